"""Autograd functions over the b200scn C ABI.  One forward / backward-input / backward-weight launch per layer.

Each function cites the upstream scn entry point it replaces and the reference call site that
reaches it.  All tensors are CUDA fp32; there is no CPU path.
"""
import collections
import os
import weakref

import torch

from . import _lib
from ._lib import alloc_rows, check, lib, ptr

# 0 = fp32 CUDA cores, 1 = TF32 tensor cores (tcgen05) where the kernel supports the shape.
# Default: the tensor-core path (stated tolerance rel 3e-3 through a net); B200SCN_PRECISION=fp32 or
# set_precision("fp32") selects the exact-fp32 CUDA-core kernels.
_precision = [0 if os.environ.get("B200SCN_PRECISION", "tf32").lower() == "fp32" else 1]


def set_precision(name):
    """'fp32' (exact-fp32 CUDA cores) or 'tf32' (tcgen05 TF32 tiles, rel 1e-3 tolerance)."""
    _precision[0] = {"fp32": 0, "tf32": 1}[name]


def get_precision():
    return "tf32" if _precision[0] else "fp32"


# ---- optional per-launch timing for bench.py's roofline block (CUDA events on the launching stream)
_prof = None


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {kind: {ms, n, bytes, flops, kernel}}; algorithmic bytes/flops per SURVEY 8d (fp32 rows touched once,
    8 B per rule pair, weights once)."""
    global _prof
    recs, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for kind, kernel, nrows_bytes, rules, per_rule_bytes, flops_per_rule, e0, e1 in recs:
        r = int(rules[1].sum()) if isinstance(rules, tuple) else int(rules)
        d = out.setdefault(kind, {"ms": 0.0, "n": 0, "bytes": 0.0, "flops": 0.0, "kernel": kernel})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
        d["bytes"] += nrows_bytes + per_rule_bytes * r
        d["flops"] += flops_per_rule * r
        sh = d.setdefault("shapes", {}).setdefault("%dx%d r%d" % (int(flops_per_rule), int(nrows_bytes), r), [0.0, 0, 0.0, 0.0])
        sh[0] += e0.elapsed_time(e1)
        sh[1] += 1
        sh[2] += nrows_bytes + per_rule_bytes * r
        sh[3] += flops_per_rule * r
    return out


_event_pool = []


def _event():
    e = _event_pool.pop() if _event_pool else torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def profile_reserve(n):
    """Pre-create n timing events so the timed region does not pay for cudaEventCreate."""
    while len(_event_pool) < n:
        _event_pool.append(torch.cuda.Event(enable_timing=True))


def _p0(kind, kernel, fixed_bytes, rules, per_rule_bytes, flops_per_rule):
    if _prof is None:
        return None
    if hasattr(rules, "rule_counts"):   # keep only the (pinned, asynchronously filled) counts, not the Level
        rules.subm_map()
        rules = ("counts", rules._counts_host)
    return (kind, kernel, fixed_bytes, rules, per_rule_bytes, flops_per_rule, _event())


def _p1(tok):
    if tok is not None:
        _prof.append(tok + (_event(),))


def _c(t):
    """fp32 contiguous-rows view: returns (tensor, leading dimension)."""
    if t.dtype != torch.float32:
        raise TypeError("b200scn computes in fp32; got %s" % t.dtype)
    if t.dim() != 2 or t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else t.shape[1])


class GemmWeight:
    """A per-offset weight stack for one GEMM direction: `w0` is the parameter (K,a,b); the GEMM multiplies rows by
    w0[k] (transposed=False: Cin_g=a, Cout_g=b) or by w0[k]^T (transposed=True: Cin_g=b, Cout_g=a), optionally with
    the offsets mirrored (flip: k -> K-1-k).  The fp32 kernels read (K,Cin_g,Cout_g), the tcgen05 kernels read the
    K-major form (K,Cout_g,Cin_g); exactly one of the two is a free view of w0, the other is one small copy."""

    def __init__(self, w0, transposed=False, flip=False, prepared=None):
        self.w0, self.transposed, self.flip = w0, transposed, flip
        self.prepared = prepared   # K-major TF32 operand made earlier (prep_both: forward makes the backward's as well)
        self.cin, self.cout = (w0.shape[2], w0.shape[1]) if transposed else (w0.shape[1], w0.shape[2])
        self.K = w0.shape[0]

    def _base(self):
        return self.w0.flip(0) if self.flip else self.w0

    def rowmajor(self):   # (K,Cin_g,Cout_g)
        w = self._base()
        return (w.transpose(1, 2) if self.transposed else w).contiguous()

    def kmajor(self):     # (K,Cout_g,Cin_g), rounded to the nearest TF32, one launch
        if self.prepared is not None:
            return self.prepared
        w0 = self.w0.contiguous()
        K, a, b = w0.shape
        out = torch.empty((K, self.cout, self.cin), dtype=torch.float32, device=w0.device)
        check(lib.b200scn_prep_weight_tf32(ptr(w0), K, a, b, 1 if self.transposed else 0, 1 if self.flip else 0,
                                           ptr(out), _lib.stream_for(w0)))
        return out


class _WeightCache:
    """Prepared (K-major, TF32-rounded) operands of every convolution weight, refreshed ONCE per optimiser step.

    Each layer keeps two persistent buffers (forward / backward-input operand).  A layer whose parameter has not changed
    since they were written (same storage, same autograd version counter) reuses them without a launch; the first layer
    that finds its parameter changed -- the first convolution after optimizer.step() -- refreshes EVERY registered layer
    in one `b200scn_prep_weight_tf32_batch` launch (cfg3: 1 launch instead of 73 per step).  In-place updates through
    autograd-visible ops (optimisers, load_state_dict, copy_) bump the version counter; edits through `.data` do not: call
    `invalidate_weight_cache()` after those.  B200SCN_WEIGHT_CACHE=0 restores one preparation launch per layer and call."""

    def __init__(self):
        self.entries = {}       # (device, data_ptr, K, a, b, flip) -> entry
        self.tables = {}        # device index -> (device table tensor, n_items, total elements, list of entries) or None

    def invalidate(self):
        for e in self.entries.values():
            e["version"] = -1

    def clear(self):
        self.entries.clear()
        self.tables.clear()

    def get(self, w, flip_bwd):
        K, a, b = w.shape
        dev = w.device.index if w.device.index is not None else torch.cuda.current_device()
        key = (dev, w.data_ptr(), K, a, b, bool(flip_bwd))
        e = self.entries.get(key)
        if e is not None and e["ref"]() is None:      # the parameter died and its address was reused
            del self.entries[key]
            self.tables.pop(dev, None)
            e = None
        if e is None:
            if len(self.entries) > 4096:                # (a process that keeps building new models)
                self.clear()
            e = {"ref": weakref.ref(w._base if w._base is not None else w), "ptr": w.data_ptr(), "shape": (K, a, b),
                 "flip": bool(flip_bwd), "dev": dev, "version": -1,
                 "fwd": torch.empty((K, b, a), dtype=torch.float32, device=w.device),
                 "bwd": torch.empty((K, a, b), dtype=torch.float32, device=w.device)}
            self.entries[key] = e
            self.tables.pop(dev, None)
            self._prep_one(w, e)
            return e["fwd"], e["bwd"]
        if e["version"] != w._version:
            self._refresh(dev, w)
        return e["fwd"], e["bwd"]

    @staticmethod
    def _alive(e):
        t = e["ref"]()
        if t is None or not t.is_cuda:
            return False
        st = t.untyped_storage()
        n = e["shape"][0] * e["shape"][1] * e["shape"][2] * 4
        return st.data_ptr() <= e["ptr"] and e["ptr"] + n <= st.data_ptr() + st.nbytes()

    def _prep_one(self, w, e):
        K, a, b = e["shape"]
        check(lib.b200scn_prep_weight_tf32_both(ptr(w), K, a, b, 1 if e["flip"] else 0, ptr(e["fwd"]), ptr(e["bwd"]),
                                                _lib.stream_for(w)))
        e["version"] = w._version

    def _refresh(self, dev, w):
        tab = self.tables.get(dev)
        if tab is None:
            live = [e for e in self.entries.values() if e["dev"] == dev and self._alive(e)]
            for k in [k for k, e in self.entries.items() if e["dev"] == dev and not self._alive(e)]:
                del self.entries[k]          # the parameter moved (model.cpu() / .cuda()) or died: its old address is not read again
            rows, first = [], 0
            for e in live:
                K, a, b = e["shape"]
                rows.append([e["ptr"], e["fwd"].data_ptr(), e["bwd"].data_ptr(), K | (a << 32), b | (int(e["flip"]) << 32), first])
                first += K * a * b
            tab = (torch.tensor(rows, dtype=torch.int64, device=w.device), len(live), first, live)
            self.tables[dev] = tab
        table, n, total, live = tab
        check(lib.b200scn_prep_weight_tf32_batch(ptr(table), n, total, _lib.stream_for(w)))
        for e in live:
            t = e["ref"]()
            if t is not None:
                e["version"] = t._version


_weight_cache = _WeightCache()
_weight_cache_on = [os.environ.get("B200SCN_WEIGHT_CACHE", "1") != "0"]


def invalidate_weight_cache():
    """Force every convolution weight to be prepared again (needed only after editing parameters through `.data`)."""
    _weight_cache.invalidate()


def set_weight_cache(on):
    _weight_cache_on[0] = bool(on)
    _weight_cache.clear()


def prep_both(w, flip_bwd):
    """K-major TF32 operands of BOTH directions of one layer (forward: w[k], backward-input: w[k]^T with the offsets
    mirrored if flip_bwd).  The forward pass gets them together and hands the second to its backward; with the weight cache
    (default) they are persistent buffers refreshed once per optimiser step for all layers in one launch, otherwise one
    launch per layer and call.  None when the tensor-core path is off."""
    if _precision[0] != 1:
        return None, None
    cacheable = _weight_cache_on[0] and w.is_contiguous() and w.data_ptr() % 16 == 0   # (a non-contiguous view is a new copy per call)
    w = w.contiguous()
    K, a, b = w.shape
    if lib.b200scn_gather_conv_tf32_ok(a, b, a) != 1 or lib.b200scn_gather_conv_tf32_ok(b, a, b) != 1:
        return None, None
    if cacheable:
        return _weight_cache.get(w, flip_bwd)
    fwd = torch.empty((K, b, a), dtype=torch.float32, device=w.device)
    bwd = torch.empty((K, a, b), dtype=torch.float32, device=w.device)
    check(lib.b200scn_prep_weight_tf32_both(ptr(w), K, a, b, 1 if flip_bwd else 0, ptr(fwd), ptr(bwd), _lib.stream_for(w)))
    return fwd, bwd


def _use_tf32(gw, lda, x):
    return _precision[0] == 1 and x.data_ptr() % 16 == 0 and lib.b200scn_gather_conv_tf32_ok(gw.cin, gw.cout, lda) == 1


# Spatially tiled submanifold convolution (conv_halo.cu): halo capacity per 128-row tile and the smallest level it is
# used for (below that the step is bound by host launch overhead, not by the kernel); B200SCN_HALO=1 forces it on for every
# size, =0 off.
_halo = {"hcap": int(os.environ.get("B200SCN_HALO_CAP", "384")), "min_rows": 128 * 148,
         "mode": os.environ.get("B200SCN_HALO", "")}   # the environment is read once, at import


def set_halo_capacity(hcap):
    _halo["hcap"] = int(hcap)


def set_tiled(mode):
    """'auto' (default: levels with >= 148 tiles), 'on' / 'off' (tests and experiments)."""
    _halo["mode"] = {"auto": "", "on": "1", "off": "0", "": "", "1": "1", "0": "0"}[mode]


def _use_tiled(n):
    mode = _halo["mode"]
    if mode == "0":
        return False
    return mode == "1" or n >= _halo["min_rows"]


def subm_conv(x, level, gw, addend=None, round_a=True):
    """out[o] = sum_k x[nbr[o,k]] @ Wg[k] over the level's 3x3x3 neighbour map (forward, and backward-input with the
    mirrored transposed weights)."""
    x, ldx = _c(x)
    if not (_use_tf32(gw, ldx, x) and _use_tiled(level.n)):
        return gather_conv(x, level.subm_map(), level.n, 27, gw, addend=addend, rules=level)
    Cin, Cout = gw.cin, gw.cout
    plan = level.tile_plan(_halo["hcap"])
    out = alloc_rows(level.n, Cout, x.device)
    lda = 0
    if addend is not None:
        addend, lda = _c(addend)
    w = gw.kmajor()
    tok = _p0("tiled27", "subm_conv_tiled", 4.0 * (x.shape[0] * Cin + level.n * Cout) + 4.0 * 27 * Cin * Cout,
              level, 8.0, 2.0 * Cin * Cout)
    check(lib.b200scn_subm_conv_tiled(ptr(x), ldx, ptr(level.nbr), ptr(plan.perm), ptr(plan.lmap), ptr(plan.halo_ids),
                                      ptr(plan.halo_n), ptr(plan.kmask), plan.hcap, level.n, ptr(w), Cin, Cout,
                                      ptr(addend), lda, ptr(out), Cout, 1 if round_a else 0, _lib.stream_for(x)))
    _p1(tok)
    return out


def gather_conv(x, map_t, n_out, K, gw, addend=None, rules=None):
    """out[o] = sum_k x[map[o,k]] @ Wg[k] (+ addend[o]);  gw: GemmWeight.  `rules`: rule count (or Level) for accounting."""
    x, ldx = _c(x)
    Cin, Cout = gw.cin, gw.cout
    out = alloc_rows(n_out, Cout, x.device)
    lda = 0
    if addend is not None:
        addend, lda = _c(addend)
    tf32 = _use_tf32(gw, ldx, x)
    w = gw.kmajor() if tf32 else gw.rowmajor()
    tok = _p0("gather%d" % K, "conv_gather", 4.0 * (x.shape[0] * Cin + n_out * Cout) + 4.0 * K * Cin * Cout,
              n_out if rules is None else rules, 8.0 if map_t is not None else 0.0, 2.0 * Cin * Cout)
    check(lib.b200scn_gather_conv(ptr(x), ldx, x.shape[0], ptr(map_t), n_out, K, ptr(w), Cin, Cout, ptr(addend), lda,
                                  ptr(out), Cout, 1 if tf32 else 0, _lib.stream_for(x)))
    _p1(tok)
    return out


def scatter_conv(x, map_t, n_out, K, gw, down=None):
    """out[map[j,k]] = x[j] @ Wg[k]; every out row is addressed exactly once by a strided child map.
    On the tensor-core path the rules are taken offset-sorted (the scn-form rulebook) and every run of <= 128 rules of one
    offset is one dense tile (b200scn_grouped_conv)."""
    x, ldx = _c(x)
    Cin, Cout = gw.cin, gw.cout
    out = alloc_rows(n_out, Cout, x.device)
    if down is not None and _use_tf32(gw, ldx, x):
        fine_ids, coarse_ids, tab, max_tiles = down.group_tiles()
        w = gw.kmajor()
        tok = _p0("grouped%d" % K, "grouped_conv", 4.0 * (x.shape[0] * Cin + n_out * Cout) + 4.0 * K * Cin * Cout,
                  n_out, 8.0, 2.0 * Cin * Cout)
        check(lib.b200scn_grouped_conv(ptr(x), ldx, ptr(coarse_ids), ptr(fine_ids), ptr(tab), max_tiles, n_out, K, ptr(w),
                                       Cin, Cout, ptr(out), Cout, _lib.stream_for(x)))
        _p1(tok)
        return out
    w = gw.rowmajor()
    tok = _p0("scatter%d" % K, "conv_scatter", 4.0 * (x.shape[0] * Cin + n_out * Cout) + 4.0 * K * Cin * Cout,
              n_out, 8.0, 2.0 * Cin * Cout)
    check(lib.b200scn_scatter_conv(ptr(x), ldx, ptr(map_t), x.shape[0], K, ptr(w), Cin, Cout, ptr(out), Cout,
                                   0, _lib.stream_for(x)))
    _p1(tok)
    return out


def pair_dw(a, g, pair_a, pair_g, offsets, K, n_pairs_max, rules=None):
    """dW[k] = sum_p a[pair_a[p]]^T (x) g[pair_g[p]] over list k."""
    a, lda = _c(a)
    g, ldg = _c(g)
    Ca, Cg = a.shape[1], g.shape[1]
    dw = torch.empty((K, Ca, Cg), dtype=torch.float32, device=a.device)
    tok = _p0("pair_dw%d" % K, "pair_dw", 4.0 * (a.shape[0] * Ca + g.shape[0] * Cg) + 4.0 * K * Ca * Cg,
              n_pairs_max if rules is None else rules, 8.0 if pair_a is not None else 0.0, 2.0 * Ca * Cg)
    check(lib.b200scn_pair_dw(ptr(a), lda, ptr(g), ldg, ptr(pair_a), ptr(pair_g), ptr(offsets), K, n_pairs_max,
                              Ca, Cg, ptr(dw), _precision[0], _lib.stream_for(a)))
    _p1(tok)
    return dw


def pair_dw_blocked(a, g, pair_a, pair_g, blk, nblk, K, rules=None):
    """Same over a b200scn_pair_lists_blocked table: one CTA per (offset, row block).  None if the shape is not taken."""
    a, lda = _c(a)
    g, ldg = _c(g)
    Ca, Cg = a.shape[1], g.shape[1]
    if not (_precision[0] == 1 and Ca % 4 == 0 and Ca >= 4 and Cg % 16 == 0 and 16 <= Cg <= 256 and lda % 4 == 0
            and ldg % 4 == 0 and a.data_ptr() % 16 == 0 and g.data_ptr() % 16 == 0):   # = pair_dw_tc_supported (conv_tc.cu)
        return None
    dw = torch.empty((K, Ca, Cg), dtype=torch.float32, device=a.device)
    tok = _p0("pair_dw%d" % K, "pair_dw", 4.0 * (a.shape[0] * Ca + g.shape[0] * Cg) + 4.0 * K * Ca * Cg, rules, 8.0,
              2.0 * Ca * Cg)
    check(lib.b200scn_pair_dw_blocked(ptr(a), lda, ptr(g), ldg, ptr(pair_a), ptr(pair_g), ptr(blk), K, nblk, Ca, Cg, ptr(dw),
                                      _lib.stream_for(a)))
    _p1(tok)
    return dw


_dw_blocked = [False]


def set_blocked_dw(on):
    """Weight gradient of tiled levels over row-block-aligned pair segments.  Default OFF: measured equal to the per-offset
    chunks at levels 1-2 and 15-40 % slower at levels 0 and 3 (profiles/r2d_time_dw.txt) -- the kernel is bound by its
    per-stage synchronisation chain, not by where the gathered rows come from."""
    _dw_blocked[0] = bool(on)


_dw_tiled = [False]


def set_tiled_dw(on):
    """Tile-stationary, bit-reproducible weight gradient (dw_tile.cu: A^T operand built in tensor memory, G tile staged once
    per tile, fixed-order reduction) on levels where the tiled forward kernel runs.  Default OFF: the deterministic
    option -- measured 1.3x (level 0, 32x32) to 2.1-3.5x (levels 1-2) slower than the pair-list kernel
    (profiles/r2f_time_dw_tile.txt): TMEM capacity forces 4-12 CTAs to re-stage every tile (one 32-channel block and a
    subset of the offsets each) and the loaders' two dependent global round trips per tile bound it."""
    _dw_tiled[0] = bool(on)


def subm_dw_tiled(a, g, level):
    """dW (27,Ca,Cg) of a submanifold convolution over the level's tile plan, or None if the shape is not taken."""
    a, lda = _c(a)
    g, ldg = _c(g)
    Ca, Cg = a.shape[1], g.shape[1]
    hcap = _halo["hcap"]
    nbytes = lib.b200scn_subm_dw_tiled_scratch_bytes(level.n, hcap, Ca, Cg)
    if nbytes == 0 or a.data_ptr() % 16 or g.data_ptr() % 16 or lda % 4 or ldg % 4:
        return None
    plan = level.tile_plan(hcap)
    dw = torch.empty((27, Ca, Cg), dtype=torch.float32, device=a.device)
    scratch = torch.empty(nbytes // 4, dtype=torch.float32, device=a.device)
    tok = _p0("tile_dw27", "subm_dw_tiled", 4.0 * (a.shape[0] * Ca + g.shape[0] * Cg) + 4.0 * 27 * Ca * Cg, level, 8.0,
              2.0 * Ca * Cg)
    check(lib.b200scn_subm_dw_tiled(ptr(a), lda, ptr(g), ldg, ptr(level.nbr), ptr(plan.perm), ptr(plan.lmap),
                                    ptr(plan.halo_ids), ptr(plan.halo_n), hcap, level.n, Ca, Cg, ptr(dw), ptr(scratch),
                                    nbytes, _lib.stream_for(a)))
    _p1(tok)
    return dw


# Experiment switch (B200SCN_DW_SIDE=1 / set_dw_side_stream): the weight gradient of a submanifold layer on a second stream,
# concurrent with the layer's input gradient (they are independent); both streams are joined before backward returns.
_dw_side = [os.environ.get("B200SCN_DW_SIDE", "0") == "1"]
_side_streams = {}


def set_dw_side_stream(on):
    _dw_side[0] = bool(on)


# Deferred weight gradients (set_deferred_dw, opt-in): every submanifold layer's weight gradient is queued on the second
# stream and NOT waited for by the layer; the streams are joined once, by an autograd-engine callback at the end of the
# backward pass; the gradients are stored into `.grad` directly (they bypass AccumulateGrad, whose copy would otherwise read
# them on the main stream too early) and the parameters' post-accumulate hooks are run by hand.  The weight gradients leave the critical path: they fill the SMs that the small
# levels of the U-Net (1.5 % of the voxels, ~2.5 ms of nearly idle GPU per backward) and the tails of the big kernels leave
# empty.  Contract: loss.backward() still returns with every .grad complete; torch.autograd.grad() and hooks on the
# weight gradients (other than post-accumulate hooks) are not supported in this mode.
_dw_defer = [False]
_deferred = {"queued": False, "device": None, "held": collections.deque()}


def set_deferred_dw(on):
    _dw_defer[0] = bool(on)


def _finish_deferred():
    """End of the backward pass (autograd-engine callback): the main stream waits for the deferred weight gradients."""
    dev = _deferred["device"]
    _deferred["queued"] = False
    if dev is not None:
        torch.cuda.current_stream(dev).wait_stream(_side_stream(dev))
    _deferred["held"].clear()      # (their memory is protected by record_stream until the second stream is done with it)


def recover_deferred():
    """Called at the start of every forward pass: if the previous backward pass raised before the engine ran its end-of-pass
    callback, join the streams now and forget the stale state (no-op otherwise)."""
    if _deferred["queued"]:
        _finish_deferred()


def _deliver_deferred(param, dw, side):
    """Store a deferred weight gradient (still being computed on `side`) into param.grad.  Autograd still runs the
    parameter's post-accumulate hooks when its AccumulateGrad node is reached with the (undefined) gradient this layer
    returns -- the data-parallel bucket hooks of b200scn_dp.FlatGrads, which in this mode pack and launch their collectives
    on `side` too, so nothing on the main stream ever waits for a weight gradient before the end of backward."""
    dw = dw.view_as(param)
    with torch.no_grad():
        if param.grad is None:
            param.grad = dw
        else:
            with torch.cuda.stream(side):
                param.grad.add_(dw)


def _param_of(w):
    """The Parameter behind a weight view (the layers pass `weight.view(...)` / `weight.unsqueeze(0)`), or None."""
    b = w._base if w._base is not None else w
    return b if isinstance(b, torch.nn.Parameter) and b.numel() == w.numel() else None


def _defer_ok(param):
    return (_dw_defer[0] and _precision[0] == 1 and param is not None and not _dw_tiled[0] and not _dw_blocked[0]
            and _prof is None)


def _run_deferred(param, inputs, fn):
    """Queue fn() -> dW on the second stream after everything the main stream has enqueued so far, and deliver it."""
    dev = inputs[0].device
    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        dw = fn()
    # the inputs belong to the main stream's allocator: it must not hand their memory out again before the second stream
    # has read them; dw is read by the main stream (optimiser) after the join
    for t in inputs:
        t.record_stream(side)
    dw.record_stream(main)
    # keep the inputs referenced until the second stream has consumed them: the autograd engine accumulates IN PLACE into
    # gradient buffers it solely owns (a layer may hand its incoming gradient on as the gradient of its addend), which
    # would change `g` under the queued kernel; a second reference makes it allocate instead
    ev = torch.cuda.Event()
    ev.record(side)
    held = _deferred["held"]
    held.append((ev, inputs))
    while held and held[0][0].query():
        held.popleft()
    _deferred["device"] = dev
    if not _deferred["queued"]:
        _deferred["queued"] = True
        torch.autograd.Variable._execution_engine.queue_callback(_finish_deferred)
    _deliver_deferred(param, dw, side)


def _side_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _side_streams:
        _side_streams[idx] = torch.cuda.Stream(device=idx)
    return _side_streams[idx]


class SubmanifoldConvFn(torch.autograd.Function):
    """scn.SubmanifoldConvolution (models/SparseConvNet.py:62,117,119): replaces upstream
    SubmanifoldConvolution_updateOutput / _backward."""

    @staticmethod
    def forward(ctx, x, w, level, addend=None, x_rounded=False):
        ctx.level = level
        ctx.param = _param_of(w)                        # (deferred-dW mode delivers the weight gradient to it directly)
        ctx.save_for_backward(x, w)
        fwd, ctx.w_bwd = prep_both(w, True) if (ctx.needs_input_grad[0] or _weight_cache_on[0]) else (None, None)
        return subm_conv(x, level, GemmWeight(w, prepared=fwd), addend=addend, round_a=not x_rounded)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        level = ctx.level
        dx = dw = None
        tiled = _precision[0] == 1 and _use_tiled(level.n)
        base = ctx.param
        if ctx.needs_input_grad[1] and _defer_ok(base):
            if tiled:
                pin, pout, offs = level.subm_pairs_ordered(level.tile_plan(_halo["hcap"]).perm)
            else:
                pin, pout, offs = level.subm_pairs()
            _run_deferred(base, (x, g), lambda: pair_dw(x, g, pin, pout, offs, 27, level.n, rules=level))
            if ctx.needs_input_grad[0]:
                dx = subm_conv(g, level, GemmWeight(w, transposed=True, flip=True, prepared=ctx.w_bwd))
            return dx, None, None, (g if ctx.needs_input_grad[3] else None), None
        if (_dw_side[0] and tiled and ctx.needs_input_grad[0] and ctx.needs_input_grad[1] and not _dw_tiled[0]
                and not _dw_blocked[0] and _prof is None):
            # rulebook first, on the main stream (it is cached on the level and used by later layers on either stream)
            pin, pout, offs = level.subm_pairs_ordered(level.tile_plan(_halo["hcap"]).perm)
            main, side = torch.cuda.current_stream(g.device), _side_stream(g.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dw = pair_dw(x, g, pin, pout, offs, 27, level.n, rules=level)
            dx = subm_conv(g, level, GemmWeight(w, transposed=True, flip=True, prepared=ctx.w_bwd))
            main.wait_stream(side)
            dw.record_stream(main)
            return dx, dw, None, (g if ctx.needs_input_grad[3] else None), None
        if ctx.needs_input_grad[0]:
            # pair (in=i, out=o) at offset k <=> o = nbr[i][26-k]:  dx[i] = sum_k' g[nbr[i][k']] @ w[26-k']^T
            dx = subm_conv(g, level, GemmWeight(w, transposed=True, flip=True, prepared=ctx.w_bwd))
        if ctx.needs_input_grad[1]:
            if tiled and _dw_tiled[0]:
                dw = subm_dw_tiled(x, g, level)
            if dw is None and tiled and _dw_blocked[0]:
                pin, pout, offs, (blk, nblk) = level.subm_pairs_blocked(level.tile_plan(_halo["hcap"]).perm)
                dw = pair_dw_blocked(x, g, pin, pout, blk, nblk, 27, rules=level)
            if dw is None:
                if tiled:
                    pin, pout, offs = level.subm_pairs_ordered(level.tile_plan(_halo["hcap"]).perm)
                else:
                    pin, pout, offs = level.subm_pairs()
                dw = pair_dw(x, g, pin, pout, offs, 27, level.n, rules=level)
        return dx, dw, None, (g if ctx.needs_input_grad[3] else None), None


class ConvolutionFn(torch.autograd.Function):
    """scn.Convolution size==stride (models/SparseConvNet.py:137-138): Convolution_updateOutput / _backward."""

    @staticmethod
    def forward(ctx, x, w, down):
        ctx.down = down
        ctx.param = _param_of(w)
        ctx.save_for_backward(x, w)
        fwd, ctx.w_bwd = prep_both(w, False) if (ctx.needs_input_grad[0] or _weight_cache_on[0]) else (None, None)
        return gather_conv(x, down.child_map(), down.coarse.n, down.K, GemmWeight(w, prepared=fwd), rules=down.fine.n)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        down = ctx.down
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = scatter_conv(g, down.child_map(), down.fine.n, down.K, GemmWeight(w, transposed=True, prepared=ctx.w_bwd), down)
        if ctx.needs_input_grad[1]:
            pin, pout, offs = down.child_pairs()
            if _defer_ok(ctx.param):
                _run_deferred(ctx.param, (x, g), lambda: pair_dw(x, g, pin, pout, offs, down.K, down.coarse.n, rules=down.fine.n))
            else:
                dw = pair_dw(x, g, pin, pout, offs, down.K, down.coarse.n, rules=down.fine.n)
        return dx, dw, None


class DeconvolutionFn(torch.autograd.Function):
    """scn.Deconvolution (inside scn.UNet, models/SparseConvNet.py:63-68): Deconvolution_updateOutput / _backward,
    on the rulebook the matching Convolution built, roles swapped."""

    @staticmethod
    def forward(ctx, x, w, down):
        ctx.down = down
        ctx.param = _param_of(w)
        ctx.save_for_backward(x, w)
        fwd, ctx.w_bwd = prep_both(w, False) if (ctx.needs_input_grad[0] or _weight_cache_on[0]) else (None, None)
        return scatter_conv(x, down.child_map(), down.fine.n, down.K, GemmWeight(w, prepared=fwd), down)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        down = ctx.down
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = gather_conv(g, down.child_map(), down.coarse.n, down.K, GemmWeight(w, transposed=True, prepared=ctx.w_bwd),
                             rules=down.fine.n)
        if ctx.needs_input_grad[1]:
            pin, pout, offs = down.child_pairs()
            if _defer_ok(ctx.param):
                _run_deferred(ctx.param, (x, g), lambda: pair_dw(x, g, pout, pin, offs, down.K, down.coarse.n, rules=down.fine.n))
            else:
                dw = pair_dw(x, g, pout, pin, offs, down.K, down.coarse.n, rules=down.fine.n)
        return dx, dw, None


class UnPoolingFn(torch.autograd.Function):
    """scn.UnPooling (models/SparseConvNet.py:140,193): UnPooling_updateOutput / _updateGradInput."""

    @staticmethod
    def forward(ctx, x, down):
        ctx.down = down
        x, ldx = _c(x)
        C = x.shape[1]
        out = alloc_rows(down.fine.n, C, x.device)
        tok = _p0("unpool", "unpool", 4.0 * (down.fine.n + down.coarse.n) * C + 4.0 * down.fine.n, 0, 0.0, 0.0)
        check(lib.b200scn_unpool(ptr(x), ldx, ptr(down.parent), down.fine.n, C, ptr(out), C, _lib.stream_for(x)))
        _p1(tok)
        return out

    @staticmethod
    def backward(ctx, g):
        down = ctx.down
        g, ldg = _c(g)
        C = g.shape[1]
        dx = alloc_rows(down.coarse.n, C, g.device)
        tok = _p0("unpool_bwd", "unpool_bwd", 4.0 * (down.fine.n + down.coarse.n) * C + 4.0 * down.fine.n, 0, 0.0, 0.0)
        check(lib.b200scn_unpool_bwd(ptr(g), ldg, ptr(down.child_map()), down.coarse.n, down.K, C, ptr(dx), C,
                                     _lib.stream_for(g)))
        _p1(tok)
        return dx, None


class NetworkInNetworkFn(torch.autograd.Function):
    """scn.NetworkInNetwork (models/SparseConvNet.py:114): NetworkInNetwork_updateOutput / updateGradInput /
    accGradParameters -- a one-offset convolution with the identity rulebook."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.param = w if isinstance(w, torch.nn.Parameter) else _param_of(w)
        ctx.save_for_backward(x, w)
        fwd, ctx.w_bwd = prep_both(w.unsqueeze(0), False) if (ctx.needs_input_grad[0] or _weight_cache_on[0]) else (None, None)
        return gather_conv(x, None, x.shape[0], 1, GemmWeight(w.unsqueeze(0), prepared=fwd))

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = gather_conv(g, None, g.shape[0], 1, GemmWeight(w.unsqueeze(0), transposed=True, prepared=ctx.w_bwd))
        if ctx.needs_input_grad[1]:
            if _defer_ok(ctx.param):
                _run_deferred(ctx.param, (x, g), lambda: pair_dw(x, g, None, None, None, 1, x.shape[0])[0])
            else:
                dw = pair_dw(x, g, None, None, None, 1, x.shape[0])[0]
        return dx, dw


_bn_scratch = {}


def bn_scratch(device):
    """The persistent, zero-initialised, self-cleaning statistics scratch of b200scn_bn_forward / _backward for the current
    stream of `device` (one buffer per (device, stream): calls on one stream are ordered, so they can share it)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _bn_scratch.get(key)
    if buf is None:
        buf = _bn_scratch[key] = torch.zeros(lib.b200scn_bn_scratch_doubles(1 << 20), dtype=torch.float64, device=device)
    return buf


def _bn_forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leak, round_tf32):
    x, ldx = _c(x)
    n, C = x.shape
    dev = x.device
    y = alloc_rows(n, C, dev)
    save_mean = torch.empty(C, dtype=torch.float32, device=dev)
    save_invstd = torch.empty(C, dtype=torch.float32, device=dev)
    scratch = bn_scratch(dev)
    tok = _p0("bn_fwd", "bn", 12.0 * n * C, 0, 0.0, 0.0)
    check(lib.b200scn_bn_forward(ptr(x), ldx, n, C, ptr(weight), ptr(bias), ptr(running_mean), ptr(running_var),
                                 ptr(save_mean), ptr(save_invstd), eps, momentum, 1 if train else 0, leak,
                                 ptr(y), C, ptr(scratch), 1 if (round_tf32 and _precision[0] == 1) else 0,
                                 _lib.stream_for(x)))
    _p1(tok)
    ctx.save_for_backward(x, weight, bias, save_mean, save_invstd)
    ctx.leak, ctx.train = leak, bool(train)
    return y


def _bn_backward(ctx, g, addend):
    x, weight, bias, save_mean, save_invstd = ctx.saved_tensors
    x, ldx = _c(x)
    g, ldg = _c(g)
    n, C = x.shape
    dev = x.device
    lda = 0
    if addend is not None:
        addend, lda = _c(addend)
    dx = alloc_rows(n, C, dev)
    dweight = torch.empty(C, dtype=torch.float32, device=dev)
    dbias = torch.empty(C, dtype=torch.float32, device=dev)
    scratch = bn_scratch(dev)
    tok = _p0("bn_bwd", "bn", (24.0 if addend is not None else 20.0) * n * C, 0, 0.0, 0.0)
    check(lib.b200scn_bn_backward(ptr(x), ldx, ptr(g), ldg, n, C, ptr(weight), ptr(bias), ptr(save_mean),
                                  ptr(save_invstd), ctx.leak, 1 if ctx.train else 0, ptr(addend), lda, ptr(dx), C,
                                  ptr(dweight), ptr(dbias), ptr(scratch), _lib.stream_for(x)))
    _p1(tok)
    return dx, dweight, dbias


class BatchNormFn(torch.autograd.Function):
    """scn.BatchNormReLU / BatchNormLeakyReLU (models/SparseConvNet.py:69,116,118,136):
    BatchNormalization_updateOutput / _backward (eps 1e-4, momentum 0.9 on the old value, App. B.8)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leak, round_tf32=False):
        return _bn_forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leak, round_tf32)

    @staticmethod
    def backward(ctx, g):
        dx, dweight, dbias = _bn_backward(ctx, g, None)
        return dx, dweight, dbias, None, None, None, None, None, None, None


class BatchNormSkipFn(torch.autograd.Function):
    """BatchNorm at the head of a residual branch: returns (bn(x), x).  The second output IS x, handed to the skip
    connection; in backward the gradient that comes back through it is added inside the BatchNorm backward kernel
    (one pass) instead of by an autograd add kernel over the whole tensor (26 such adds per step in the m=32 UNet)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leak, round_tf32=False):
        y = _bn_forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leak, round_tf32)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, g, gskip):
        if g is None:   # only the skip was used
            return gskip, None, None, None, None, None, None, None, None, None
        dx, dweight, dbias = _bn_backward(ctx, g, gskip)
        return dx, dweight, dbias, None, None, None, None, None, None, None


class InputFeaturesFn(torch.autograd.Function):
    """scn.InputLayer feature half (models/SparseConvNet.py:61): InputLayer_updateOutput / _updateGradInput."""

    @staticmethod
    def forward(ctx, feats, md, n0):
        feats = feats.contiguous()
        if feats.dtype != torch.float32:
            raise TypeError("InputLayer: features must be float32, got %s" % feats.dtype)
        P, C = feats.shape
        out = alloc_rows(n0, C, feats.device, zero=True)
        tok = _p0("input_feats", "input_features", 4.0 * (P + n0) * C + 4.0 * P, 0, 0.0, 0.0)
        check(lib.b200scn_input_features(ptr(feats), P, C, ptr(md.pv), ptr(md.count), ptr(md.first_row),
                                         ptr(md.last_row), md.mode, ptr(out), _lib.stream_for(feats)))
        _p1(tok)
        ctx.md = md
        return out

    @staticmethod
    def backward(ctx, g):
        md = ctx.md
        g = g.contiguous()
        C = g.shape[1]
        d = alloc_rows(md.P, C, g.device)
        check(lib.b200scn_input_features_bwd(ptr(g), md.P, C, ptr(md.pv), ptr(md.count), ptr(md.first_row),
                                             ptr(md.last_row), md.mode, ptr(d), _lib.stream_for(g)))
        return d, None, None


class OutputFeaturesFn(torch.autograd.Function):
    """scn.OutputLayer (models/SparseConvNet.py:70): OutputLayer_updateOutput / _updateGradInput."""

    @staticmethod
    def forward(ctx, feats, md):
        feats, ldf = _c(feats)
        C = feats.shape[1]
        out = alloc_rows(md.P, C, feats.device)
        tok = _p0("output_feats", "output_features", 4.0 * (feats.shape[0] + md.P) * C + 4.0 * md.P, 0, 0.0, 0.0)
        check(lib.b200scn_output_features(ptr(feats), ldf, md.P, C, ptr(md.pv), ptr(md.first_row), ptr(md.last_row),
                                          md.mode, ptr(out), _lib.stream_for(feats)))
        _p1(tok)
        ctx.md, ctx.n = md, feats.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        md = ctx.md
        g = g.contiguous()
        C = g.shape[1]
        d = alloc_rows(ctx.n, C, g.device)
        start, rows = md.site_rows()
        tok = _p0("output_feats_bwd", "output_features_bwd", 4.0 * (ctx.n + md.P) * C + 4.0 * md.P, 0, 0.0, 0.0)
        check(lib.b200scn_output_features_bwd_csr(ptr(g), ctx.n, C, ptr(start), ptr(md.count), ptr(rows),
                                                  ptr(md.first_row), ptr(md.last_row), md.mode, ptr(d), C,
                                                  _lib.stream_for(g)))
        _p1(tok)
        return d, None


class SceneMeanFn(torch.autograd.Function):
    """Per-scene mean of the per-point features (SparseConvBase_.postProcessing, models/SparseConvNet.py:20-26) straight
    from the level-0 voxel features: OutputLayer + the Python loop of torch.mean without the (sum P, C) tensor."""

    @staticmethod
    def forward(ctx, feats, md, level, batch_size):
        feats, ldf = _c(feats)
        n, C = feats.shape
        out = torch.empty((batch_size, C), dtype=torch.float32, device=feats.device)
        npts = torch.empty(batch_size, dtype=torch.float32, device=feats.device)
        check(lib.b200scn_scene_mean(ptr(feats), ldf, ptr(level.ukeys), ptr(md.count), md.mode, n, C, batch_size,
                                     ptr(out), ptr(npts), _lib.stream_for(feats)))
        ctx.md, ctx.level, ctx.n = md, level, n
        ctx.save_for_backward(npts)
        return out

    @staticmethod
    def backward(ctx, g):
        npts, = ctx.saved_tensors
        g = g.contiguous()
        C = g.shape[1]
        d = alloc_rows(ctx.n, C, g.device)
        check(lib.b200scn_scene_mean_bwd(ptr(g), ptr(ctx.level.ukeys), ptr(ctx.md.count), ctx.md.mode, ptr(npts), ctx.n, C,
                                         ptr(d), C, _lib.stream_for(g)))
        return d, None, None, None


class MaxPoolingFn(torch.autograd.Function):
    """scn.MaxPooling, size == stride (models/projector/components.py:78-100): MaxPooling_updateOutput / _updateGradInput."""

    @staticmethod
    def forward(ctx, x, down):
        x, ldx = _c(x)
        C = x.shape[1]
        out = alloc_rows(down.coarse.n, C, x.device)
        check(lib.b200scn_maxpool(ptr(x), ldx, ptr(down.child_map()), down.coarse.n, down.K, C, ptr(out), C,
                                  _lib.stream_for(x)))
        ctx.down = down
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        down = ctx.down
        x, ldx = _c(x)
        g, ldg = _c(g)
        C = x.shape[1]
        d = alloc_rows(down.fine.n, C, x.device)
        check(lib.b200scn_maxpool_bwd(ptr(g), ldg, ptr(x), ldx, ptr(out), C, ptr(down.parent), down.fine.n, C, ptr(d), C,
                                      _lib.stream_for(x)))
        return d, None


class SparseToDenseFn(torch.autograd.Function):
    """scn.SparseToDense (Function_test.py:46): SparseToDense_updateOutput / _updateGradInput."""

    @staticmethod
    def forward(ctx, x, level, batch_size):
        x, ldx = _c(x)
        n, C = x.shape
        S = level.size
        dense = torch.zeros((batch_size, C, S, S, S), dtype=torch.float32, device=x.device)
        check(lib.b200scn_sparse_to_dense(ptr(x), ldx, ptr(level.ukeys), n, C, S, ptr(dense), _lib.stream_for(x)))
        ctx.level, ctx.n = level, n
        return dense

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        C = g.shape[1]
        d = alloc_rows(ctx.n, C, g.device)
        check(lib.b200scn_sparse_to_dense_bwd(ptr(g), ptr(ctx.level.ukeys), ctx.n, C, ctx.level.size, ptr(d), C,
                                              _lib.stream_for(g)))
        return d, None, None
