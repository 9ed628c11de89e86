"""Device-resident Metadata: active-site grids and rulebooks for one forward/backward.

Mirrors the role of upstream scn's `Metadata_3`: one object per InputLayer call,
shared by reference by every SparseConvNetTensor derived from it, owning every grid / rulebook.
Unlike upstream (CPU hash maps, rules copied H2D on every conv call) everything lives in HBM and
the host synchronises exactly once per forward (to learn the per-level site counts).
"""
import torch

from . import _lib
from ._lib import alloc_flat, alloc_rows, check, lib, ptr

# speculative pyramid built inside InputLayer before its single host sync: (stride, coarse levels)
_pyramid_hint = [2, 6]


def set_pyramid_hint(stride, depth):
    """Stride and number of coarse levels pre-built per forward (default 2, 6 = the 7-level UNet/FCNet).
    Nets with another stride (e.g. downsample=[4,4]) still work: their pyramid is built on demand with
    one extra host sync."""
    _pyramid_hint[0], _pyramid_hint[1] = int(stride), int(depth)


class PackedKeys:
    """Level-0 coordinates already on the device in the library's packed form (b << 48 | x << 32 | y << 16 | z, int64, one per
    point, all inside the spatial size): what b200scn_augment_voxelize produces.  `InputLayer` accepts it in place of the
    (sum P, 3|4) LongTensor and skips the packing / range check."""

    def __init__(self, keys, batch_size=None):
        assert keys.dtype == torch.int64 and keys.dim() == 1 and keys.is_cuda
        self.keys, self.batch_size = keys, batch_size

    def size(self, dim=None):
        return self.keys.shape[0] if dim in (0, None) else 4

    @property
    def shape(self):
        return (self.keys.shape[0], 4)


class Level:
    """One spatial size: its sites (keys in id order), hash table, lazily built 3^3 neighbour map and pair lists."""

    def __init__(self, size, n, ukeys, hkeys, hvals, cap):
        self.size, self.n, self.ukeys, self.hkeys, self.hvals, self.cap = size, n, ukeys, hkeys, hvals, cap
        self.nbr = None           # (n,27) int32
        self.nbr_counts = None    # (27,) int32 device: rules per offset
        self.pairs = None         # (pair_in, pair_out, offsets_dev)
        self.pairs_ordered = None
        self.pairs_blocked = None
        self._counts = None
        self.plan = None          # TilePlan of the spatially tiled convolution
        self.batch_bits = 15      # sample-index bits sorted by the Morton ordering (InputLayer narrows it to the batch)

    def tile_plan(self, hcap):
        """Morton-ordered 128-row tiles + per-tile halo lists (b200scn_tile_plan), built once per level and step."""
        if self.plan is None or self.plan.hcap != hcap:
            self.plan = TilePlan(self, hcap)
        return self.plan

    def subm_map(self):
        if self.nbr is None:
            dev = self.ukeys.device
            self.nbr = alloc_rows(self.n, 27, dev, torch.int32)
            self.nbr_counts = torch.zeros(27, dtype=torch.int32, device=dev)
            st = _lib.stream_for(self.ukeys)
            from . import ops as _ops
            tok = _ops._p0("subm_map", "subm_map", 8.0 * self.n + 4.0 * 27 * self.n, 0, 0.0, 0.0)   # keys read + map written
            check(lib.b200scn_subm_map(ptr(self.ukeys), self.n, None, ptr(self.hkeys), ptr(self.hvals), self.cap,
                                       self.size, ptr(self.nbr), ptr(self.nbr_counts), st))
            _ops._p1(tok)
            # rule counts travel to the host asynchronously; nobody waits for them unless the op counters
            # are read or a weight gradient needs exact pair-list sizes (long after this point)
            self._counts_host = torch.empty(27, dtype=torch.int32, pin_memory=True)
            self._counts_host.copy_(self.nbr_counts, non_blocking=True)
            self._counts_event = torch.cuda.Event()
            self._counts_event.record(torch.cuda.current_stream(dev))
        return self.nbr

    def rule_counts(self):
        """Rules per kernel offset (host ints)."""
        self.subm_map()
        if self._counts is None:
            self._counts_event.synchronize()
            self._counts = [int(v) for v in self._counts_host]
        return self._counts

    def subm_pairs(self):
        if self.pairs is None:
            self.pairs = build_pairs(self.subm_map(), self.n, 27, sum(self.rule_counts()))
        return self.pairs

    def subm_pairs_ordered(self, perm):
        """The same pairs with every offset's list walking the sites along the Morton curve (weight gradient only:
        CTAs that run at the same time then gather rows that are neighbours in space and hit in L2)."""
        if self.pairs_ordered is None:
            self.pairs_ordered = build_pairs(self.subm_map(), self.n, 27, sum(self.rule_counts()), order=perm)
        return self.pairs_ordered

    def subm_pairs_blocked(self, perm):
        """Morton-ordered lists plus the row-block table: -> (pair_in, pair_out, offsets, (blk_offsets, nblk))."""
        if self.pairs_blocked is None:
            self.pairs_blocked = build_pairs(self.subm_map(), self.n, 27, sum(self.rule_counts()), order=perm,
                                             row_block=pair_row_block(self.n, 27))
            self.pairs_ordered = self.pairs_blocked[:3]
        return self.pairs_blocked


class TilePlan:
    """perm (n,) int32 site ids along the Morton curve; lmap (T,27,128) uint16; halo_ids (T,hcap) int32; halo_n, kmask (T,)."""

    def __init__(self, level, hcap):
        nbr = level.subm_map()
        dev = nbr.device
        n = level.n
        st = _lib.stream_for(nbr)
        from . import ops as _ops
        tok = _ops._p0("morton_sort", "morton_perm (keys + radix sort)", 8.0 * n + 2 * 12.0 * n * 4 + 4.0 * n, 0, 0.0, 0.0)
        # site ids along the Morton curve: key build + radix sort inside the library (b200scn_morton_perm); the sample
        # index needs ceil(log2(batch)) bits, bounded here by the 15 bits the packed key reserves for it
        nbytes = lib.b200scn_morton_perm_scratch_bytes(_lib.round_rows(max(n, 1)))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.perm = alloc_flat(max(n, 1), dev, torch.int32)[:n]
        check(lib.b200scn_morton_perm(ptr(level.ukeys), n, level.size, level.batch_bits, ptr(self.perm), ptr(scratch), nbytes,
                                      st))
        _ops._p1(tok)
        T = (n + 127) // 128
        self.hcap = hcap
        self.lmap = alloc_flat(T * 27 * 128, dev, torch.int16)
        self.halo_ids = alloc_flat(T * hcap, dev, torch.int32)
        self.halo_n = alloc_flat(T, dev, torch.int32)
        self.kmask = alloc_flat(T, dev, torch.int32)
        tok = _ops._p0("tile_plan", "tile_plan", 4.0 * 27 * n + 4.0 * n + 2.0 * 27 * T * 128 + 4.0 * T * hcap, 0, 0.0, 0.0)
        check(lib.b200scn_tile_plan(ptr(nbr), ptr(self.perm), n, hcap, ptr(self.lmap), ptr(self.halo_ids),
                                    ptr(self.halo_n), ptr(self.kmask), st))
        _ops._p1(tok)


def pair_row_block(n, K):
    """Rows per block of a blocked pair table: a power of two in [1024, 8192] that leaves ~4 CTAs (offset x block) per SM."""
    want = max(1, (4 * 148 + K - 1) // K)
    rb = 8192
    while rb > 1024 and (n + rb - 1) // rb < want:
        rb //= 2
    return rb


def build_pairs(map_t, n, K, total, order=None, row_block=0):
    """scn-form rulebook (per-offset (in,out) pair lists, ascending out) from a map[n][K] with `total` entries >= 0.
    `order`: optional int32 permutation of the rows replacing "ascending".
    `row_block` > 0: also return the (K * nblk + 1,) table of row-block boundaries (b200scn_pair_lists_blocked)."""
    dev = map_t.device
    offsets = torch.empty(K + 1, dtype=torch.int32, device=dev)
    nbytes = lib.b200scn_pair_scratch_bytes(n, K)
    scratch = alloc_flat(nbytes, dev, torch.uint8)
    pair_in = alloc_flat(max(total, 1), dev, torch.int32)
    pair_out = alloc_flat(max(total, 1), dev, torch.int32)
    st = _lib.stream_for(map_t)
    from . import ops as _ops
    tok = _ops._p0("pair_lists", "pair_lists", 2.0 * 4.0 * K * n + 8.0 * total, 0, 0.0, 0.0)   # map read twice + pairs written
    if row_block:
        nblk = (max(n, 1) + row_block - 1) // row_block
        blk = torch.empty(K * nblk + 1, dtype=torch.int32, device=dev)
        check(lib.b200scn_pair_lists_blocked(ptr(map_t), ptr(order), n, K, row_block, ptr(pair_in), ptr(pair_out),
                                             ptr(offsets), ptr(blk), ptr(scratch), nbytes, st))
        _ops._p1(tok)
        return pair_in, pair_out, offsets, (blk, nblk)
    check(lib.b200scn_pair_lists_ordered(ptr(map_t), ptr(order), n, K, ptr(pair_in), ptr(pair_out), ptr(offsets),
                                         ptr(scratch), nbytes, st))
    _ops._p1(tok)
    return pair_in, pair_out, offsets


class Down:
    """Strided relation fine(size) -> coarse(size//s): parent/off per fine site, child map per coarse site."""

    def __init__(self, s, fine, coarse, parent, off):
        self.s, self.K, self.fine, self.coarse, self.parent, self.off = s, s ** 3, fine, coarse, parent, off
        self.child = None
        self.pairs = None
        self.gtab = None

    def child_map(self):
        if self.child is None:
            dev = self.parent.device
            self.child = alloc_rows(self.coarse.n, self.K, dev, torch.int32)
            st = _lib.stream_for(self.parent)
            check(lib.b200scn_child_map(ptr(self.parent), ptr(self.off), self.fine.n, None, self.K,
                                        ptr(self.child), self.coarse.n, st))
        return self.child

    def group_tiles(self):
        """Tile table of the offset-sorted rulebook for the fine-side tensor-core convolutions (b200scn_grouped_conv):
        (fine ids, coarse ids, table, max_tiles); built once per step and shared by Deconvolution forward and
        Convolution backward-input."""
        if self.gtab is None:
            pin, pout, offs = self.child_pairs()
            max_tiles = (self.fine.n + 127) // 128 + self.K
            tab = alloc_flat(4 * max_tiles, self.parent.device, torch.int32)
            check(lib.b200scn_group_tiles(ptr(offs), self.K, max_tiles, ptr(tab), _lib.stream_for(self.parent)))
            self.gtab = (pin, pout, tab, max_tiles)
        return self.gtab

    def child_pairs(self):
        """pair_in = fine ids, pair_out = coarse ids, grouped by offset."""
        if self.pairs is None:
            self.pairs = build_pairs(self.child_map(), self.coarse.n, self.K, self.fine.n)
        return self.pairs


class Metadata:
    def __init__(self, dimension=3):
        self.dimension = dimension
        self.levels = {}   # spatial size -> Level
        self.downs = {}    # (fine size, s) -> Down
        self.P = 0
        self.mode = 4
        self.pv = self.count = self.first_row = self.last_row = None
        self.syncs = 0     # host synchronisations this forward (diagnostic)
        self._site_rows = None

    # ------------------------------------------------------------------ InputLayer
    def build_input(self, coords, spatial_size, mode, device):
        """coords (P,3|4) int64 on any device -> level-0 grid (+ speculative strided pyramid)."""
        packed = coords if isinstance(coords, PackedKeys) else None
        if packed is None:
            if coords.dim() != 2 or coords.shape[1] not in (3, 4):
                raise ValueError("InputLayer: coords must be (N,3) or (N,4), got %s" % (tuple(coords.shape),))
            coords = coords.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
        P = coords.shape[0]
        Pc = _lib.round_rows(max(P, 1))   # quantised capacity of every per-point / per-site buffer
        self.P, self.mode = P, mode
        st = _lib.stream_for(coords if packed is None else packed.keys)
        i32, i64 = torch.int32, torch.int64
        keys = torch.empty(Pc, dtype=i64, device=device)
        s_hint, depth = _pyramid_hint
        sizes = [int(spatial_size)]
        while len(sizes) <= depth and sizes[-1] % s_hint == 0 and sizes[-1] // s_hint >= 1 and sizes[-1] > 1:
            sizes.append(sizes[-1] // s_hint)
        nlev = len(sizes)
        # header: [err, n_0, n_1, ..., n_{nlev-1}]
        hdr = torch.zeros(1 + nlev, dtype=i32, device=device)
        from . import ops as _ops
        # algorithmic bytes: 32 P coords read + 4 P ids written; the coarse levels re-hash <= P keys each
        tok = _ops._p0("grid_build", "pack_coords+hash+scan (all levels)", 36.0 * P if packed is None else 12.0 * P, 0, 0.0, 0.0)
        if packed is None:
            check(lib.b200scn_pack_coords(ptr(coords), P, coords.shape[1], int(spatial_size), ptr(keys), ptr(hdr), st))
        else:
            keys[:P].copy_(packed.keys)   # (the packed point keys are reused below as scratch of the coarse levels)
        cap = lib.b200scn_hash_capacity(Pc)
        sbytes = lib.b200scn_grid_scratch_bytes(Pc)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=device)
        self.pv = torch.empty(Pc, dtype=i32, device=device)
        self.first_row = torch.empty(Pc, dtype=i32, device=device)
        self.last_row = torch.empty(Pc, dtype=i32, device=device)
        self.count = torch.empty(Pc, dtype=i32, device=device)
        raw = []
        ukeys = torch.empty(Pc, dtype=i64, device=device)
        hkeys = torch.empty(cap, dtype=i64, device=device)
        hvals = torch.empty(cap, dtype=i32, device=device)
        check(lib.b200scn_grid_build(ptr(keys), P, None, ptr(hkeys), ptr(hvals), cap, ptr(self.pv), ptr(ukeys),
                                     ptr(self.first_row), ptr(self.last_row), ptr(self.count),
                                     hdr[1:].data_ptr(), ptr(scratch), sbytes, st))
        raw.append((ukeys, hkeys, hvals, None, None))
        for li in range(1, nlev):
            fu = raw[-1][0]
            ckeys = keys  # reuse: the packed point keys are dead after the level-0 build
            off = torch.empty(Pc, dtype=torch.uint8, device=device)
            parent = torch.empty(Pc, dtype=i32, device=device)
            n_dev = hdr[li:].data_ptr()
            check(lib.b200scn_coarse_keys(ptr(fu), P, n_dev, s_hint, ptr(ckeys), ptr(off), st))
            cu = torch.empty(Pc, dtype=i64, device=device)
            ck = torch.empty(cap, dtype=i64, device=device)
            cv = torch.empty(cap, dtype=i32, device=device)
            check(lib.b200scn_grid_build(ptr(ckeys), P, n_dev, ptr(ck), ptr(cv), cap, ptr(parent), ptr(cu),
                                         None, None, None, hdr[li + 1:].data_ptr(), ptr(scratch), sbytes, st))
            raw.append((cu, ck, cv, parent, off))
        _ops._p1(tok)
        host = hdr.cpu()  # the one host sync of the forward
        self.syncs += 1
        if int(host[0]) != 0:
            raise ValueError("InputLayer: coordinates outside [0, %d) or bad sample index" % int(spatial_size))
        counts = [int(v) for v in host[1:]]
        for li, size in enumerate(sizes):
            u, hk, hv, parent, off = raw[li]
            n = counts[li]
            lvl = Level(size, n, u[:n], hk, hv, cap)
            self.levels[size] = lvl
            if li > 0:
                fine = self.levels[sizes[li - 1]]
                self.downs[(sizes[li - 1], s_hint)] = Down(s_hint, fine, lvl, parent[:fine.n], off[:fine.n])
        n0 = counts[0]
        self.pv, self.count = self.pv[:P], self.count[:n0]
        self.first_row, self.last_row = self.first_row[:n0], self.last_row[:n0]
        return self.levels[sizes[0]]

    def site_rows(self):
        """(start, rows): input rows grouped by level-0 site, for the atomic-free OutputLayer backward."""
        if self._site_rows is None:
            n0 = self.count.shape[0]
            dev = self.pv.device
            start = alloc_flat(max(n0, 1), dev, torch.int32)
            rows = alloc_flat(max(self.P, 1), dev, torch.int32)
            nb = lib.b200scn_site_rows_scratch_bytes(n0)
            scratch = alloc_flat(nb, dev, torch.uint8)
            check(lib.b200scn_site_rows(ptr(self.pv), self.P, ptr(self.count), n0, ptr(start), ptr(rows), ptr(scratch),
                                        nb, _lib.stream_for(self.pv)))
            self._site_rows = (start, rows)
        return self._site_rows

    # ------------------------------------------------------------------ strided levels on demand
    def get_down(self, size, s):
        key = (size, s)
        if key not in self.downs:
            if size % s != 0:
                raise ValueError("Convolution: spatial size %d not divisible by stride %d" % (size, s))
            fine = self.levels[size]
            device = fine.ukeys.device
            st = _lib.stream_for(fine.ukeys)
            n = fine.n
            i32, i64 = torch.int32, torch.int64
            ckeys = torch.empty(max(n, 1), dtype=i64, device=device)
            off = torch.empty(max(n, 1), dtype=torch.uint8, device=device)
            parent = torch.empty(max(n, 1), dtype=i32, device=device)
            cu = torch.empty(max(n, 1), dtype=i64, device=device)
            cap = lib.b200scn_hash_capacity(n)
            ck = torch.empty(cap, dtype=i64, device=device)
            cv = torch.empty(cap, dtype=i32, device=device)
            cnt = torch.zeros(1, dtype=i32, device=device)
            sbytes = lib.b200scn_grid_scratch_bytes(n)
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=device)
            check(lib.b200scn_coarse_keys(ptr(fine.ukeys), n, None, s, ptr(ckeys), ptr(off), st))
            check(lib.b200scn_grid_build(ptr(ckeys), n, None, ptr(ck), ptr(cv), cap, ptr(parent), ptr(cu),
                                         None, None, None, ptr(cnt), ptr(scratch), sbytes, st))
            nc = int(cnt.cpu()[0])
            self.syncs += 1
            csize = (size - s) // s + 1
            coarse = self.levels.get(csize)
            if coarse is None or coarse.n != nc:
                # scn registers one grid per spatial size; a second, different grid of the same size can
                # only come from mixing stride chains, which the reference's nets never do
                if coarse is not None:
                    raise ValueError("two different grids for spatial size %d" % csize)
                coarse = Level(csize, nc, cu[:nc], ck, cv, cap)
                self.levels[csize] = coarse
            self.downs[key] = Down(s, fine, coarse, parent[:n], off[:n])
        return self.downs[key]

    def get_up(self, coarse_size, s):
        """The (fine, coarse) relation whose coarse side has `coarse_size` (Deconvolution / UnPooling reuse the
        rulebook built on the way down)."""
        fsize = (coarse_size - 1) * s + s
        key = (fsize, s)
        if key not in self.downs:
            raise ValueError("Deconvolution/UnPooling to spatial size %d: no matching Convolution was run" % fsize)
        return self.downs[key]
