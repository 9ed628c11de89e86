"""point2mask_ext -- B200 replacement for the reference's torch C++/CUDA extension of the same name
(ops/point2mask/_ext_src/src/bindings.cpp:4-9: ball_query, group_points, group_points_grad), over the b200scn C ABI.
Same argument order, dtypes and checks (utils.h:5-25: CUDA + contiguous + float32 / int32); results identical,
including the -1 sentinel and the `n - ptnum` scan bound of ball_query_gpu.cu:28.  No CPU path ("CPU not supported",
ball_query.cpp:28-30)."""
import torch

from sparseconvnet import _lib
from sparseconvnet._lib import check, lib, ptr


def _chk(t, name, dtype):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be a %s tensor" % (name, "float" if dtype == torch.float32 else "int"))


def ball_query(new_coords, coords, pointsnum, radius, nsample, bucketed=None):
    _chk(new_coords, "new_coords", torch.float32)
    _chk(coords, "coords", torch.float32)
    _chk(pointsnum, "pointsnum", torch.int32)
    b, n, _ = coords.shape
    m = new_coords.shape[1]
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=coords.device)
    st = _lib.stream_for(coords)
    # Large inputs (the pseudo-dataset generator: 65 536 queries x ~200 k points per instance) go through the cell-bucketed
    # kernel -- identical results; the grid size comes from the queries' extent (one small read-back, this is an offline tool
    # path).  Small inputs keep the shared-memory scan.
    if bucketed is None:
        bucketed = (m * n >= (1 << 24))
    if bucketed and radius > 0 and m > 0 and n > 0:
        q = new_coords.reshape(-1, 2)
        ext = float((q.amax(0) - q.amin(0)).max())
        side = int(ext / float(radius)) + 5
        if b * side * side < (1 << 27):
            nb = lib.b200scn_p2m_ball_query_scratch_bytes(b, n, m, side)
            scratch = torch.empty(nb, dtype=torch.uint8, device=coords.device)
            check(lib.b200scn_p2m_ball_query_bucketed(b, n, m, float(radius), int(nsample), ptr(new_coords), ptr(coords),
                                                      ptr(pointsnum), ptr(idx), side, ptr(scratch), nb, st))
            return idx
    check(lib.b200scn_p2m_ball_query(b, n, m, float(radius), int(nsample), ptr(new_coords), ptr(coords), ptr(pointsnum),
                                     ptr(idx), st))
    return idx


def group_points(points, idx):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((b, c, npoints, nsample), dtype=torch.float32, device=points.device)
    check(lib.b200scn_p2m_group_points(b, c, n, npoints, nsample, ptr(points), ptr(idx), ptr(out), _lib.stream_for(points)))
    return out


def group_points_grad(grad_out, idx, n):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    b, c, npoints, nsample = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    check(lib.b200scn_p2m_group_points_grad(b, c, int(n), npoints, nsample, ptr(grad_out), ptr(idx), ptr(out),
                                            _lib.stream_for(grad_out)))
    return out
