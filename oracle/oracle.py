"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(rayjoin_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF_LSI = os.path.join(_HERE, "_ref", "libref_lsi.so")
NO_HIT = 0xFFFFFFFF


def build(ref=True):
    """Compile liboracle.so (and the reference-backed pins when the read-only
    reference tree is present).  Building the checker is not using it."""
    targets = [_LIB]
    subprocess.run(["make", "-s", "-C", _HERE, _LIB] + (["ref"] if ref else []),
                   check=True)
    return targets


class Scaling(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("rx", "ry", "rrx", "rry", "deltax", "deltay", "ddeltax", "ddeltay")] + \
               [(n, C.c_int64) for n in ("imin", "imax", "irange")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
            build(ref=False)
        _lib = C.CDLL(_LIB)
        _lib.orc_lsi_brute.restype = C.c_uint64
        _lib.orc_lsi_grid.restype = C.c_uint64
        _lib.orc_intersect_test.restype = C.c_int
        _lib.orc_intersect_point.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP threads of the oracle's drivers (bench.py: all the cores the process may use)."""
    lib().orc_set_num_threads(C.c_int(int(n)))


def scaling_init(min_x, min_y, max_x, max_y):
    s = Scaling()
    lib().orc_scaling_init(C.byref(s), C.c_double(min_x), C.c_double(min_y),
                           C.c_double(max_x), C.c_double(max_y))
    return s


def scale_points(s, xy, device_semantics=True):
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    out = np.empty(xy.shape, dtype=np.int64)
    f = lib().orc_scale_points_dev if device_semantics else lib().orc_scale_points_host
    f(C.byref(s), _p(xy), C.c_uint64(len(xy)), _p(out))
    return out


def unscale_points_host(s, xy):
    xy = _i64(xy).reshape(-1, 2)
    out = np.empty(xy.shape, dtype=np.float64)
    lib().orc_unscale_points_host(C.byref(s), _p(xy), C.c_uint64(len(xy)), _p(out))
    return out


def build_edges(row_index):
    row_index = _u32(row_index)
    n_chains = len(row_index) - 1
    n_edges = int(row_index[-1]) - n_chains if n_chains > 0 else 0
    p1 = np.empty(n_edges, dtype=np.uint32)
    ch = np.empty(n_edges, dtype=np.uint32)
    lib().orc_build_edges(_p(row_index), C.c_uint64(n_chains), _p(p1), _p(ch))
    return p1, ch


def intersect_batch(pts):
    """pts: (n, 8) int64 = e1p1x,e1p1y,e1p2x,e1p2y,e2p1x,... -> hit, x, y"""
    pts = _i64(pts).reshape(-1, 8)
    n = len(pts)
    hit = np.zeros(n, dtype=np.uint8)
    x = np.zeros(n, dtype=np.int64)
    y = np.zeros(n, dtype=np.int64)
    lib().orc_intersect_batch(_p(pts), C.c_uint64(n), _p(hit), _p(x), _p(y))
    return hit, x, y


def i128_to_double(words):
    """(n, 2) uint64 {lo, hi} words of signed 128-bit integers -> (double) value (host libgcc)."""
    words = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, 2)
    out = np.zeros(len(words), np.float64)
    lib().orc_i128_to_double(_p(words), C.c_uint64(len(words)), _p(out))
    return out


def lsi_refgrid(xy0, p1_0, xy1, p1_1, s, gsize, sort_map=1, brute=False, cap=None):
    """LSI with the semantics of the reference's GRID backend (src/app/lsi_grid.h:19-78):
    intersect_test(map-0 edge, map-1 edge), kept iff the cell of the intersection point is a
    cell both edges are registered in.  -> (eids of map sort_map, eids of the other map, x, y)
    sorted by (first, second)."""
    xy0, xy1, p1_0, p1_1 = _i64(xy0), _i64(xy1), _u32(p1_0), _u32(p1_1)
    cap = cap or max(1024, 2 * (len(p1_0) + len(p1_1)))
    a, b, x, y = _lsi_out(cap)
    lib().orc_lsi_refgrid.restype = C.c_uint64
    n = lib().orc_lsi_refgrid(_p(xy0), _p(p1_0), C.c_uint64(len(p1_0)), _p(xy1), _p(p1_1),
                              C.c_uint64(len(p1_1)), C.c_int64(s.imin), C.c_int64(s.irange),
                              C.c_int(gsize), C.c_int(sort_map), C.c_int(1 if brute else 0),
                              _p(a), _p(b), _p(x), _p(y), C.c_uint64(cap))
    if n > cap:
        return lsi_refgrid(xy0, p1_0, xy1, p1_1, s, gsize, sort_map, brute, cap=int(n))
    return a[:n], b[:n], x[:n], y[:n]


def refgrid_cells(pts, s, gsize):
    """pts: (n, 8) int64 {map-0 edge, map-1 edge} -> hit, cell x, cell y of the intersection
    point as the reference's grid computes it (src/grid/cell.h:15-22), and whether its grid
    backend reports the pair (src/app/lsi_grid.h:62-67)."""
    pts = _i64(pts).reshape(-1, 8)
    n = len(pts)
    hit, owned = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    cx, cy = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().orc_refgrid_cells(_p(pts), C.c_uint64(n), C.c_int(gsize), C.c_int64(s.imin), C.c_int64(s.irange),
                            _p(hit), _p(cx), _p(cy), _p(owned))
    return hit, cx, cy, owned


def ref_lsi_available():
    return os.path.exists(_REF_LSI)


_ref = None


def ref_intersect_batch(pts):
    """Same contract as intersect_batch, but evaluated by the REFERENCE's own
    src/algo/lsi.h compiled on the host (oracle/_ref/libref_lsi.so)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_LSI)
    pts = _i64(pts).reshape(-1, 8)
    n = len(pts)
    hit = np.zeros(n, dtype=np.uint8)
    hit2 = np.zeros(n, dtype=np.uint8)
    x = np.zeros(n, dtype=np.int64)
    y = np.zeros(n, dtype=np.int64)
    _ref.ref_intersect_batch(_p(pts), C.c_uint64(n), _p(hit), _p(x), _p(y), _p(hit2))
    assert np.array_equal(hit, hit2)
    return hit, x, y


def _lsi_out(cap):
    return (np.empty(cap, np.uint32), np.empty(cap, np.uint32),
            np.empty(cap, np.int64), np.empty(cap, np.int64))


def lsi_brute(xyq, q_p1, xyb, b_p1, cap=None):
    xyq, xyb, q_p1, b_p1 = _i64(xyq), _i64(xyb), _u32(q_p1), _u32(b_p1)
    cap = cap or max(1024, 4 * (len(q_p1) + len(b_p1)))
    eq, eb, x, y = _lsi_out(cap)
    n = lib().orc_lsi_brute(_p(xyq), _p(q_p1), C.c_uint64(len(q_p1)), _p(xyb), _p(b_p1),
                            C.c_uint64(len(b_p1)), _p(eq), _p(eb), _p(x), _p(y),
                            C.c_uint64(cap))
    if n > cap:
        return lsi_brute(xyq, q_p1, xyb, b_p1, cap=int(n))
    return eq[:n], eb[:n], x[:n], y[:n]


def lsi_grid(xyq, q_p1, xyb, b_p1, s, cap=None, return_candidates=False):
    xyq, xyb, q_p1, b_p1 = _i64(xyq), _i64(xyb), _u32(q_p1), _u32(b_p1)
    cap = cap or max(1024, 2 * (len(q_p1) + len(b_p1)))
    eq, eb, x, y = _lsi_out(cap)
    ncand = C.c_uint64(0)
    n = lib().orc_lsi_grid(_p(xyq), _p(q_p1), C.c_uint64(len(q_p1)), _p(xyb), _p(b_p1),
                           C.c_uint64(len(b_p1)), C.c_int64(s.imin), C.c_int64(s.irange),
                           _p(eq), _p(eb), _p(x), _p(y), C.c_uint64(cap), C.byref(ncand))
    if n > cap:
        return lsi_grid(xyq, q_p1, xyb, b_p1, s, cap=int(n),
                        return_candidates=return_candidates)
    res = (eq[:n], eb[:n], x[:n], y[:n])
    return res + (ncand.value,) if return_candidates else res


def pip_brute(xyb, b_p1, pts, query_map_id):
    xyb, b_p1, pts = _i64(xyb), _u32(b_p1), _i64(pts).reshape(-1, 2)
    out = np.empty(len(pts), np.uint32)
    lib().orc_pip_brute(_p(xyb), _p(b_p1), C.c_uint64(len(b_p1)), _p(pts),
                        C.c_uint64(len(pts)), C.c_int(query_map_id), _p(out))
    return out


def pip_grid(xyb, b_p1, s, pts, query_map_id, return_candidates=False):
    xyb, b_p1, pts = _i64(xyb), _u32(b_p1), _i64(pts).reshape(-1, 2)
    out = np.empty(len(pts), np.uint32)
    ncand = C.c_uint64(0)
    lib().orc_pip_grid(_p(xyb), _p(b_p1), C.c_uint64(len(b_p1)), C.c_int64(s.imin),
                       C.c_int64(s.irange), _p(pts), C.c_uint64(len(pts)),
                       C.c_int(query_map_id), _p(out), C.byref(ncand))
    return (out, ncand.value) if return_candidates else out


def face_ids(xyb, b_p1, b_chain, left, right, eids):
    xyb, b_p1, b_chain = _i64(xyb), _u32(b_p1), _u32(b_chain)
    left, right, eids = _i64(left), _i64(right), _u32(eids)
    out = np.empty(len(eids), np.int32)
    lib().orc_face_ids(_p(xyb), _p(b_p1), _p(b_chain), _p(left), _p(right), _p(eids),
                       C.c_uint64(len(eids)), _p(out))
    return out


def ref_scaling_apply(bbox, xy, ixy):
    """Reference Scaling<double> on the host (no FMA): ScaleX/Y of xy, UnscaleX/Y of ixy."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_LSI)
    bbox = np.ascontiguousarray(bbox, dtype=np.float64)
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    ixy = _i64(ixy).reshape(-1, 2)
    scaled = np.empty(xy.shape, np.int64)
    unscaled = np.empty(ixy.shape, np.float64)
    limits = np.zeros(3, np.int64)
    _ref.ref_scaling_apply(_p(bbox), _p(xy), C.c_uint64(len(xy)), _p(scaled), _p(ixy),
                           C.c_uint64(len(ixy)), _p(unscaled), _p(limits))
    return scaled, unscaled, limits
