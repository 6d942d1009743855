"""ctypes binding of librjb200.so (include/rjb200.h) and a thin host-side mirror
of RayJoin's operator interface (Context / LSI / PIP / MapOverlay).

There is no CPU fallback: if the CUDA library is missing or no device is
present every entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RJB_LIB", os.path.join(_HERE, "librjb200.so"))  # (RJB_LIB: tuning builds)

MODE_GRID, MODE_LBVH, MODE_BRUTE = 0, 1, 2
MODES = {"grid": MODE_GRID, "lbvh": MODE_LBVH, "brute": MODE_BRUTE}
NO_HIT = 0xFFFFFFFF
ERR_QUEUE_OVERFLOW = 3

# every symbol include/rjb200.h declares (tests check the library exports them)
EXPORTS = [
    "rjb_last_error", "rjb_version", "rjb_create", "rjb_destroy", "rjb_set_stream",
    "rjb_set_bounding_box", "rjb_get_scaling", "rjb_set_map", "rjb_map_info",
    "rjb_map_device_views", "rjb_build_index", "rjb_set_option", "rjb_lsi", "rjb_lsi_launch", "rjb_lsi_wait",
    "rjb_last_launches", "rjb_pip",
    "rjb_pip_host", "rjb_pip_host_scaled", "rjb_overlay_run", "rjb_overlay_finish", "rjb_overlay_finish_device", "rjb_overlay_results", "rjb_overlay_write",
    "rjb_debug_sort_pairs", "rjb_debug_sort_packed", "rjb_debug_intersect_batch", "rjb_debug_i128_batch", "rjb_debug_pip_batch",
    "rjb_last_kernel_ms", "rjb_last_stage_ms", "rjb_last_stats", "rjb_index_info", "rjb_copy_to_host", "rjb_sync",
    "rjb_graph_load", "rjb_graph_read_text", "rjb_graph_read_bin", "rjb_graph_write_bin",
    "rjb_graph_free",
]


class RjbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rjb error %d: %s" % (code, msg))
        self.code = code


class Xsect(C.Structure):
    _fields_ = [("x", C.c_int64), ("y", C.c_int64), ("eid", C.c_uint32 * 2),
                ("mid_point_polygon_id", C.c_int32), ("_pad", C.c_int32)]


XSECT_DTYPE = np.dtype([("x", "<i8"), ("y", "<i8"), ("eid", "<u4", (2,)),
                        ("mid_point_polygon_id", "<i4"), ("_pad", "<i4")])
assert XSECT_DTYPE.itemsize == C.sizeof(Xsect) == 32


class Scaling(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("rx", "ry", "rrx", "rry", "deltax", "deltay", "ddeltax", "ddeltay")] + \
               [(n, C.c_int64) for n in ("internal_min", "internal_max", "internal_range")]


class Graph(C.Structure):
    _fields_ = [("n_chains", C.c_uint64), ("n_points", C.c_uint64),
                ("chain_id", C.POINTER(C.c_int64)), ("first_point", C.POINTER(C.c_int64)),
                ("last_point", C.POINTER(C.c_int64)), ("left", C.POINTER(C.c_int64)),
                ("right", C.POINTER(C.c_int64)), ("row_index", C.POINTER(C.c_uint32)),
                ("xy", C.POINTER(C.c_double)), ("min_x", C.c_double), ("min_y", C.c_double),
                ("max_x", C.c_double), ("max_y", C.c_double), ("_owner", C.c_void_p)]


_lib = None


def load_library():
    """dlopen librjb200.so; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RjbError(-1, "librjb200.so is not built (run python -m rayjoin_b200.build); "
                               "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.rjb_last_error.restype = C.c_char_p
        _lib.rjb_version.restype = C.c_char_p
        _lib.rjb_destroy.restype = None
        _lib.rjb_graph_free.restype = None
    return _lib


def _check(rc):
    if rc != 0:
        raise RjbError(rc, load_library().rjb_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PlanarGraph:
    """Host-side planar graph (reference: PlanarGraph<double>, planar_graph.h:32-39)."""

    def __init__(self, xy, row_index, left, right, chain_id=None, first_point=None,
                 last_point=None, bbox=None):
        self.xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        self.row_index = np.ascontiguousarray(row_index, dtype=np.uint32)
        self.left = np.ascontiguousarray(left, dtype=np.int64)
        self.right = np.ascontiguousarray(right, dtype=np.int64)
        n = len(self.left)
        self.chain_id = np.arange(n, dtype=np.int64) if chain_id is None else \
            np.ascontiguousarray(chain_id, dtype=np.int64)
        self.first_point = np.zeros(n, np.int64) if first_point is None else \
            np.ascontiguousarray(first_point, dtype=np.int64)
        self.last_point = np.zeros(n, np.int64) if last_point is None else \
            np.ascontiguousarray(last_point, dtype=np.int64)
        if bbox is None and len(self.xy):
            bbox = (self.xy[:, 0].min(), self.xy[:, 1].min(), self.xy[:, 0].max(), self.xy[:, 1].max())
        self.bbox = bbox

    @property
    def n_points(self):
        return len(self.xy)

    @property
    def n_chains(self):
        return len(self.left)

    @property
    def n_edges(self):
        return self.n_points - self.n_chains


def _graph_to_py(g):
    nc, npnt = g.n_chains, g.n_points

    def arr(p, n, dt):
        if n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)
    out = PlanarGraph(arr(g.xy, 2 * npnt, np.float64).reshape(-1, 2),
                      arr(g.row_index, nc + 1 if npnt else 0, np.uint32),
                      arr(g.left, nc, np.int64), arr(g.right, nc, np.int64),
                      arr(g.chain_id, nc, np.int64), arr(g.first_point, nc, np.int64),
                      arr(g.last_point, nc, np.int64),
                      (g.min_x, g.min_y, g.max_x, g.max_y))
    return out


def load_from(path, serialize_prefix=""):
    """load_from (planar_graph.h:222-252): .bin cache under serialize_prefix, else text."""
    lib = load_library()
    g = Graph()
    _check(lib.rjb_graph_load(path.encode(), serialize_prefix.encode(), C.byref(g)))
    try:
        return _graph_to_py(g)
    finally:
        lib.rjb_graph_free(C.byref(g))


def read_pgraph(path):
    lib = load_library()
    g = Graph()
    _check(lib.rjb_graph_read_text(path.encode(), C.byref(g)))
    try:
        return _graph_to_py(g)
    finally:
        lib.rjb_graph_free(C.byref(g))


def deserialize_pgraph(path):
    lib = load_library()
    g = Graph()
    _check(lib.rjb_graph_read_bin(path.encode(), C.byref(g)))
    try:
        return _graph_to_py(g)
    finally:
        lib.rjb_graph_free(C.byref(g))


def serialize_pgraph(pg, path):
    lib = load_library()
    g = Graph()
    g.n_chains, g.n_points = pg.n_chains, pg.n_points
    keep = [pg.chain_id, pg.first_point, pg.last_point, pg.left, pg.right, pg.row_index, pg.xy]
    g.chain_id = pg.chain_id.ctypes.data_as(C.POINTER(C.c_int64))
    g.first_point = pg.first_point.ctypes.data_as(C.POINTER(C.c_int64))
    g.last_point = pg.last_point.ctypes.data_as(C.POINTER(C.c_int64))
    g.left = pg.left.ctypes.data_as(C.POINTER(C.c_int64))
    g.right = pg.right.ctypes.data_as(C.POINTER(C.c_int64))
    g.row_index = pg.row_index.ctypes.data_as(C.POINTER(C.c_uint32))
    g.xy = pg.xy.ctypes.data_as(C.POINTER(C.c_double))
    g.min_x, g.min_y, g.max_x, g.max_y = pg.bbox
    _check(lib.rjb_graph_write_bin(C.byref(g), path.encode()))
    del keep


class Context:
    """Context (reference src/context.h): owns the stream, both maps and the scaling."""

    def __init__(self, graphs=None, device=0, bbox=None, stream=None):
        self.lib = load_library()
        self._h = C.c_void_p()
        _check(self.lib.rjb_create(C.c_int(device), C.byref(self._h)))
        self.device = device
        self.graphs = [None, None]
        if stream is not None:
            self.set_stream(stream)
        if graphs is not None:
            graphs = list(graphs) + [None] * (2 - len(graphs))
            if bbox is None:
                boxes = [g.bbox for g in graphs if g is not None and g.n_points]
                bbox = (min(b[0] for b in boxes), min(b[1] for b in boxes),
                        max(b[2] for b in boxes), max(b[3] for b in boxes))
            self.set_bounding_box(*bbox)
            for im, g in enumerate(graphs):
                if g is not None:
                    self.set_map(im, g)

    def close(self):
        if self._h:
            self.lib.rjb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(self.lib.rjb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_bounding_box(self, min_x, min_y, max_x, max_y):
        _check(self.lib.rjb_set_bounding_box(self._h, C.c_double(min_x), C.c_double(min_y),
                                             C.c_double(max_x), C.c_double(max_y)))

    def get_scaling(self):
        s = Scaling()
        _check(self.lib.rjb_get_scaling(self._h, C.byref(s)))
        return s

    def set_option(self, name, value):
        _check(self.lib.rjb_set_option(self._h, name.encode(), C.c_int64(value)))

    def set_map(self, map_id, g):
        self.graphs[map_id] = g
        _check(self.lib.rjb_set_map(self._h, C.c_int(map_id), _ptr(g.xy), C.c_uint64(g.n_points),
                                    _ptr(g.row_index), _ptr(g.left), _ptr(g.right),
                                    C.c_uint64(g.n_chains)))

    def set_map_raw(self, map_id, xy_ptr, n_points, row_index_ptr, left_ptr, right_ptr, n_chains):
        """Pointers to (pinned) host buffers, e.g. torch tensors' data_ptr()."""
        _check(self.lib.rjb_set_map(self._h, C.c_int(map_id), C.c_void_p(xy_ptr),
                                    C.c_uint64(n_points), C.c_void_p(row_index_ptr),
                                    C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                    C.c_uint64(n_chains)))

    def map_info(self, map_id):
        out = (C.c_uint64 * 3)()
        _check(self.lib.rjb_map_info(self._h, C.c_int(map_id), out))
        return {"points": out[0], "edges": out[1], "chains": out[2]}

    def map_points(self, map_id):
        """Scaled (internal) coordinates of the map's vertices, copied to the host."""
        info = self.map_info(map_id)
        d = C.c_void_p()
        _check(self.lib.rjb_map_device_views(self._h, C.c_int(map_id), C.byref(d), None))
        out = np.empty((info["points"], 2), np.int64)
        self.copy_to_host(d.value, out)
        return out

    def map_points_device_ptr(self, map_id):
        d = C.c_void_p()
        _check(self.lib.rjb_map_device_views(self._h, C.c_int(map_id), C.byref(d), None))
        return d.value

    def build_index(self, map_id, mode, grid_size=2048):
        ms = C.c_double(0)
        _check(self.lib.rjb_build_index(self._h, C.c_int(map_id), C.c_int(MODES.get(mode, mode)),
                                        C.c_uint32(grid_size), C.byref(ms)))
        return ms.value

    def index_info(self, map_id, mode):
        out = (C.c_uint64 * 4)()
        _check(self.lib.rjb_index_info(self._h, C.c_int(map_id), C.c_int(MODES.get(mode, mode)), out))
        return {"units": out[0], "bytes": out[1], "param": out[2], "cell_directory_bytes": out[3]}

    def last_kernel_ms(self):
        out = (C.c_double * 2)()
        _check(self.lib.rjb_last_kernel_ms(self._h, out))
        return out[0], out[1]

    def last_stage_ms(self):
        """(times[4], layout): per-kernel device times of the last query (rjb_last_stage_ms)."""
        out = (C.c_double * 4)()
        layout = C.c_int(0)
        _check(self.lib.rjb_last_stage_ms(self._h, out, C.byref(layout)))
        return list(out), layout.value

    def debug_intersect_batch(self, pts, mode=1):
        """Device run of lsi_intersect + the intersection point on (n, 8) int64 cases."""
        pts = np.ascontiguousarray(pts, dtype=np.int64).reshape(-1, 8)
        n = len(pts)
        flags, x, y = np.zeros(n, np.uint8), np.zeros(n, np.int64), np.zeros(n, np.int64)
        _check(self.lib.rjb_debug_intersect_batch(self._h, _ptr(pts), C.c_uint64(n), C.c_int(mode),
                                                  _ptr(flags), _ptr(x), _ptr(y)))
        return flags, x, y

    def debug_i128_batch(self, v, d=None):
        """v, d: (n, 2) uint64 {lo, hi} words of signed 128-bit integers -> (double) v,
        (double) v / (double) d, (int64) of the quotient -- computed on the device."""
        v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 2)
        n = len(v)
        cvt = np.zeros(n, np.float64)
        if d is None:
            _check(self.lib.rjb_debug_i128_batch(self._h, _ptr(v), None, C.c_uint64(n), _ptr(cvt), None, None))
            return cvt
        d = np.ascontiguousarray(d, dtype=np.uint64).reshape(-1, 2)
        div, tr = np.zeros(n, np.float64), np.zeros(n, np.int64)
        _check(self.lib.rjb_debug_i128_batch(self._h, _ptr(v), _ptr(d), C.c_uint64(n), _ptr(cvt), _ptr(div),
                                             _ptr(tr)))
        return cvt, div, tr

    def debug_pip_batch(self, edges, pts, query_map_id):
        edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 4)
        pts = np.ascontiguousarray(pts, dtype=np.int64).reshape(-1, 2)
        out = np.zeros(len(pts), np.uint32)
        _check(self.lib.rjb_debug_pip_batch(self._h, _ptr(edges), C.c_uint64(len(edges)), _ptr(pts),
                                            C.c_uint64(len(pts)), C.c_int(query_map_id), _ptr(out)))
        return out

    def debug_sort_pairs(self, keys, vals, begin_bit=0, end_bit=64):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        _check(self.lib.rjb_debug_sort_pairs(self._h, _ptr(keys), _ptr(vals), C.c_uint64(len(keys)),
                                             C.c_int(begin_bit), C.c_int(end_bit)))
        return keys, vals

    def debug_sort_packed(self, words, begin_bit=0, end_bit=32):
        words = np.ascontiguousarray(words, dtype=np.uint64).copy()
        _check(self.lib.rjb_debug_sort_packed(self._h, _ptr(words), C.c_uint64(len(words)),
                                              C.c_int(begin_bit), C.c_int(end_bit)))
        return words

    def last_stats(self):
        out = (C.c_uint64 * 8)()
        _check(self.lib.rjb_last_stats(self._h, out))
        return list(out)

    def copy_to_host(self, d_ptr, out):
        if out.nbytes:
            _check(self.lib.rjb_copy_to_host(self._h, C.c_void_p(d_ptr), _ptr(out),
                                             C.c_uint64(out.nbytes)))
        return out

    def sync(self):
        _check(self.lib.rjb_sync(self._h))

    # -- raw queries (device-resident results) ---------------------------------
    def lsi_device(self, query_map_id, mode, xsect_factor):
        d = C.c_void_p()
        n = C.c_uint64(0)
        cand = C.c_uint64(0)
        rc = self.lib.rjb_lsi(self._h, C.c_int(query_map_id), C.c_int(MODES.get(mode, mode)),
                              C.c_double(xsect_factor), C.byref(d), C.byref(n), C.byref(cand))
        if rc != 0:
            err = RjbError(rc, self.lib.rjb_last_error().decode("utf-8", "replace"))
            err.needed = n.value
            raise err
        return d.value, n.value, cand.value

    def lsi_launch(self, query_map_id, mode, xsect_factor):
        _check(self.lib.rjb_lsi_launch(self._h, C.c_int(query_map_id), C.c_int(MODES.get(mode, mode)),
                                       C.c_double(xsect_factor)))

    def lsi_wait(self):
        d = C.c_void_p()
        n = C.c_uint64(0)
        cand = C.c_uint64(0)
        rc = self.lib.rjb_lsi_wait(self._h, C.byref(d), C.byref(n), C.byref(cand))
        if rc != 0:
            err = RjbError(rc, self.lib.rjb_last_error().decode("utf-8", "replace"))
            err.needed = n.value
            raise err
        return d.value, n.value, cand.value

    def last_launches(self):
        out = C.c_uint32(0)
        _check(self.lib.rjb_last_launches(self._h, C.byref(out)))
        return out.value

    def pip_host_scaled(self, query_map_id, mode, h_points_ptr, n_points, h_eid_ptr, h_face_ptr):
        """End-to-end PIP from (pinned) host buffers: scaled int64 points in, closest edge ids
        and face ids out (rjb_pip_host_scaled); raw pointers, e.g. torch tensors' data_ptr()."""
        _check(self.lib.rjb_pip_host_scaled(self._h, C.c_int(query_map_id), C.c_int(MODES.get(mode, mode)),
                                            C.c_void_p(h_points_ptr), C.c_uint64(n_points),
                                            C.c_void_p(h_eid_ptr), C.c_void_p(h_face_ptr)))

    def pip_device(self, query_map_id, mode, d_points_ptr=None, n_points=0):
        de, df = C.c_void_p(), C.c_void_p()
        cand = C.c_uint64(0)
        _check(self.lib.rjb_pip(self._h, C.c_int(query_map_id), C.c_int(MODES.get(mode, mode)),
                                C.c_void_p(d_points_ptr), C.c_uint64(n_points), C.byref(de),
                                C.byref(df), C.byref(cand)))
        return de.value, df.value, cand.value


class LSI:
    """LSI<CTX> (reference src/app/lsi.h:7-43): Init / Query / get_xsects."""

    def __init__(self, ctx, mode="lbvh"):
        self.ctx, self.mode = ctx, mode
        self.xsect_factor = 0.2  # FLAGS_xsect_factor default, src/flags.cc:7
        self._res = None
        self.n_candidates = 0

    def Init(self, xsect_factor):
        self.xsect_factor = xsect_factor

    def Query(self, query_map_id):
        d, n, cand = self.ctx.lsi_device(query_map_id, self.mode, self.xsect_factor)
        self._res = (d, n)
        self.n_candidates = cand
        return n

    def Launch(self, query_map_id):
        """First half of Query: enqueue only (rjb_lsi_launch)."""
        self.ctx.lsi_launch(query_map_id, self.mode, self.xsect_factor)

    def Wait(self):
        """Second half of Query: complete the launched query (rjb_lsi_wait)."""
        d, n, cand = self.ctx.lsi_wait()
        self._res = (d, n)
        self.n_candidates = cand
        return n

    def get_xsects(self):
        """Host copy of the result queue as a structured array (rjb_xsect)."""
        d, n = self._res
        out = np.empty(n, XSECT_DTYPE)
        return self.ctx.copy_to_host(d, out)


class PIP:
    """PIP<CTX> (reference src/app/pip.h:8-38): Query / get_closest_eids."""

    def __init__(self, ctx, mode="lbvh"):
        self.ctx, self.mode = ctx, mode
        self._res = None
        self.n_candidates = 0

    def Query(self, query_map_id, points_xy=None):
        """points_xy: host int64 (n, 2) scaled points, or None for all vertices of the
        query map (src/run_query.cu:346)."""
        if points_xy is None:
            n = self.ctx.map_info(query_map_id)["points"]
            de, df, cand = self.ctx.pip_device(query_map_id, self.mode)
        else:
            import torch  # device memory plumbing only
            pts = np.ascontiguousarray(points_xy, dtype=np.int64).reshape(-1, 2)
            n = len(pts)
            self._dev_pts = torch.from_numpy(pts).to("cuda:%d" % self.ctx.device)
            torch.cuda.synchronize()
            de, df, cand = self.ctx.pip_device(query_map_id, self.mode,
                                               self._dev_pts.data_ptr(), n)
        self._res = (de, df, n)
        self.n_candidates = cand
        return n

    def get_closest_eids(self):
        de, _, n = self._res
        return self.ctx.copy_to_host(de, np.empty(n, np.uint32))

    def get_face_ids(self):
        _, df, n = self._res
        return self.ctx.copy_to_host(df, np.empty(n, np.int32))


class MapOverlay:
    """MapOverlay<CTX> (reference src/app/map_overlay.h:9-56, driver order of
    src/run_overlay.cu:196-226)."""

    def __init__(self, ctx, mode="lbvh", grid_size=2048, xsect_factor=0.2):
        self.ctx, self.mode, self.grid_size, self.xsect_factor = ctx, mode, grid_size, xsect_factor
        self.phase_ms = None

    def Run(self):
        ms = (C.c_double * 6)()
        _check(self.ctx.lib.rjb_overlay_run(self.ctx._h, C.c_int(MODES.get(self.mode, self.mode)),
                                            C.c_uint32(self.grid_size),
                                            C.c_double(self.xsect_factor), ms))
        self.phase_ms = dict(zip(("build", "lsi", "pip0", "pip1", "polygons", "total"), ms))
        return self.phase_ms

    def Finish(self, xsects, closest_eids, point_in_polygon):
        """Final step of the multi-GPU overlay on one rank (rjb_overlay_finish): the
        gathered results of the sharded LSI / vertex-location phases go in,
        ComputeOutputPolygons runs here."""
        xs = np.ascontiguousarray(xsects, dtype=XSECT_DTYPE)
        ce = [np.ascontiguousarray(a, dtype=np.uint32) for a in closest_eids]
        pf = [np.ascontiguousarray(a, dtype=np.int32) for a in point_in_polygon]
        ms = (C.c_double * 6)()
        _check(self.ctx.lib.rjb_overlay_finish(
            self.ctx._h, C.c_int(MODES.get(self.mode, self.mode)), C.c_uint32(self.grid_size),
            _ptr(xs), C.c_uint64(len(xs)), _ptr(ce[0]), _ptr(pf[0]), _ptr(ce[1]), _ptr(pf[1]), ms))
        self.phase_ms = dict(zip(("build", "lsi", "pip0", "pip1", "polygons", "total"), ms))
        return self.phase_ms

    def FinishDevice(self, d_xsects, n_xsects, d_closest_eids, d_point_in_polygon):
        """Finish with the gathered arrays in device memory (raw pointers, e.g. torch data_ptr())."""
        ms = (C.c_double * 6)()
        _check(self.ctx.lib.rjb_overlay_finish_device(
            self.ctx._h, C.c_int(MODES.get(self.mode, self.mode)), C.c_uint32(self.grid_size),
            C.c_void_p(d_xsects), C.c_uint64(n_xsects), C.c_void_p(d_closest_eids[0]),
            C.c_void_p(d_point_in_polygon[0]), C.c_void_p(d_closest_eids[1]),
            C.c_void_p(d_point_in_polygon[1]), ms))
        self.phase_ms = dict(zip(("build", "lsi", "pip0", "pip1", "polygons", "total"), ms))
        return self.phase_ms

    def _results(self, im):
        dx, de, dp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_uint64(0)
        _check(self.ctx.lib.rjb_overlay_results(self.ctx._h, C.c_int(im), C.byref(dx), C.byref(n),
                                                C.byref(de), C.byref(dp)))
        return dx.value, n.value, de.value, dp.value

    def get_xsect_edges(self, im=0):
        dx, n, _, _ = self._results(im)
        return self.ctx.copy_to_host(dx, np.empty(n, XSECT_DTYPE))

    def get_closet_eids(self, im):
        _, _, de, _ = self._results(im)
        return self.ctx.copy_to_host(de, np.empty(self.ctx.map_info(im)["points"], np.uint32))

    def get_point_in_polygon(self, im):
        _, _, _, dp = self._results(im)
        return self.ctx.copy_to_host(dp, np.empty(self.ctx.map_info(im)["points"], np.int32))

    def WriteResult(self, path):
        _check(self.ctx.lib.rjb_overlay_write(self.ctx._h, path.encode()))
