"""ctypes binding of libyafaray_b200/libb200rt.so (C ABI: include/b200rt.h).

The shared library is the product; this module only marshals pointers.  There is no CPU fallback: if the
library has not been built, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RT_LIB") or os.path.join(_HERE, "libb200rt.so")  # B200RT_LIB: tuning builds (tools/sweep.sh)
MISS = 0xFFFFFFFF
QUERY_CLOSEST, QUERY_SHADOW, QUERY_TSHADOW = 0, 1, 2
RAYS_TREE_SPACE = 1
BUFFERS_PINNED = 2
TSHADOW_MAX = 8

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("tmin", np.float32), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32)])
TSHADOW_DTYPE = np.dtype([("shadowed", np.uint32), ("n_transparent", np.uint32), ("occluder", np.uint32), ("pad_", np.uint32),
                          ("transparent", HIT_DTYPE, TSHADOW_MAX)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16 and TSHADOW_DTYPE.itemsize == 144


class BuildParams(C.Structure):
    _fields_ = [("max_depth", C.c_int), ("max_leaf_size", C.c_int), ("cost_ratio", C.c_float), ("empty_bonus", C.c_float),
                ("build_threads", C.c_int), ("reserved_", C.c_int * 3)]


class Stats(C.Structure):
    _fields_ = [("n_faces", C.c_uint64), ("n_triangles", C.c_uint64), ("n_quads", C.c_uint64),
                ("n_nodes", C.c_uint64), ("n_interior", C.c_uint64), ("n_leaves", C.c_uint64), ("n_empty_leaves", C.c_uint64),
                ("n_leaf_refs", C.c_uint64), ("max_depth", C.c_uint32), ("max_leaf_prims", C.c_uint32),
                ("build_seconds", C.c_double), ("upload_seconds", C.c_double), ("device_bytes", C.c_uint64),
                ("n_spheres", C.c_uint64), ("n_bezier_faces", C.c_uint64), ("n_moving_faces", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Job(C.Structure):
    """b200rt_job (include/b200rt.h)."""
    _fields_ = [("scene", C.c_void_p), ("query", C.c_int), ("flags", C.c_uint), ("rays", C.c_void_p), ("n", C.c_size_t), ("out", C.c_void_p), ("max_depth", C.c_int),
                ("times", C.c_void_p)]


class B200RTError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libb200rt error {code}: {text}")
        self.code = code


#: every symbol include/b200rt.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "b200rt_device_count", "b200rt_create", "b200rt_destroy", "b200rt_add_mesh", "b200rt_add_spheres", "b200rt_build", "b200rt_get_bound",
    "b200rt_get_stats", "b200rt_update_face_flags", "b200rt_trace_closest", "b200rt_trace_shadow", "b200rt_trace_tshadow",
    "b200rt_trace_closest_device", "b200rt_trace_shadow_device", "b200rt_trace_tshadow_device", "b200rt_trace",
    "b200rt_trace_device", "b200rt_trace_timed", "b200rt_trace_timed_device", "b200rt_add_mesh_bezier", "b200rt_add_mesh_moving", "b200rt_trace_tshadow_deep", "b200rt_trace_tshadow_deep_device", "b200rt_trace_jobs", "b200rt_trace_jobs_begin", "b200rt_trace_jobs_end", "b200rt_host_alloc",
    "b200rt_host_free", "b200rt_host_tree_build", "b200rt_host_tree_sizes", "b200rt_host_tree_export",
    "b200rt_host_tree_destroy", "b200rt_launch_count", "b200rt_last_error", "b200rt_version",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C libyafaray_b200/csrc); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        P, Z = C.c_void_p, C.c_size_t
        L.b200rt_device_count.argtypes = [P]
        L.b200rt_create.argtypes = [C.c_int, P, P]
        L.b200rt_destroy.argtypes = [P]
        L.b200rt_destroy.restype = None
        L.b200rt_add_mesh.argtypes = [P, P, Z, P, Z, P]
        L.b200rt_add_spheres.argtypes = [P, P, Z, P]
        L.b200rt_build.argtypes = [P]
        L.b200rt_get_bound.argtypes = [P, P]
        L.b200rt_get_stats.argtypes = [P, P]
        L.b200rt_update_face_flags.argtypes = [P, P, Z]
        L.b200rt_trace_closest.argtypes = [P, P, Z, P]
        L.b200rt_trace_shadow.argtypes = [P, P, Z, P]
        L.b200rt_trace_tshadow.argtypes = [P, P, Z, C.c_int, P]
        L.b200rt_trace_closest_device.argtypes = [P, P, Z, P, P]
        L.b200rt_trace_shadow_device.argtypes = [P, P, Z, P, P]
        L.b200rt_trace_tshadow_device.argtypes = [P, P, Z, C.c_int, P, P]
        L.b200rt_trace.argtypes = [P, C.c_int, C.c_uint, P, Z, P, C.c_int]
        L.b200rt_trace_device.argtypes = [P, C.c_int, C.c_uint, P, Z, P, C.c_int, P]
        L.b200rt_trace_timed.argtypes = [P, C.c_int, C.c_uint, P, P, Z, P, C.c_int]
        L.b200rt_trace_timed_device.argtypes = [P, C.c_int, C.c_uint, P, P, Z, P, C.c_int, P]
        L.b200rt_add_mesh_bezier.argtypes = [P, P, P, P, Z, P, Z, P, C.c_float, C.c_float]
        L.b200rt_add_mesh_moving.argtypes = [P, P, Z, P, Z, P, P, C.c_float, C.c_float]
        L.b200rt_trace_tshadow_deep.argtypes = [P, C.c_uint, P, P, Z, C.c_int, C.c_int, P]
        L.b200rt_trace_tshadow_deep_device.argtypes = [P, C.c_uint, P, P, Z, C.c_int, C.c_int, P, P]
        L.b200rt_trace_jobs.argtypes = [P, Z]
        L.b200rt_trace_jobs_begin.argtypes = [P, Z, P]
        L.b200rt_trace_jobs_end.argtypes = [P]
        L.b200rt_host_alloc.argtypes = [P, Z]
        L.b200rt_host_free.argtypes = [P]
        L.b200rt_host_tree_build.argtypes = [P, Z, P, Z, P, P]
        L.b200rt_host_tree_sizes.argtypes = [P, P, P]
        L.b200rt_host_tree_export.argtypes = [P, P, P, P, P]
        L.b200rt_host_tree_destroy.argtypes = [P]
        L.b200rt_host_tree_destroy.restype = None
        L.b200rt_launch_count.restype = C.c_uint64
        L.b200rt_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise B200RTError(rc, lib().b200rt_last_error().decode(errors="replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().b200rt_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def launch_count() -> int:
    return int(lib().b200rt_launch_count())


def make_params(depth=0, max_leaf_size=0, cost_ratio=0.0, empty_bonus=0.0, build_threads=0) -> BuildParams:
    p = BuildParams()
    p.max_depth, p.max_leaf_size, p.cost_ratio, p.empty_bonus, p.build_threads = int(depth), int(max_leaf_size), float(cost_ratio), float(empty_bonus), int(build_threads)
    return p


def as_rays(rays) -> np.ndarray:
    """Accept float32 [n, 8] (ox oy oz tmin dx dy dz tmax) or a RAY_DTYPE array; return contiguous float32 [n, 8]."""
    r = np.asarray(rays)
    if r.dtype == RAY_DTYPE:
        r = r.view(np.float32).reshape(-1, 8)
    r = np.ascontiguousarray(r, dtype=np.float32)
    if r.ndim != 2 or r.shape[1] != 8:
        raise ValueError("rays must have shape [n, 8]")
    return r


class PinnedBuffer:
    """Page-locked host memory from b200rt_host_alloc, viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = C.c_void_p(0)
        _check(lib().b200rt_host_alloc(C.byref(self._ptr), nbytes))
        buf = (C.c_char * max(1, nbytes)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr and self._ptr.value:
            self.array = None
            lib().b200rt_host_free(self._ptr)
            self._ptr = C.c_void_p(0)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Scene:
    """Owns one b200rt_scene: add meshes, build, trace."""

    def __init__(self, device: int = 0, params: BuildParams | None = None):
        self._h = C.c_void_p(0)
        _check(lib().b200rt_create(int(device), C.byref(params) if params is not None else None, C.byref(self._h)))
        self.device = device
        self.n_faces = 0

    def close(self):
        if self._h and self._h.value:
            lib().b200rt_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_mesh(self, xyz, idx, flags=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
            if flags.shape[0] != idx.shape[0]:
                raise ValueError("one flag byte per face")
        _check(lib().b200rt_add_mesh(self._h, _p(xyz), xyz.shape[0], _p(idx), idx.shape[0], _p(flags)))
        self.n_faces += idx.shape[0]

    def add_spheres(self, center_radius, flags=None):
        """Spheres as rows cx, cy, cz, radius (b200rt_add_spheres); they take the next face ids."""
        cr = np.ascontiguousarray(center_radius, dtype=np.float32).reshape(-1, 4)
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
            if flags.shape[0] != cr.shape[0]:
                raise ValueError("one flag byte per sphere")
        _check(lib().b200rt_add_spheres(self._h, _p(cr), cr.shape[0], _p(flags)))
        self.n_faces += cr.shape[0]

    def add_mesh_bezier(self, xyz0, xyz1, xyz2, idx, flags=None, time_range=(0.0, 1.0)):
        """Faces of a Bezier motion-blur mesh: the vertex arrays of its three time steps as the mesh stores them (b200rt_add_mesh_bezier)."""
        steps = [np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3) for x in (xyz0, xyz1, xyz2)]
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
        flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        _check(lib().b200rt_add_mesh_bezier(self._h, _p(steps[0]), _p(steps[1]), _p(steps[2]), steps[0].shape[0], _p(idx), idx.shape[0], _p(flags),
                                            float(time_range[0]), float(time_range[1])))
        self.n_faces += idx.shape[0]

    def add_mesh_moving(self, xyz, idx, matrices, flags=None, time_range=(0.0, 1.0)):
        """Faces of a moving instance: base vertices + three row-major 4x4 obj_to_world matrices (b200rt_add_mesh_moving)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(48)
        flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        _check(lib().b200rt_add_mesh_moving(self._h, _p(xyz), xyz.shape[0], _p(idx), idx.shape[0], _p(flags), _p(m), float(time_range[0]), float(time_range[1])))
        self.n_faces += idx.shape[0]

    def build(self):
        _check(lib().b200rt_build(self._h))

    def bound(self) -> np.ndarray:
        b = np.zeros(6, np.float32)
        _check(lib().b200rt_get_bound(self._h, _p(b)))
        return b

    def stats(self) -> dict:
        s = Stats()
        _check(lib().b200rt_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def update_face_flags(self, flags):
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        _check(lib().b200rt_update_face_flags(self._h, _p(flags), flags.shape[0]))

    # ---- host-buffer queries (numpy in, numpy out) ----
    def trace_closest(self, rays, out=None) -> np.ndarray:
        r = as_rays(rays)
        if out is None:
            out = np.empty(r.shape[0], HIT_DTYPE)
        _check(lib().b200rt_trace_closest(self._h, _p(r), r.shape[0], _p(out)))
        return out

    def trace_shadow(self, rays, out=None) -> np.ndarray:
        r = as_rays(rays)
        if out is None:
            out = np.empty(r.shape[0], np.uint32)
        _check(lib().b200rt_trace_shadow(self._h, _p(r), r.shape[0], _p(out)))
        return out

    def trace_tshadow(self, rays, max_depth, out=None) -> np.ndarray:
        r = as_rays(rays)
        if out is None:
            out = np.empty(r.shape[0], TSHADOW_DTYPE)
        _check(lib().b200rt_trace_tshadow(self._h, _p(r), r.shape[0], int(max_depth), _p(out)))
        return out

    def trace(self, query, rays, flags=0, max_depth=0, out=None, times=None) -> np.ndarray:
        """Generic entry point: query = QUERY_CLOSEST / QUERY_SHADOW / QUERY_TSHADOW, flags = RAYS_TREE_SPACE or 0; times = one ray
        time per ray (b200rt_trace_timed) or None."""
        r = as_rays(rays)
        if out is None:
            out = np.empty(r.shape[0], {QUERY_CLOSEST: HIT_DTYPE, QUERY_SHADOW: np.uint32, QUERY_TSHADOW: TSHADOW_DTYPE}.get(int(query), TSHADOW_DTYPE))
        if times is not None:
            times = np.ascontiguousarray(times, dtype=np.float32)
            if times.shape[0] != r.shape[0]:
                raise ValueError("one time per ray")
            _check(lib().b200rt_trace_timed(self._h, int(query), int(flags), _p(r), _p(times), r.shape[0], _p(out), int(max_depth)))
        else:
            _check(lib().b200rt_trace(self._h, int(query), int(flags), _p(r), r.shape[0], _p(out), int(max_depth)))
        return out

    def trace_tshadow_deep(self, rays, max_depth, capacity=None, flags=0, times=None) -> np.ndarray:
        """Transparent shadows with more than TSHADOW_MAX distinct transparent casters (b200rt_trace_tshadow_deep): a structured array
        with the fields of TSHADOW_DTYPE whose `transparent` list has `capacity` entries."""
        r = as_rays(rays)
        capacity = int(max_depth if capacity is None else capacity)
        dt = np.dtype([("shadowed", np.uint32), ("n_transparent", np.uint32), ("occluder", np.uint32), ("pad_", np.uint32), ("transparent", HIT_DTYPE, capacity)])
        out = np.zeros(r.shape[0], dt)
        if times is not None:
            times = np.ascontiguousarray(times, dtype=np.float32)
        _check(lib().b200rt_trace_tshadow_deep(self._h, int(flags), _p(r), _p(times), r.shape[0], int(max_depth), capacity, _p(out)))
        return out

    # ---- device-buffer queries (raw device pointers, e.g. torch.Tensor.data_ptr()) ----
    def trace_closest_device(self, d_rays: int, n: int, d_out: int, stream: int = 0):
        _check(lib().b200rt_trace_closest_device(self._h, C.c_void_p(d_rays), n, C.c_void_p(d_out), C.c_void_p(stream)))

    def trace_shadow_device(self, d_rays: int, n: int, d_out: int, stream: int = 0):
        _check(lib().b200rt_trace_shadow_device(self._h, C.c_void_p(d_rays), n, C.c_void_p(d_out), C.c_void_p(stream)))

    def trace_tshadow_device(self, d_rays: int, n: int, max_depth: int, d_out: int, stream: int = 0):
        _check(lib().b200rt_trace_tshadow_device(self._h, C.c_void_p(d_rays), n, int(max_depth), C.c_void_p(d_out), C.c_void_p(stream)))


def trace_jobs(jobs, split=False):
    """b200rt_trace_jobs over a list of (scene, query, flags, rays[n,8] float32, out array, max_depth); split=True goes through
    the _begin / _end pair.  The arrays must stay alive (and, for the in-place path, be PinnedBuffer arrays)."""
    arr = (Job * len(jobs))()
    for k, job in enumerate(jobs):
        scene, query, flags, rays, out, max_depth = job[:6]
        times = job[6] if len(job) > 6 else None
        arr[k] = Job(scene._h, int(query), int(flags), rays.ctypes.data, rays.shape[0], out.ctypes.data, int(max_depth), times.ctypes.data if times is not None else None)
    if not split:
        _check(lib().b200rt_trace_jobs(arr, len(jobs)))
        return
    flight = C.c_void_p(0)
    rc = lib().b200rt_trace_jobs_begin(arr, len(jobs), C.byref(flight))
    if flight.value:
        rc2 = lib().b200rt_trace_jobs_end(flight)
        rc = rc or rc2
    _check(rc)


def host_tree(xyz, idx, params: BuildParams | None = None) -> dict:
    """Run only the host-side builder (no GPU needed) and return the tree in the export format of b200rt.h."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
    h = C.c_void_p(0)
    L = lib()
    _check(L.b200rt_host_tree_build(_p(xyz), xyz.shape[0], _p(idx), idx.shape[0], C.byref(params) if params is not None else None, C.byref(h)))
    try:
        nn, nr = C.c_size_t(0), C.c_size_t(0)
        _check(L.b200rt_host_tree_sizes(h, C.byref(nn), C.byref(nr)))
        a = np.zeros(nn.value, np.uint32); b = np.zeros(nn.value, np.uint32)
        refs = np.zeros(max(1, nr.value), np.uint32); bound = np.zeros(6, np.float32)
        _check(L.b200rt_host_tree_export(h, _p(a), _p(b), _p(refs), _p(bound)))
    finally:
        L.b200rt_host_tree_destroy(h)
    return dict(a=a, b=b, refs=refs[: nr.value], bound=bound)
