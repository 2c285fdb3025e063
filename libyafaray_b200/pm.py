"""ctypes binding of the photon-map queries of libyafaray_b200/libb200rt.so (C ABI: include/b200pm.h).

Host-side mirror of the reference's PhotonMap for this path (include/photon/photon.h:57-92): `PhotonMap.gather` and
`PhotonMap.find_nearest` keep the reference's names, argument meaning and results, but take a batch of points per call.
The shared library is the product; this module only marshals pointers.  No CPU fallback: without the library or a CUDA device
the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import rt

NONE = 0xFFFFFFFF
FOUND_DTYPE = np.dtype([("photon", np.uint32), ("dist_square", np.float32)])
assert FOUND_DTYPE.itemsize == 8

#: every symbol include/b200pm.h declares (tests check that the library exports all of them)
SYMBOLS = ["b200pm_create", "b200pm_destroy", "b200pm_get_stats", "b200pm_gather", "b200pm_gather_device", "b200pm_find_nearest",
           "b200pm_find_nearest_device", "b200pm_host_tree_build", "b200pm_debug_set_tuning"]


class Stats(C.Structure):
    _fields_ = [("n_photons", C.c_uint64), ("n_nodes", C.c_uint64), ("depth", C.c_uint32), ("reserved_", C.c_uint32),
                ("build_seconds", C.c_double), ("upload_seconds", C.c_double), ("device_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved_"}


_ready = False


def lib():
    global _ready
    L = rt.lib()
    if not _ready:
        P, Z, U, F = C.c_void_p, C.c_size_t, C.c_uint32, C.c_float
        L.b200pm_create.argtypes = [C.c_int, P, P, Z, C.c_int, P]
        L.b200pm_destroy.argtypes = [P]
        L.b200pm_destroy.restype = None
        L.b200pm_get_stats.argtypes = [P, P]
        L.b200pm_gather.argtypes = [P, P, Z, U, F, P, P, P, P]
        L.b200pm_gather_device.argtypes = [P, P, Z, U, F, P, P, P, P, P]
        L.b200pm_find_nearest.argtypes = [P, P, P, Z, F, P]
        L.b200pm_find_nearest_device.argtypes = [P, P, P, Z, F, P, P]
        L.b200pm_host_tree_build.argtypes = [P, Z, C.c_int, P, P]
        L.b200pm_debug_set_tuning.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        _ready = True
    return L


def _f32(a, cols=3):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected an [n, {cols}] array, got {a.shape}")
    return a


def set_tuning(kernel: int = -1, round_steps: int = -1, smem_k: int = -1, patience: int = -1):
    """b200pm_debug_set_tuning: gather kernel (0 plain, 1 phased, 2 phased + single pop) / round length / shared-memory heap
    threshold / make_heap patience (tuning aid; results do not change)."""
    rt._check(lib().b200pm_debug_set_tuning(kernel, round_steps, smem_k, patience))


def host_tree(pos, build_threads: int = 0):
    """The host-side builder alone (no device): the reference's KdNode array as (a, b) (include/b200pm.h, diagnostics)."""
    pos = _f32(pos)
    n = len(pos)
    a, b = np.zeros(max(1, 2 * n - 1), np.uint32), np.zeros(max(1, 2 * n - 1), np.uint32)
    rt._check(lib().b200pm_host_tree_build(rt._p(pos), n, build_threads, rt._p(a), rt._p(b)))
    return a, b


def pack_nodes(a, b, pos):
    """The 16-byte device nodes (pm_kernels.cuh) from a KdNode array -- what b200pm_create uploads.  For the host model test."""
    pos = _f32(pos)
    nodes = np.zeros((len(a), 4), np.uint32)
    leaf = (b & 3) == 3
    nodes[:, 0] = a
    nodes[:, 3] = b
    photon = a[leaf]
    nodes[leaf, 0:3] = pos[photon].view(np.uint32)
    nodes[leaf, 3] = (photon << 2) | 3
    return nodes


class PhotonMap:
    """Owns one b200pm_map (PhotonMap::updateTree, src/photon/photon.cc:46-56)."""

    def __init__(self, pos, dirs=None, device: int = 0, build_threads: int = 0):
        pos = _f32(pos)
        dirs = None if dirs is None else _f32(dirs)
        if dirs is not None and len(dirs) != len(pos):
            raise ValueError("one direction per photon")
        self._h = C.c_void_p(0)
        rt._check(lib().b200pm_create(device, rt._p(pos), rt._p(dirs), len(pos), build_threads, C.byref(self._h)))
        self.n_photons = len(pos)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().b200pm_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def stats(self) -> dict:
        s = Stats()
        rt._check(lib().b200pm_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def gather(self, points, k: int, sq_radius: float = 0.0, sq_radii=None, out=None):
        """PhotonMap::gather (src/photon/photon.cc:58-64) for every row of `points`.

        Returns (found [n, k] FOUND_DTYPE, n_found [n] u32, sq_radius_out [n] f32); entries of `found` past n_found are
        photon = NONE, dist_square = 0.  out = the three result arrays to fill (e.g. page-locked ones from rt.PinnedBuffer); with
        `out` the entries past n_found are left as the library wrote them (unspecified)."""
        points = _f32(points)
        n = len(points)
        if out is not None:
            found, n_found, radius_out = out
            assert found.dtype == FOUND_DTYPE and found.shape == (n, k) and found.flags.c_contiguous
            assert n_found.dtype == np.uint32 and n_found.shape == (n,) and radius_out.dtype == np.float32 and radius_out.shape == (n,)
        else:
            found = np.zeros((n, k), FOUND_DTYPE)
            n_found = np.zeros(n, np.uint32)
            radius_out = np.zeros(n, np.float32)
        radii = None
        if sq_radii is not None:
            radii = np.ascontiguousarray(sq_radii, np.float32)
            if radii.shape != (n,):
                raise ValueError("one squared radius per point")
        rt._check(lib().b200pm_gather(self._h, rt._p(points), n, k, float(sq_radius), rt._p(radii), rt._p(found), rt._p(n_found), rt._p(radius_out)))
        if n and out is None:
            past = np.arange(k)[None, :] >= n_found[:, None]
            found["photon"][past] = NONE
            found["dist_square"][past] = 0.0
        return found, n_found, radius_out

    def find_nearest(self, points, normals, dist: float):
        """PhotonMap::findNearest (src/photon/photon.cc:66-72) for every row: photon index or NONE."""
        points, normals = _f32(points), _f32(normals)
        if len(points) != len(normals):
            raise ValueError("one normal per point")
        out = np.zeros(len(points), np.uint32)
        rt._check(lib().b200pm_find_nearest(self._h, rt._p(points), rt._p(normals), len(points), float(dist), rt._p(out)))
        return out

    # device-buffer variants: torch CUDA tensors (float32 [n,3] points; results allocated here) on the current stream
    def gather_device(self, d_points, k: int, sq_radius: float = 0.0, d_sq_radii=None, out=None):
        import torch

        n = int(d_points.shape[0])
        assert d_points.is_cuda and d_points.dtype == torch.float32 and d_points.is_contiguous()
        if out is None:
            out = (torch.empty((n, k, 2), dtype=torch.int32, device=d_points.device), torch.empty(n, dtype=torch.int32, device=d_points.device),
                   torch.empty(n, dtype=torch.float32, device=d_points.device))
        found, n_found, radius_out = out
        stream = torch.cuda.current_stream(d_points.device).cuda_stream
        rt._check(lib().b200pm_gather_device(self._h, d_points.data_ptr(), n, k, float(sq_radius), d_sq_radii.data_ptr() if d_sq_radii is not None else None,
                                             found.data_ptr(), n_found.data_ptr(), radius_out.data_ptr(), stream))
        return out

    def find_nearest_device(self, d_points, d_normals, dist: float, out=None):
        import torch

        n = int(d_points.shape[0])
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=d_points.device)
        stream = torch.cuda.current_stream(d_points.device).cuda_stream
        rt._check(lib().b200pm_find_nearest_device(self._h, d_points.data_ptr(), d_normals.data_ptr(), n, float(dist), out.data_ptr(), stream))
        return out
