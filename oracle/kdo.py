"""ctypes binding of oracle/libkdoracle.so -- the plain-C restatement (kd_oracle.c) of the reference's ray queries.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under libyafaray_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkdoracle.so")


class _Motion(C.Structure):
    _fields_ = [("kind", C.c_void_p), ("xyz1", C.c_void_p), ("xyz2", C.c_void_p), ("face_times", C.c_void_p), ("face_matrix", C.c_void_p),
                ("matrices", C.c_void_p)]


class _Mesh(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("n_verts", C.c_size_t), ("idx", C.c_void_p), ("n_faces", C.c_size_t), ("flags", C.c_void_p),
                ("motion", C.c_void_p)]


class _Tree(C.Structure):
    _fields_ = [("split", C.c_void_p), ("flags", C.c_void_p), ("first_ref", C.c_void_p), ("refs", C.c_void_p),
                ("n_nodes", C.c_size_t), ("bound", C.c_float * 6)]


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("interior", C.c_uint64), ("leaves", C.c_uint64), ("refs", C.c_uint64), ("tests", C.c_uint64)]

    def per_ray(self):
        n = max(1, self.rays)
        return dict(interior=self.interior / n, leaves=self.leaves / n, refs=self.refs / n, tests=self.tests / n)

    def bytes_per_ray(self, io_bytes):
        """IO + 8*(interior+leaves) + 4*refs + 36*tests."""
        p = self.per_ray()
        return io_bytes + 8.0 * (p["interior"] + p["leaves"]) + 4.0 * p["refs"] + 36.0 * p["tests"]


def build_library(force: bool = False) -> str:
    """Compile kd_oracle.c with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "kd_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_library()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.kdo_tree_bound.argtypes = [P, P]
        L.kdo_build.restype = P
        L.kdo_build.argtypes = [P, C.c_int, C.c_int]
        L.kdo_built_view.argtypes = [P, P]
        L.kdo_built_num_refs.restype = C.c_size_t
        L.kdo_built_num_refs.argtypes = [P]
        L.kdo_built_free.argtypes = [P]
        L.kdo_trace_closest.argtypes = [P, P, P, C.c_size_t, P, P, P, P, C.c_int, P]
        L.kdo_trace_shadow.argtypes = [P, P, P, C.c_size_t, P, P, C.c_int, P]
        L.kdo_trace_tshadow.argtypes = [P, P, P, C.c_size_t, C.c_int, P, P, P, C.c_int, C.c_int]
        L.kdo_brute_closest.argtypes = [P, P, P, C.c_size_t, P, P, P, P, C.c_int]
        L.kdo_trace_closest_timed.argtypes = [P, P, P, P, C.c_size_t, P, P, P, P, C.c_int]
        L.kdo_trace_shadow_timed.argtypes = [P, P, P, P, C.c_size_t, P, P, C.c_int]
        L.kdo_trace_tshadow_timed.argtypes = [P, P, P, P, C.c_size_t, C.c_int, P, P, P, C.c_int, C.c_int]
        L.kdo_brute_closest_timed.argtypes = [P, P, P, P, C.c_size_t, P, P, P, P, C.c_int]
        L.kdo_poly_intersect.restype = C.c_float
        L.kdo_poly_intersect.argtypes = [P, P, P, P, C.c_int, P, P, P, P]
        L.kdo_bound_cross.restype = C.c_int
        L.kdo_bound_cross.argtypes = [P, P, P, C.c_float, P, P]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """The restated queries over one mesh, on either its own small SAH tree or a tree exported from the
    unmodified reference (oracle.yref.RefScene.export_tree()), which reproduces the reference's ties too."""

    def __init__(self, xyz, idx, flags=None, *, tree=None, bound=None, max_leaf=2, max_depth=0, motion=None):
        """motion (optional): dict(kind u8[n_faces] 0 static / 1 Bezier face / 2 face of a moving instance, xyz1, xyz2 f32[n_verts, 3],
        face_times f32[n_faces, 2], face_matrix u32[n_faces], matrices f32[n_instances, 3, 16]) -- libyafaray_b200/scenes.py::motion_scene."""
        L = lib()
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.idx = np.ascontiguousarray(idx, dtype=np.uint32)
        n_faces = self.idx.shape[0]
        self.flags = np.full(n_faces, 3, np.uint8) if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        self._motion = None
        if motion is not None:
            self._mo = dict(kind=np.ascontiguousarray(motion["kind"], np.uint8), xyz1=np.ascontiguousarray(motion["xyz1"], np.float32),
                            xyz2=np.ascontiguousarray(motion["xyz2"], np.float32), face_times=np.ascontiguousarray(motion["face_times"], np.float32),
                            face_matrix=np.ascontiguousarray(motion["face_matrix"], np.uint32), matrices=np.ascontiguousarray(motion["matrices"], np.float32))
            self._motion = _Motion(*[_p(self._mo[k]) for k in ("kind", "xyz1", "xyz2", "face_times", "face_matrix", "matrices")])
        self.mesh = _Mesh(_p(self.xyz), self.xyz.shape[0], _p(self.idx), n_faces, _p(self.flags),
                          C.cast(C.pointer(self._motion), C.c_void_p) if self._motion is not None else None)
        self.tree = _Tree()
        self._built = None
        if tree is None:
            self._built = L.kdo_build(C.byref(self.mesh), max_leaf, max_depth)
            L.kdo_built_view(self._built, C.byref(self.tree))
            self.n_refs = L.kdo_built_num_refs(self._built)
        else:
            self._keep = {k: np.ascontiguousarray(tree[k]) for k in ("split", "flags", "first_ref", "refs")}
            assert self._keep["split"].dtype == np.float32 and self._keep["flags"].dtype == np.uint32
            self.tree.split = _p(self._keep["split"])
            self.tree.flags = _p(self._keep["flags"])
            self.tree.first_ref = _p(self._keep["first_ref"])
            self.tree.refs = _p(self._keep["refs"])
            self.tree.n_nodes = self._keep["split"].shape[0]
            b = self.tree_bound() if bound is None else np.asarray(bound, dtype=np.float32)
            for i in range(6):
                self.tree.bound[i] = float(b[i])
            self.n_refs = self._keep["refs"].shape[0]
        self.n_nodes = self.tree.n_nodes

    def close(self):
        if self._built:
            lib().kdo_built_free(self._built)
            self._built = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree_bound(self):
        b = np.zeros(6, np.float32)
        lib().kdo_tree_bound(C.byref(self.mesh), _p(b))
        return b

    def bound(self):
        return np.array(list(self.tree.bound), dtype=np.float32)

    def trace_closest(self, rays, threads=1, counters=False, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        prim = np.zeros(n, np.int32)
        if times is not None:
            times = np.ascontiguousarray(times, np.float32)
            lib().kdo_trace_closest_timed(C.byref(self.mesh), C.byref(self.tree), _p(rays), _p(times), n, _p(t), _p(u), _p(v), _p(prim), threads)
            return dict(t=t, u=u, v=v, prim=prim, counters=None)
        cnt = Counters() if counters else None
        lib().kdo_trace_closest(C.byref(self.mesh), C.byref(self.tree), _p(rays), n, _p(t), _p(u), _p(v), _p(prim), threads,
                                C.byref(cnt) if counters else None)
        return dict(t=t, u=u, v=v, prim=prim, counters=cnt)

    def trace_shadow(self, rays, threads=1, counters=False, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        sh = np.zeros(n, np.uint8); prim = np.zeros(n, np.int32)
        if times is not None:
            times = np.ascontiguousarray(times, np.float32)
            lib().kdo_trace_shadow_timed(C.byref(self.mesh), C.byref(self.tree), _p(rays), _p(times), n, _p(sh), _p(prim), threads)
            return dict(shadowed=sh, prim=prim, counters=None)
        cnt = Counters() if counters else None
        lib().kdo_trace_shadow(C.byref(self.mesh), C.byref(self.tree), _p(rays), n, _p(sh), _p(prim), threads,
                               C.byref(cnt) if counters else None)
        return dict(shadowed=sh, prim=prim, counters=cnt)

    def trace_tshadow(self, rays, max_depth, threads=1, max_list=8, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        sh = np.zeros(n, np.uint8); nt = np.zeros(n, np.int32); lst = np.zeros((n, max_list), np.int32)
        if times is not None:
            times = np.ascontiguousarray(times, np.float32)
            lib().kdo_trace_tshadow_timed(C.byref(self.mesh), C.byref(self.tree), _p(rays), _p(times), n, int(max_depth), _p(sh), _p(nt), _p(lst), max_list, threads)
            return dict(shadowed=sh, n_transparent=nt, list=lst)
        lib().kdo_trace_tshadow(C.byref(self.mesh), C.byref(self.tree), _p(rays), n, int(max_depth), _p(sh), _p(nt), _p(lst), max_list, threads)
        return dict(shadowed=sh, n_transparent=nt, list=lst)

    def brute_closest(self, rays, threads=1, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        prim = np.zeros(n, np.int32)
        b = self.bound()
        if times is not None:
            times = np.ascontiguousarray(times, np.float32)
            lib().kdo_brute_closest_timed(C.byref(self.mesh), _p(b), _p(rays), _p(times), n, _p(t), _p(u), _p(v), _p(prim), threads)
            return dict(t=t, u=u, v=v, prim=prim)
        lib().kdo_brute_closest(C.byref(self.mesh), _p(b), _p(rays), n, _p(t), _p(u), _p(v), _p(prim), threads)
        return dict(t=t, u=u, v=v, prim=prim)


def poly_intersect(verts, origin, direction):
    """One polygon test; verts [3|4, 3].  Returns (t, u, v); t == 0 is a miss."""
    vv = np.ascontiguousarray(verts, dtype=np.float32)
    o = np.ascontiguousarray(origin, dtype=np.float32)
    d = np.ascontiguousarray(direction, dtype=np.float32)
    u = C.c_float(0); v = C.c_float(0)
    nv = vv.shape[0]
    rows = [vv[i].copy() for i in range(nv)]
    t = lib().kdo_poly_intersect(_p(rows[0]), _p(rows[1]), _p(rows[2]), _p(rows[3]) if nv == 4 else None, nv, _p(o), _p(d), C.byref(u), C.byref(v))
    return float(t), float(u.value), float(v.value)


def bound_cross(bound6, origin, direction, t_max):
    b = np.ascontiguousarray(bound6, dtype=np.float32)
    o = np.ascontiguousarray(origin, dtype=np.float32)
    d = np.ascontiguousarray(direction, dtype=np.float32)
    e = C.c_float(0); l = C.c_float(0)
    ok = lib().kdo_bound_cross(_p(b), _p(o), _p(d), C.c_float(t_max), C.byref(e), C.byref(l))
    return bool(ok), float(e.value), float(l.value)
