"""ctypes bindings of the photon-map checkers: oracle/libkdoracle.so (pm_oracle.c, the plain-C restatement) and
oracle/_ref/libyafref.so (ref_pm_driver.cc around the UNMODIFIED reference's PhotonMap).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke(), the golden-vector generator and the CPU legs of
tools/pm_bench.py.  Nothing under libyafaray_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import kdo, yref

MISS = 0xFFFFFFFF


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, cols=3):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


class _Base:
    """Common result handling: gather() returns (idx [n,k] u32, d2 [n,k] f32, n_found [n] u32, radius_out [n] f32);
    entries past n_found are MISS / 0."""

    def _alloc(self, n, k):
        return (np.full((n, k), MISS, np.uint32), np.zeros((n, k), np.float32), np.zeros(n, np.uint32), np.zeros(n, np.float32))


_olib = None


def _oracle_lib():
    global _olib
    if _olib is None:
        kdo.build_library()
        L = C.CDLL(kdo.LIB_PATH)
        L.pmo_create.restype = C.c_void_p
        L.pmo_create.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.pmo_destroy.argtypes = [C.c_void_p]
        L.pmo_tree_export.restype = C.c_int64
        L.pmo_tree_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pmo_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_float, C.c_void_p] + [C.c_void_p] * 4
        L.pmo_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
        L.pmo_set_unpruned.argtypes = [C.c_int]
        _olib = L
    return _olib


def set_unpruned(on: bool):
    """Cross-check mode of pm_oracle.c: visit every leaf in the lookup's order, no split-plane pruning."""
    _oracle_lib().pmo_set_unpruned(1 if on else 0)


class OracleMap(_Base):
    """pm_oracle.c"""

    def __init__(self, pos, dirs=None):
        self.pos = _f32(pos)
        self.dirs = None if dirs is None else _f32(dirs)
        self.h = _oracle_lib().pmo_create(_p(self.pos), _p(self.dirs), len(self.pos))
        if not self.h:
            raise ValueError("empty photon map")

    def close(self):
        if self.h:
            _oracle_lib().pmo_destroy(self.h)
            self.h = None

    __del__ = close

    def tree(self):
        n = _oracle_lib().pmo_tree_export(self.h, None, None)
        a, b = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        _oracle_lib().pmo_tree_export(self.h, _p(a), _p(b))
        return a, b

    def gather(self, points, k, sq_radius, sq_radii=None):
        points = _f32(points)
        out = self._alloc(len(points), k)
        radii = None if sq_radii is None else np.ascontiguousarray(sq_radii, np.float32)
        _oracle_lib().pmo_gather(self.h, _p(points), len(points), k, float(sq_radius), _p(radii), *[_p(o) for o in out])
        return out

    def nearest(self, points, normals, dist):
        points, normals = _f32(points), _f32(normals)
        out = np.zeros(len(points), np.uint32)
        _oracle_lib().pmo_nearest(self.h, _p(points), _p(normals), len(points), float(dist), _p(out))
        return out


_rlib = None


def ref_available() -> bool:
    if not yref.available():
        return False
    try:
        _ref_lib()
        return True
    except AttributeError:  # a libyafref.so from before ref_pm_driver.cc
        return False


def _ref_lib():
    global _rlib
    if _rlib is None:
        L = C.CDLL(yref.LIB_PATH)
        L.yref_pm_create.restype = C.c_void_p
        L.yref_pm_create.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.yref_pm_destroy.argtypes = [C.c_void_p]
        L.yref_pm_build_seconds.restype = C.c_double
        L.yref_pm_build_seconds.argtypes = [C.c_void_p]
        L.yref_pm_tree_export.restype = C.c_int64
        L.yref_pm_tree_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.yref_pm_gather.restype = C.c_double
        L.yref_pm_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_float, C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        L.yref_pm_nearest.restype = C.c_double
        L.yref_pm_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p, C.c_int]
        _rlib = L
    return _rlib


class RefMap(_Base):
    """The unmodified reference's PhotonMap (oracle/_ref)."""

    def __init__(self, pos, dirs=None, build_threads=1, query_threads=1):
        self.pos = _f32(pos)
        self.dirs = None if dirs is None else _f32(dirs)
        self.query_threads = query_threads
        self.h = _ref_lib().yref_pm_create(_p(self.pos), _p(self.dirs), len(self.pos), build_threads)
        self.build_seconds = _ref_lib().yref_pm_build_seconds(self.h)
        self.seconds = 0.0

    def close(self):
        if self.h:
            _ref_lib().yref_pm_destroy(self.h)
            self.h = None

    __del__ = close

    def tree(self):
        n = _ref_lib().yref_pm_tree_export(self.h, None, None)
        a, b = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        _ref_lib().yref_pm_tree_export(self.h, _p(a), _p(b))
        return a, b

    def gather(self, points, k, sq_radius, sq_radii=None):
        points = _f32(points)
        out = self._alloc(len(points), k)
        radii = None if sq_radii is None else np.ascontiguousarray(sq_radii, np.float32)
        self.seconds = _ref_lib().yref_pm_gather(self.h, _p(points), len(points), k, float(sq_radius), _p(radii), *[_p(o) for o in out], self.query_threads)
        return out

    def nearest(self, points, normals, dist):
        points, normals = _f32(points), _f32(normals)
        out = np.zeros(len(points), np.uint32)
        self.seconds = _ref_lib().yref_pm_nearest(self.h, _p(points), _p(normals), len(points), float(dist), _p(out), self.query_threads)
        return out
