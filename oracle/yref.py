"""ctypes binding of oracle/_ref/libyafref.so -- the UNMODIFIED reference behind oracle/ref_driver.cc.

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline / --impl reference legs and the
golden-vector generator.  Nothing under libyafaray_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libyafref.so")

_VIS = {3: "normal", 0: "invisible", 2: "shadow_only", 1: "no_shadows"}


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.yref_scene_create.restype = C.c_void_p
        L.yref_scene_create.argtypes = [C.c_int]
        L.yref_scene_destroy.argtypes = [C.c_void_p]
        L.yref_add_material.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
        L.yref_add_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_char_p]
        L.yref_add_mesh_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_char_p, C.c_int, C.c_float, C.c_float]
        L.yref_add_instance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.yref_set_times.argtypes = [C.c_void_p, C.c_void_p]
        L.yref_add_sphere.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_char_p]
        L.yref_add_sphere.restype = C.c_int
        L.yref_build.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.yref_build_seconds.restype = C.c_double
        L.yref_build_seconds.argtypes = [C.c_void_p]
        L.yref_num_prims.restype = C.c_size_t
        L.yref_num_prims.argtypes = [C.c_void_p]
        L.yref_get_bound.argtypes = [C.c_void_p, C.c_void_p]
        for f in (L.yref_trace_closest, L.yref_trace_shadow, L.yref_trace_tshadow):
            f.restype = C.c_double
        L.yref_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.yref_trace_shadow.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.yref_trace_tshadow.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.yref_tree_counts.restype = C.c_int64
        L.yref_tree_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.yref_tree_export.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.yref_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """One reference scene: a single mesh object whose faces carry per-face flags.

    Per-face flag bits (libyafaray_b200/scenes.py): the reference keeps visibility on the object AND on the
    material (include/accelerator/accelerator.h:126-127); here the object stays "normal" and each distinct
    (visibility, transparent) combination becomes one shinydiffuse material, which is equivalent for the
    accept rules.  `transparency` is the material parameter used for faces with the transparent bit.
    """

    def __init__(self, xyz, idx, flags=None, *, accel_type=None, depth=-1, max_leaf_size=-1, cost_ratio=-1.0,
                 empty_bonus=-1.0, transparency=0.5, verbose=False, motion=None):
        L = lib()
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        n_faces = idx.shape[0]
        if flags is None:
            flags = np.full(n_faces, 3, dtype=np.uint8)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        self.h = L.yref_scene_create(1 if verbose else 0)
        combos = sorted(set(int(f) & 7 for f in np.unique(flags))) or [3]
        mat_of = {}
        for cmb in combos:
            m = L.yref_add_material(self.h, _VIS[cmb & 3].encode(), float(transparency) if (cmb & 4) else 0.0)
            if m < 0:
                raise RuntimeError("reference refused the material")
            mat_of[cmb] = m
        face_mat = np.array([mat_of[int(f) & 7] for f in flags], dtype=np.int32) if len(combos) > 1 else np.zeros(n_faces, dtype=np.int32)
        # sphere faces (idx[:, 2] == 0xFFFFFFFE, libyafaray_b200/scenes.py::with_spheres) become "sphere" objects created
        # after the mesh, so that the reference's primitive order (objects in creation order, src/scene/scene.cc:320-326)
        # is the face order of the arrays; they must therefore be the LAST faces
        is_sphere = idx[:, 2] == 0xFFFFFFFE if n_faces else np.zeros(0, bool)
        n_mesh = int(n_faces - is_sphere.sum())
        if is_sphere[:n_mesh].any():
            raise ValueError("sphere faces must follow all mesh faces")
        if motion is not None:
            # Motion blur (libyafaray_b200/scenes.py::motion_scene): runs of static faces become plain meshes, runs of Bezier faces with
            # one time range become motion-blur meshes, runs of faces of one moving instance become a base object plus an instance
            # with three matrices.  The reference lists all object primitives before the instance primitives
            # (src/scene/scene.cc:320-341), so the instance faces must be the last ones for face ids to equal array rows.
            kind = np.ascontiguousarray(motion["kind"], np.uint8)
            # the client passes the mid-time POSITIONS; the mesh turns them into the control points motion["xyz1"] holds
            xyz1 = np.ascontiguousarray(motion["xyz1_user"], np.float32); xyz2 = np.ascontiguousarray(motion["xyz2"], np.float32)
            ft = np.asarray(motion["face_times"], np.float32); fm = np.asarray(motion["face_matrix"], np.uint32)
            mats = np.asarray(motion["matrices"], np.float64)
            if is_sphere.any():
                raise ValueError("motion scenes with spheres are not built by this driver")
            seen_instance = False
            f = 0
            while f < n_faces:
                e = f
                while e < n_faces and kind[e] == kind[f] and np.array_equal(ft[e], ft[f]) and (kind[f] != 2 or fm[e] == fm[f]):
                    e += 1
                sub_idx = np.ascontiguousarray(idx[f:e]); sub_mat = np.ascontiguousarray(face_mat[f:e])
                if kind[f] == 2:
                    seen_instance = True
                    oid = L.yref_add_mesh_ex(self.h, _p(xyz), None, None, xyz.shape[0], _p(sub_idx), e - f, _p(sub_mat), b"normal", 1, 0.0, 0.0)
                    m3 = np.ascontiguousarray(mats[fm[f]].reshape(3, 16))
                    tt = np.array([ft[f, 0], 0.5 * (float(ft[f, 0]) + float(ft[f, 1])), ft[f, 1]], np.float32)
                    if oid < 0 or L.yref_add_instance(self.h, oid, _p(m3), _p(tt), 3) < 0:
                        raise RuntimeError("reference refused the moving instance")
                else:
                    if seen_instance:
                        raise ValueError("faces of moving instances must follow all other faces")
                    bez = kind[f] == 1
                    if L.yref_add_mesh_ex(self.h, _p(xyz), _p(xyz1) if bez else None, _p(xyz2) if bez else None, xyz.shape[0], _p(sub_idx), e - f,
                                          _p(sub_mat), b"normal", 0, float(ft[f, 0]), float(ft[f, 1])) < 0:
                        raise RuntimeError("reference refused the mesh")
                f = e
        elif n_mesh and L.yref_add_mesh(self.h, _p(xyz), xyz.shape[0], _p(np.ascontiguousarray(idx[:n_mesh])), n_mesh, _p(np.ascontiguousarray(face_mat[:n_mesh])), b"normal") < 0:
            raise RuntimeError("reference refused the mesh")
        for f in (range(n_mesh, n_faces) if motion is None else ()):
            c = xyz[idx[f, 0]]
            if L.yref_add_sphere(self.h, float(c[0]), float(c[1]), float(c[2]), float(xyz[idx[f, 1], 0]), int(face_mat[f]), b"normal") < 0:
                raise RuntimeError("reference refused the sphere")
        rc = L.yref_build(self.h, accel_type.encode() if accel_type else None, depth, max_leaf_size, cost_ratio, empty_bonus)
        if rc != 0:
            raise RuntimeError(f"reference preprocess failed ({rc})")
        self.n_prims = L.yref_num_prims(self.h)
        self.build_seconds = L.yref_build_seconds(self.h)

    def close(self):
        if self.h:
            lib().yref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bound(self):
        b = np.zeros(6, dtype=np.float32)
        lib().yref_get_bound(self.h, _p(b))
        return b

    def _set_times(self, times, n):
        if times is None:
            lib().yref_set_times(self.h, None)
            return None
        t = np.ascontiguousarray(times, np.float32)
        assert t.shape[0] == n
        lib().yref_set_times(self.h, _p(t))
        return t

    def trace_closest(self, rays, threads=0, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        _keep = self._set_times(times, rays.shape[0])
        n = rays.shape[0]
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        prim = np.zeros(n, np.int32)
        secs = lib().yref_trace_closest(self.h, _p(rays), n, _p(t), _p(u), _p(v), _p(prim), threads)
        return dict(t=t, u=u, v=v, prim=prim, seconds=secs)

    def trace_shadow(self, rays, threads=0, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        _keep = self._set_times(times, rays.shape[0])
        n = rays.shape[0]
        sh = np.zeros(n, np.uint8); prim = np.zeros(n, np.int32)
        secs = lib().yref_trace_shadow(self.h, _p(rays), n, _p(sh), _p(prim), threads)
        return dict(shadowed=sh, prim=prim, seconds=secs)

    def trace_tshadow(self, rays, max_depth, threads=0, times=None):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        _keep = self._set_times(times, rays.shape[0])
        n = rays.shape[0]
        sh = np.zeros(n, np.uint8); rgb = np.zeros((n, 3), np.float32)
        secs = lib().yref_trace_tshadow(self.h, _p(rays), n, int(max_depth), _p(sh), _p(rgb), threads)
        return dict(shadowed=sh, rgb=rgb, seconds=secs)

    def export_tree(self):
        """The reference's own kd-tree, flat: split f32[n], flags u32[n] (Node::flags_ verbatim),
        first_ref u32[n], refs u32[n_refs] (primitive indices in factory order)."""
        L = lib()
        n_refs = C.c_int64(0)
        n = L.yref_tree_counts(self.h, C.byref(n_refs))
        if n < 0:
            raise RuntimeError("accelerator is not yafaray-kdtree-original")
        split = np.zeros(n, np.float32); flags = np.zeros(n, np.uint32); first = np.zeros(n, np.uint32)
        refs = np.zeros(max(1, n_refs.value), np.uint32)
        L.yref_tree_export(self.h, _p(split), _p(flags), _p(first), _p(refs))
        return dict(split=split, flags=flags, first_ref=first, refs=refs[: n_refs.value])


def hardware_threads() -> int:
    return lib().yref_hardware_threads()
