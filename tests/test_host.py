"""CPU tests of the product's host side: the C-ABI library loads and exports what include/b200rt.h declares,
fails loudly without a device, and its kd-tree builder produces trees on which the (restated) reference
traversal finds exactly the reference's hits.  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from libyafaray_b200 import rt, scenes
from oracle import kdo
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    declared = set(re.findall(r"\b(b200rt_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(rt.SYMBOLS), declared ^ set(rt.SYMBOLS)
    L = rt.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libb200rt.so does not export {name}"
    assert L.b200rt_version() == 2


def test_struct_sizes_match_header():
    assert rt.RAY_DTYPE.itemsize == 32 and rt.HIT_DTYPE.itemsize == 16 and rt.TSHADOW_DTYPE.itemsize == 16 + 16 * rt.TSHADOW_MAX
    assert C.sizeof(rt.BuildParams) == 32


def test_no_silent_cpu_fallback(built):
    if rt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.B200RTError) as e:
        rt.Scene(0)
    assert e.value.code in (-2, -3)
    assert rt.lib().b200rt_last_error()


def test_argument_validation(built):
    L = rt.lib()
    assert L.b200rt_create(0, None, None) == -1
    assert L.b200rt_get_bound(None, None) == -1
    assert L.b200rt_trace_closest(None, None, 4, None) == -1
    xyz = np.zeros((3, 3), np.float32)
    bad = np.array([[0, 1, 7, 0xFFFFFFFF]], np.uint32)  # vertex 7 does not exist
    h = C.c_void_p(0)
    assert L.b200rt_host_tree_build(xyz.ctypes.data, 3, bad.ctypes.data, 1, None, C.byref(h)) == -1
    assert b"vertex" in L.b200rt_last_error()


ZOO = helpers.scene_zoo()


@pytest.mark.parametrize("name", sorted(ZOO))
def test_host_tree_is_valid_and_finds_reference_hits(built, name):
    xyz, idx, flags = ZOO[name]
    t = rt.host_tree(xyz, idx)
    a, b, refs = t["a"], t["b"], t["refs"]
    leaf = (b & 3) == 3
    n = len(a)
    # structure: right child in range and after the node; leaf reference ranges tile refs[] in order
    right = b[~leaf] >> 2
    assert np.all(right > np.nonzero(~leaf)[0]) and np.all(right < n)
    counts = (b[leaf] >> 2).astype(np.int64)
    assert np.array_equal(a[leaf].astype(np.int64), np.concatenate([[0], np.cumsum(counts)[:-1]]))
    assert counts.sum() == len(refs)
    assert set(np.unique(refs)) == set(range(idx.shape[0])), "every face must be referenced by some leaf"
    # bound: the reference's inflation arithmetic
    o_ref = kdo.Oracle(xyz, idx, flags)
    assert np.array_equal(t["bound"], o_ref.bound())
    # the reference traversal (restated) over OUR tree vs over the oracle's own tree
    o_mine = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(t), bound=t["bound"])
    closest, shadow = helpers.ray_zoo(t["bound"], n=8000, seed=5)
    r_mine, r_ref = o_mine.trace_closest(closest, threads=4), o_ref.trace_closest(closest, threads=4)
    helpers.check_closest_parity(r_mine["prim"], r_mine["t"], r_mine["u"], r_mine["v"], r_ref,
                                 min_agree=0.98 if name.startswith("cubes") else 0.9999)
    assert np.array_equal(o_mine.trace_shadow(shadow, threads=4)["shadowed"], o_ref.trace_shadow(shadow, threads=4)["shadowed"])


def test_host_tree_parameters(built):
    xyz, idx, _ = scenes.heightfield(48)
    shallow = rt.host_tree(xyz, idx, rt.make_params(depth=6))
    leaf = (shallow["b"] & 3) == 3
    assert leaf.sum() <= 64
    big_leaves = rt.host_tree(xyz, idx, rt.make_params(max_leaf_size=16))
    small_leaves = rt.host_tree(xyz, idx, rt.make_params(max_leaf_size=1))
    assert len(big_leaves["a"]) < len(small_leaves["a"])
    one_thread = rt.host_tree(xyz, idx, rt.make_params(build_threads=1))
    many = rt.host_tree(xyz, idx, rt.make_params(build_threads=8))
    assert np.array_equal(one_thread["a"], many["a"]) and np.array_equal(one_thread["b"], many["b"])  # deterministic


def test_host_tree_empty_and_single(built):
    t = rt.host_tree(np.zeros((0, 3), np.float32), np.zeros((0, 4), np.uint32))
    assert len(t["a"]) == 1 and (t["b"][0] & 3) == 3 and np.array_equal(t["bound"], np.zeros(6, np.float32))
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    t = rt.host_tree(xyz, np.array([[0, 1, 2, 0xFFFFFFFF]], np.uint32))
    assert list(t["refs"]) == [0]
    # a flat mesh has zero extent on one axis: the bound stays flat (reference note at :97-98) and the build survives
    assert t["bound"][2] == 0 and t["bound"][5] == 0
