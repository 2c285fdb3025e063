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
    assert L.b200rt_version() == 3


def test_struct_sizes_match_header():
    assert rt.RAY_DTYPE.itemsize == 32 and rt.HIT_DTYPE.itemsize == 16 and rt.TSHADOW_DTYPE.itemsize == 16 + 16 * rt.TSHADOW_MAX
    assert C.sizeof(rt.BuildParams) == 32
    # b200rt_stats: 8 uint64 counters, 2 uint32, 2 doubles, device_bytes, n_spheres, n_bezier_faces, n_moving_faces (include/b200rt.h)
    assert C.sizeof(rt.Stats) == 8 * 8 + 2 * 4 + 2 * 8 + 8 + 8 + 16
    assert C.sizeof(rt.Job) == 8 + 4 + 4 + 8 + 8 + 8 + 8 + 8  # b200rt_job: max_depth is padded to 8 bytes before the times pointer


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
    closest, shadow = helpers.ray_zoo(t["bound"], n=8000, seed=5, edge_cases=not name.startswith("cube_grid"))
    r_mine, r_ref = o_mine.trace_closest(closest, threads=4), o_ref.trace_closest(closest, threads=4)
    helpers.check_closest_parity(r_mine["prim"], r_mine["t"], r_mine["u"], r_mine["v"], r_ref,
                                 min_agree=0.97 if name.startswith("cube") else 0.9999)
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


def _visited_faces(tree, ray, strict):
    """The kernel's t-interval descent (kd_kernels.cuh) restated on the exported host tree, unbounded stack: the set of faces in
    the leaves a ray opens.  strict=False is the rule the kernel had before the 1-ulp-slab fix."""
    a, b, refs, bound = tree["a"], tree["b"], tree["refs"], tree["bound"]
    split = a.view(np.float32)
    o, d = ray[0:3].astype(np.float32), ray[4:7].astype(np.float32)
    if strict:  # the kernel's traversal copy of the origin: one ulp lower where the direction component is exactly zero
        o = np.where(d == 0, np.nextafter(o, np.float32(-np.inf)), o).astype(np.float32)
    with np.errstate(all="ignore"):
        inv = np.where(d == 0, np.float32(3.4e38), np.float32(1) / np.where(d == 0, np.float32(1), d)).astype(np.float32)
        t0, t1 = (bound[:3] - o) * inv, (bound[3:] - o) * inv
    lo, hi = np.float32(max(np.minimum(t0, t1).max(), 0)), np.float32(np.maximum(t0, t1).min())
    seen = set()
    if not lo <= hi:
        return seen
    stack, node, seg_lo, seg_hi = [], 0, lo, hi
    while True:
        nb = int(b[node]); axis = nb & 3
        if axis == 3:
            seen.update(refs[int(a[node]):int(a[node]) + (nb >> 2)].tolist())
            if not stack:
                return seen
            node, seg_lo, seg_hi = stack.pop()
            continue
        t_plane = np.float32((split[node] - o[axis]) * inv[axis])
        near, far = (node + 1, nb >> 2) if inv[axis] >= 0 else (nb >> 2, node + 1)
        near_only, far_only = (t_plane > seg_hi, t_plane < seg_lo) if strict else (t_plane >= seg_hi, t_plane <= seg_lo)
        if near_only:
            node = near
        elif far_only:
            node = far
        else:
            stack.append((far, t_plane, seg_hi)); node = near; seg_hi = t_plane


def test_strict_interval_rule_opens_one_ulp_slabs(built):
    """Faces that are planar up to one ulp (instanced cubes of the reference's tests/test02) get 1-ulp kd slabs whose entry and
    exit parameter coincide along most rays.  With strict comparisons (t_plane < seg_lo / > seg_hi) every leaf holding the
    reference's hit is opened; with the former <= / >= rule some are skipped."""
    xyz, idx, flags = scenes.cube_grid(5)
    tree = rt.host_tree(xyz, idx)
    closest, _ = helpers.ray_zoo(tree["bound"], n=3000, seed=5)
    ref = kdo.Oracle(xyz, idx, flags).trace_closest(closest, threads=4)
    hit = np.nonzero(ref["prim"] >= 0)[0][:1500]
    missed_strict = sum(int(ref["prim"][i]) not in _visited_faces(tree, closest[i], True) for i in hit)
    missed_loose = sum(int(ref["prim"][i]) not in _visited_faces(tree, closest[i], False) for i in hit)
    assert missed_strict == 0
    assert missed_loose > 0, "the scene no longer produces the thin slabs this test is about"


def test_rays_lying_in_split_planes_take_the_reference_side(built):
    """A ray whose direction component is exactly zero and whose origin lies exactly on a split plane of that axis: the
    reference descends into the LEFT child only (accelerator_kdtree_common.h:148-174).  The kernel's rule (origin copy one ulp
    lower on such axes) must open every leaf in which the reference traversal -- run over the same tree -- finds its hit."""
    xyz, idx, flags = scenes.cube_grid(5)
    tree = rt.host_tree(xyz, idx)
    a, b = tree["a"], tree["b"]
    split = a.view(np.float32)
    interior = np.nonzero((b & 3) != 3)[0]
    rng = np.random.default_rng(1)
    lo, ext = tree["bound"][:3], tree["bound"][3:] - tree["bound"][:3]
    rays = np.zeros((1200, 8), np.float32)
    for k, node in enumerate(rng.choice(interior, size=rays.shape[0])):
        axis = int(b[node]) & 3
        o = (lo - 0.3 * ext + rng.random(3).astype(np.float32) * 1.6 * ext).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        o[axis], d[axis] = split[node], 0.0
        rays[k] = [o[0], o[1], o[2], 0.0, d[0], d[1], d[2], -1.0]
    same_tree = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(tree), bound=tree["bound"]).trace_closest(rays, threads=4)
    hit = np.nonzero(same_tree["prim"] >= 0)[0]
    assert len(hit) > 300
    assert sum(int(same_tree["prim"][i]) not in _visited_faces(tree, rays[i], True) for i in hit) == 0


@pytest.mark.parametrize("name", [n for n in sorted(ZOO) if not n.endswith("_flags")])
def test_kernel_rule_opens_every_leaf_in_which_the_reference_finds_its_hit(built, name):
    """The kernel's descent rule restated on the CPU (_visited_faces: strict t-interval comparisons, origin copy one ulp lower on
    zero-direction axes) against the REFERENCE traversal run over the same exported tree: for primary, edge-case and
    on-surface rays (tmin 0) every leaf that holds the reference's hit must be among the leaves the rule opens."""
    import warnings
    xyz, idx, flags = ZOO[name]
    tree = rt.host_tree(xyz, idx)
    same_tree = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(tree), bound=tree["bound"])
    closest, _ = helpers.ray_zoo(tree["bound"], n=400, seed=9)
    surface = helpers.surface_rays(closest, same_tree.trace_closest(closest, threads=4)["t"], seed=10, tmin=0.0)[:600]
    rays = np.concatenate([closest, surface])
    ref = same_tree.trace_closest(rays, threads=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        missed = [int(i) for i in np.nonzero(ref["prim"] >= 0)[0] if int(ref["prim"][i]) not in _visited_faces(tree, rays[i], True)]
    assert not missed, missed[:10]


def _leaf_order(tree, ray, ring_entries=None):
    """Leaves in the order the kernel's descent opens them.  ring_entries=None: unbounded stack.  Otherwise the kernel's short
    stack restated: a ring that keeps the newest `ring_entries` - 1 usable entries (pushes are unconditional stores), `floor`
    marking lost entries, and the exact kd-restart (replayTo in kd_kernels.cuh): replay the descent from the root along the path
    to the leaf just left -- "target < right" tells the child in depth-first order -- re-postponing the far children, then pop."""
    a, b, bound = tree["a"], tree["b"], tree["bound"]
    split = a.view(np.float32)
    o, d = ray[0:3].astype(np.float32), ray[4:7].astype(np.float32)
    o = np.where(d == 0, np.nextafter(o, np.float32(-np.inf)), o).astype(np.float32)
    with np.errstate(all="ignore"):
        inv = np.where(d == 0, np.float32(3.4e38), np.float32(1) / np.where(d == 0, np.float32(1), d)).astype(np.float32)
        t0, t1 = (bound[:3] - ray[0:3].astype(np.float32)) * inv, (bound[3:] - ray[0:3].astype(np.float32)) * inv
        lo, hi = np.float32(max(np.minimum(t0, t1).max(), 0)), np.float32(np.maximum(t0, t1).min())
    order = []
    if not lo <= hi:
        return order
    t_exit = hi
    ring, sp, floor = {}, 0, 0   # ring slot -> entry; sp counts pushes minus pops; entries below floor are lost

    def plane(node):
        axis = int(b[node]) & 3
        with np.errstate(all="ignore"):
            t_plane = np.float32((split[node] - o[axis]) * inv[axis])
        left, right = node + 1, int(b[node]) >> 2
        return t_plane, ((left, right) if inv[axis] >= 0 else (right, left)), right, inv[axis] < 0

    def push(entry):
        nonlocal sp, floor
        if ring_entries is None:
            ring[sp] = entry
        else:
            ring[sp % ring_entries] = entry
            floor = max(floor, sp + 1 - ring_entries + 1)  # the kernel's speculative store also clobbers one more slot: K - 1 usable
        sp += 1

    node, seg_lo, seg_hi = 0, lo, hi
    for _ in range(100000):
        if (int(b[node]) & 3) != 3:
            t_plane, (near, far), _, _ = plane(node)
            if t_plane > seg_hi:
                node = near
            elif t_plane < seg_lo:
                node = far
            else:
                push((far, seg_hi)); node = near; seg_hi = t_plane
            continue
        order.append(node)
        if sp > floor:
            sp -= 1
            entry = ring[sp if ring_entries is None else sp % ring_entries]
            node, seg_lo, seg_hi = entry[0], seg_hi, entry[1]
            continue
        if floor == 0 or not seg_hi < t_exit:
            return order
        # exact kd-restart
        target, sp, floor, cur, r_hi = node, 0, 0, 0, t_exit
        while cur != target:
            t_plane, (near, far), right, negative = plane(cur)
            if (target < right) != negative:
                if not t_plane > r_hi:
                    push((far, r_hi)); r_hi = t_plane
                cur = near
            else:
                cur = far
        if sp <= floor:
            return order
        sp -= 1
        entry = ring[sp % ring_entries]
        node, seg_lo, seg_hi = entry[0], r_hi, entry[1]
    raise AssertionError("the descent did not terminate")


@pytest.mark.parametrize("name", ["cube_grid", "objects", "soup"])
def test_replay_restart_visits_the_same_leaves_in_the_same_order(built, name):
    """The exact kd-restart, restated: with a ring of 2, 3 or 8 entries the descent must open exactly the leaves, in exactly the
    order, of the unbounded-stack descent -- and terminate -- also through the zero-length intervals of the thin-slab scene."""
    xyz, idx, flags = ZOO[name]
    tree = rt.host_tree(xyz, idx)
    closest, _ = helpers.ray_zoo(tree["bound"], n=150, seed=12)
    for ray in closest[::3]:
        want = _leaf_order(tree, ray)
        for entries in (2, 3, 8):
            assert _leaf_order(tree, ray, entries) == want, (name, entries, ray.tolist())
