"""CPU tests of the oracle (oracle/kd_oracle.c): pinned against the golden vectors generated from the
unmodified reference, and against the live reference (oracle/_ref/libyafref.so) where it was built."""
import glob
import os

import numpy as np
import pytest

from libyafaray_b200 import scenes
from oracle import kdo, yref
from tests import helpers

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _golden_tree(g):
    return dict(split=g["tree_split"], flags=g["tree_flags"], first_ref=g["tree_first_ref"], refs=g["tree_refs"])


def test_golden_files_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_bit_exact_on_reference_tree(built, path):
    """Restated traversal + polygon test + accept rules on the reference's own tree == the reference, ties included."""
    g = np.load(path)
    o = kdo.Oracle(g["xyz"], g["idx"], g["flags"], tree=_golden_tree(g))
    assert np.array_equal(o.bound(), g["bound"])  # tree-bound arithmetic (accelerator_kdtree_original.cc:88-103)
    c = o.trace_closest(g["closest_rays"])
    assert np.array_equal(c["prim"], g["closest_prim"])
    for k in ("t", "u", "v"):
        assert np.array_equal(c[k], g["closest_" + k]), k
    s = o.trace_shadow(g["shadow_rays"])
    assert np.array_equal(s["shadowed"], g["shadow_shadowed"])
    assert np.array_equal(s["prim"], g["shadow_prim"])
    depth = int(g["tshadow_depth"])
    t = o.trace_tshadow(g["shadow_rays"], depth)
    assert np.array_equal(t["shadowed"], g["tshadow_shadowed"])
    # colour = product of per-occluder transparencies (0.5 * (0.8, 0.6, 0.4) for the generator's material)
    lit = g["tshadow_shadowed"] == 0
    expect = np.power(np.float64(0.4), t["n_transparent"][lit])
    assert np.allclose(g["tshadow_rgb"][lit, 0], expect, rtol=1e-5)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_own_tree_meets_parity_bar(built, path):
    """With a different (valid) kd-tree only exact-t ties may change."""
    g = np.load(path)
    o = kdo.Oracle(g["xyz"], g["idx"], g["flags"])
    c = o.trace_closest(g["closest_rays"])
    ref = dict(prim=g["closest_prim"], t=g["closest_t"], u=g["closest_u"], v=g["closest_v"])
    min_agree = 0.98 if "cubes" in path else 0.9999  # the cube scene has coplanar faces: genuine ties
    helpers.check_closest_parity(c["prim"], c["t"], c["u"], c["v"], ref, min_agree=min_agree)
    s = o.trace_shadow(g["shadow_rays"])
    assert np.array_equal(s["shadowed"], g["shadow_shadowed"])
    t = o.trace_tshadow(g["shadow_rays"], int(g["tshadow_depth"]))
    assert np.array_equal(t["shadowed"], g["tshadow_shadowed"])


@pytest.mark.parametrize("path", [p for p in GOLDEN if "cubes" in p or "hf_quads" in p])
def test_brute_force_agrees(built, path):
    g = np.load(path)
    o = kdo.Oracle(g["xyz"], g["idx"], g["flags"])
    b = o.brute_closest(g["closest_rays"])
    ref = dict(prim=g["closest_prim"], t=g["closest_t"], u=g["closest_u"], v=g["closest_v"])
    helpers.check_closest_parity(b["prim"], b["t"], b["u"], b["v"], ref, min_agree=0.98)


@pytest.mark.skipif(not yref.available(), reason="oracle/_ref/libyafref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["hf_flags", "hf_quads_flags", "cubes_flags", "soup_flags", "objects_flags", "spheres_flags", "spheres"])
def test_oracle_vs_live_reference(built, name):
    xyz, idx, flags = helpers.scene_zoo()[name]
    ref = yref.RefScene(xyz, idx, flags)
    closest, shadow = helpers.ray_zoo(ref.bound(), n=20000, seed=77)
    o = kdo.Oracle(xyz, idx, flags, tree=ref.export_tree())
    assert np.array_equal(o.bound(), ref.bound())
    a, b = ref.trace_closest(closest, threads=4), o.trace_closest(closest, threads=4)
    for k in ("prim", "t", "u", "v"):
        assert np.array_equal(a[k], b[k]), k
    a, b = ref.trace_shadow(shadow, threads=4), o.trace_shadow(shadow, threads=4)
    assert np.array_equal(a["shadowed"], b["shadowed"]) and np.array_equal(a["prim"], b["prim"])
    for depth in (0, 1, 4):
        a, b = ref.trace_tshadow(shadow, depth, threads=4), o.trace_tshadow(shadow, depth, threads=4)
        assert np.array_equal(a["shadowed"], b["shadowed"]), depth
    ref.close()


@pytest.mark.skipif(not yref.available(), reason="oracle/_ref/libyafref.so not built (needs /root/reference)")
def test_reference_accelerator_types_agree(built):
    """The reference's own two kd-tree types give the same answers (sanity of the oracle driver)."""
    xyz, idx, flags = scenes.objects(8000, n_spheres=10)
    a = yref.RefScene(xyz, idx, flags)
    b = yref.RefScene(xyz, idx, flags, accel_type="yafaray-kdtree-multi-thread")
    rays = scenes.rays_incoherent(20000, seed=2)
    ra, rb = a.trace_closest(rays, threads=4), b.trace_closest(rays, threads=4)
    assert np.mean(ra["prim"] == rb["prim"]) >= 0.9999
    a.close(); b.close()


# ---- unit cases of the two leaf functions -------------------------------------------------------
def test_poly_intersect_triangle_cases(built):
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    t, u, v = kdo.poly_intersect(tri, [0.25, 0.25, 1], [0, 0, -1])
    assert t == 1.0 and u == 0.25 and v == 0.25
    assert kdo.poly_intersect(tri, [0.25, 0.25, -1], [0, 0, -1])[0] == 0.0      # behind the origin
    assert kdo.poly_intersect(tri, [0.25, 0.25, 1], [1, 0, 0])[0] == 0.0        # parallel: det == 0
    assert kdo.poly_intersect(tri, [0.75, 0.75, 1], [0, 0, -1])[0] == 0.0       # u + v > 1
    assert kdo.poly_intersect(tri, [0.0, 0.0, 1], [0, 0, -1])[0] == 1.0         # a vertex counts (u = v = 0)
    assert kdo.poly_intersect(tri, [0.5, 0.5, 1], [0, 0, -1])[0] == 1.0         # the hypotenuse counts (u + v = 1)
    assert kdo.poly_intersect(tri, [0.25, 0.25, 1], [0, 0, 1])[0] == 0.0        # pointing away
    t2, _, _ = kdo.poly_intersect(tri, [0.25, 0.25, -1], [0, 0, 1])             # no back-face culling
    assert t2 == 1.0


def test_poly_intersect_quad_rules(built):
    quad = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    # first triangle (v0 v1 v2): uv = {u+v, v}
    t, u, v = kdo.poly_intersect(quad, [0.75, 0.25, 1], [0, 0, -1])
    assert t == 1.0 and (u, v) == (0.75, 0.25)
    # the second triangle (v0 v2 v3) is tried only when the first fails its u range test
    # (shape_polygon.h:139,153); its uv = {u, u+v}
    t, u, v = kdo.poly_intersect(quad, [0.25, 0.75, 1], [0, 0, -1])
    assert t == 1.0 and (u, v) == (0.25, 0.75)
    # outside both
    assert kdo.poly_intersect(quad, [1.5, 0.5, 1], [0, 0, -1])[0] == 0.0
    # a non-convex "quad": a point inside the second triangle that passes the first triangle's u test but
    # fails its v test is reported as a miss -- the reference's rule, reproduced as is
    dart = np.array([[0, 0, 0], [1, 0, 0], [0.2, 0.2, 0], [0, 1, 0]], np.float32)
    assert kdo.poly_intersect(dart, [0.05, 0.5, 1], [0, 0, -1])[0] in (0.0, 1.0)


def test_bound_cross_cases(built):
    b = [0, 0, 0, 1, 1, 1]
    ok, e, l = kdo.bound_cross(b, [-1, 0.5, 0.5], [1, 0, 0], 3.4e38)
    assert ok and e == 1.0 and l == 2.0
    assert not kdo.bound_cross(b, [-1, 0.5, 0.5], [-1, 0, 0], 3.4e38)[0]          # pointing away
    assert not kdo.bound_cross(b, [-1, 0.5, 0.5], [1, 0, 0], 0.5)[0]              # t_max before the box
    ok, e, l = kdo.bound_cross(b, [0.5, 0.5, 0.5], [0, 0, 1], 3.4e38)             # inside, two zero components
    assert ok and e == -0.5 and l == 0.5
    # a zero component skips that axis entirely -- even when the origin is outside the slab (reference behaviour)
    assert kdo.bound_cross(b, [0.5, 2.0, 0.5], [0, 0, 1], 3.4e38)[0]


def test_empty_and_degenerate(built):
    xyz = np.zeros((0, 3), np.float32)
    idx = np.zeros((0, 4), np.uint32)
    o = kdo.Oracle(xyz, idx)
    assert np.array_equal(o.bound(), np.zeros(6, np.float32))
    r = o.trace_closest(scenes.rays_incoherent(100))
    assert np.all(r["prim"] == -1) and np.all(r["t"] == 0)
    # a zero-area triangle never hits (det == 0)
    xyz = np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]], np.float32)
    idx = np.array([[0, 1, 2, 0xFFFFFFFF]], np.uint32)
    o = kdo.Oracle(xyz, idx)
    assert np.all(o.trace_closest(scenes.rays_incoherent(1000, lo=(0, 0, 0), hi=(2, 2, 2)))["prim"] == -1)


@pytest.mark.skipif(not yref.available(), reason="oracle/_ref/libyafref.so was not built")
def test_reference_itself_is_not_exact_on_cube_grid_far(built):
    """Pins the claim behind the tolerance of tests/test_gpu.py::test_adversarial_scenes[cube_grid_far] WITHOUT any code of
    ours in the loop: at coordinates of 300 with cube faces planar only up to one ulp, the UNMODIFIED reference on ITS OWN
    kd-tree returns, for a few rays in 100 000, something else than a brute-force test of every primitive with the same
    arithmetic and accept rules (it misses hits outright or reports a farther one: rays grazing cube edges walk through a
    slab that is thinner than the resolution of its stored entry/exit points).  So on this scene no tree -- the reference's
    included -- reproduces "the" answer, and two valid trees may differ by that much.  The rate is bounded here; the GPU test
    holds the kernel to the same bound against brute force and to EXACT agreement with the reference traversal restated on
    the same tree."""
    xyz, idx, flags = helpers.cube_grid_far()
    ref = yref.RefScene(xyz, idx, flags)
    b = ref.bound().astype(np.float64)
    ext = np.maximum(b[3:] - b[:3], 1e-3)
    primary = scenes.rays_incoherent(300000, seed=51, lo=b[:3] - 0.3 * ext, hi=b[3:] + 0.3 * ext)
    o = kdo.Oracle(xyz, idx, flags)
    ncpu = os.cpu_count() or 1
    r1 = ref.trace_closest(primary, threads=ncpu)
    surface = helpers.surface_rays(primary, r1["t"], seed=52, tmin=0.0005 * float(np.linalg.norm(ext)))
    total_bad = 0
    for rays in (primary, surface):
        r = ref.trace_closest(rays, threads=ncpu)
        brute = o.brute_closest(rays, threads=ncpu)
        _, n_diff, n_bad = helpers.parity_counts(r["prim"].astype(np.int64), r["t"], brute)
        assert n_bad <= 2.5e-4 * rays.shape[0], (n_diff, n_bad)  # measured: 10 of 300 000 primary, 8 of 66 693 on-surface rays
        total_bad += n_bad
    assert total_bad > 0, "the reference agreed with brute force everywhere: tighten test_adversarial_scenes[cube_grid_far]"
    ref.close()


# ---------------------------------------------------------------------------------------------- motion blur
MOTION_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "motion", "motion.npz")


def _golden_motion(g):
    return {k[len("motion_"):]: g[k] for k in g.files if k.startswith("motion_")}


def test_oracle_motion_blur_bit_exact_on_reference_tree(built):
    """Bezier motion-blur faces and faces of moving instances at per-ray times: the restated vertex interpolation (Bezier factors,
    interpolated instance matrix, matrix * point) + polygon test on the reference's own tree == the unmodified reference, bit
    for bit (golden vector generated by tests/golden/make_golden.py motion)."""
    g = np.load(MOTION_GOLDEN)
    mo = _golden_motion(g)
    o = kdo.Oracle(g["xyz"], g["idx"], g["flags"], tree=_golden_tree(g), motion=mo)
    assert np.array_equal(o.bound(), g["bound"])  # the tree bound covers every time step / matrix
    c = o.trace_closest(g["closest_rays"], times=g["closest_times"])
    assert np.array_equal(c["prim"], g["closest_prim"])
    for k in ("t", "u", "v"):
        assert np.array_equal(c[k], g["closest_" + k]), k
    hit = g["closest_prim"] >= 0
    assert set(np.unique(mo["kind"][g["closest_prim"][hit]])) == {0, 1, 2}, "the golden rays must hit all three kinds of faces"
    s = o.trace_shadow(g["shadow_rays"], times=g["shadow_times"])
    assert np.array_equal(s["shadowed"], g["shadow_shadowed"]) and np.array_equal(s["prim"], g["shadow_prim"])
    t = o.trace_tshadow(g["shadow_rays"], int(g["tshadow_depth"]), times=g["shadow_times"])
    assert np.array_equal(t["shadowed"], g["tshadow_shadowed"])
    # time matters: the same rays at time 0 give other answers
    c0 = o.trace_closest(g["closest_rays"], times=np.zeros_like(g["closest_times"]))
    assert (c0["prim"] != c["prim"]).mean() > 0.01


@pytest.mark.skipif(not yref.available(), reason="oracle/_ref/libyafref.so was not built")
def test_oracle_motion_blur_vs_live_reference(built):
    xyz, idx, flags, mo = scenes.motion_scene(seed=11)
    flags = helpers.flag_mix(idx.shape[0], seed=12)
    ref = yref.RefScene(xyz, idx, flags, motion=mo)
    o = kdo.Oracle(xyz, idx, flags, tree=ref.export_tree(), motion=mo)
    assert np.array_equal(o.bound(), ref.bound())
    b = ref.bound()
    ncpu = os.cpu_count() or 1
    rays = scenes.rays_incoherent(100000, seed=13, lo=b[:3], hi=b[3:])
    times = scenes.ray_times(rays.shape[0], 14)
    r, c = ref.trace_closest(rays, threads=ncpu, times=times), o.trace_closest(rays, threads=ncpu, times=times)
    assert np.array_equal(c["prim"], r["prim"]) and np.array_equal(c["t"], r["t"]) and np.array_equal(c["u"], r["u"]) and np.array_equal(c["v"], r["v"])
    srays = scenes.rays_shadow(100000, seed=15, lo=b[:3], hi=b[3:], t_max=0.4)
    assert np.array_equal(o.trace_shadow(srays, threads=ncpu, times=times)["shadowed"], ref.trace_shadow(srays, threads=ncpu, times=times)["shadowed"])
    assert np.array_equal(o.trace_tshadow(srays, 2, threads=ncpu, times=times)["shadowed"], ref.trace_tshadow(srays, 2, threads=ncpu, times=times)["shadowed"])
    # the oracle's own tree (boxes over all time steps) gives the same answers up to ties
    own = kdo.Oracle(xyz, idx, flags, motion=mo)
    assert np.array_equal(own.bound(), ref.bound())
    c2 = own.trace_closest(rays, threads=ncpu, times=times)
    helpers.check_closest_parity(c2["prim"], c2["t"], c2["u"], c2["v"], r)
    ref.close()
