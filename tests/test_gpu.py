"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of libb200rt.so
and is checked against (a) the golden vectors generated from the unmodified reference, (b) the oracle's C
restatement on the same seeded inputs, (c) the live reference when oracle/_ref travelled with the repo, and
(d) size-independent properties at the full BASELINE.json size (1 M triangles, 16 M rays)."""
import glob
import os
import threading

import numpy as np
import pytest

from libyafaray_b200 import rt, scenes
from oracle import kdo, yref
from tests import helpers

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
NCPU = os.cpu_count() or 1


def make_scene(xyz, idx, flags=None, params=None):
    return helpers.make_rt_scene(rt, xyz, idx, flags, params)


def tie_floor(name):
    return 0.97 if "cube" in name else 0.9999  # coplanar cube / floor faces and shared cube edges are genuine exact-t ties


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors(built, path):
    g = np.load(path)
    s = make_scene(g["xyz"], g["idx"], g["flags"])
    assert np.array_equal(s.bound(), g["bound"])
    h = s.trace_closest(g["closest_rays"])
    ref = dict(prim=g["closest_prim"], t=g["closest_t"], u=g["closest_u"], v=g["closest_v"])
    helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], ref, min_agree=tie_floor(path))
    sh = s.trace_shadow(g["shadow_rays"])
    assert np.array_equal((sh != rt.MISS).astype(np.uint8), g["shadow_shadowed"])
    depth = int(g["tshadow_depth"])
    ts = s.trace_tshadow(g["shadow_rays"], depth)
    assert np.array_equal(ts["shadowed"].astype(np.uint8), g["tshadow_shadowed"])
    lit = g["tshadow_shadowed"] == 0
    assert np.allclose(g["tshadow_rgb"][lit, 0], np.power(0.4, ts["n_transparent"][lit].astype(np.float64)), rtol=1e-5)
    assert np.all(ts["occluder"][lit] == rt.MISS)
    s.close()


ZOO = helpers.scene_zoo(small=False)


@pytest.mark.parametrize("name", sorted(ZOO))
def test_zoo_against_oracle(built, name):
    xyz, idx, flags = ZOO[name]
    s = make_scene(xyz, idx, flags)
    o = kdo.Oracle(xyz, idx, flags)
    assert np.array_equal(s.bound(), o.bound())
    closest, shadow = helpers.ray_zoo(s.bound(), n=200000, seed=11, edge_cases=not name.startswith("cube_grid"))
    h = s.trace_closest(closest)
    ref = o.trace_closest(closest, threads=NCPU)
    helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], ref, min_agree=tie_floor(name))
    sh = s.trace_shadow(shadow)
    rs = o.trace_shadow(shadow, threads=NCPU)
    assert np.array_equal((sh != rt.MISS).astype(np.uint8), rs["shadowed"])
    # the occluder is any-hit: it may differ from the oracle's, but it must be a real shadow caster
    occ = sh[sh != rt.MISS]
    assert np.all(occ < idx.shape[0]) and np.all(flags[occ] & scenes.F_SHADOW)
    for depth in (0, 2, 8):
        ts = s.trace_tshadow(shadow, depth)
        rts = o.trace_tshadow(shadow, depth, threads=NCPU, max_list=8)
        assert np.array_equal(ts["shadowed"].astype(np.uint8), rts["shadowed"]), depth
        lit = rts["shadowed"] == 0
        assert np.array_equal(ts["n_transparent"][lit].astype(np.int32), rts["n_transparent"][lit]), depth
        # same SET of distinct transparent casters (order follows the traversal, which is ours)
        got = np.sort(helpers.prim_signed(ts["transparent"]["prim"][lit]), axis=1)
        exp = np.sort(rts["list"][lit].astype(np.int64), axis=1)
        assert np.array_equal(got, exp), depth
    s.close()


@pytest.mark.skipif(not yref.available(), reason="oracle/_ref/libyafref.so did not travel")
@pytest.mark.parametrize("name", ["hf_flags", "objects_flags", "soup", "hf_quads", "spheres_flags", "cube_grid"])
def test_zoo_against_live_reference(built, name):
    xyz, idx, flags = ZOO[name]
    s = make_scene(xyz, idx, flags)
    ref = yref.RefScene(xyz, idx, flags)
    assert np.array_equal(s.bound(), ref.bound())
    closest, shadow = helpers.ray_zoo(s.bound(), n=200000, seed=13, edge_cases=not name.startswith("cube_grid"))
    h = s.trace_closest(closest)
    r = ref.trace_closest(closest, threads=NCPU)
    helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], r, min_agree=tie_floor(name))
    sh = s.trace_shadow(shadow)
    assert np.array_equal((sh != rt.MISS).astype(np.uint8), ref.trace_shadow(shadow, threads=NCPU)["shadowed"])
    ts = s.trace_tshadow(shadow, 3)
    assert np.array_equal(ts["shadowed"].astype(np.uint8), ref.trace_tshadow(shadow, 3, threads=NCPU)["shadowed"])
    ref.close(); s.close()


@pytest.fixture(scope="module")
def big():
    xyz, idx, flags = scenes.heightfield(707)  # S1M-hf: 999 698 triangles (BASELINE.json configs[1])
    s = make_scene(xyz, idx, flags)
    yield s, xyz, idx, flags
    s.close()


def test_full_size_sample_against_oracle(built, big):
    s, xyz, idx, flags = big
    assert s.stats()["n_faces"] == 999698
    o = kdo.Oracle(xyz, idx, flags)
    rays = scenes.rays_incoherent(400000, seed=99)
    h = s.trace_closest(rays)
    ref = o.trace_closest(rays, threads=NCPU)
    helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], ref)
    srays = scenes.rays_shadow(400000, seed=98)
    sh = s.trace_shadow(srays)
    assert np.array_equal((sh != rt.MISS).astype(np.uint8), o.trace_shadow(srays, threads=NCPU)["shadowed"])


def test_full_size_properties(built, big):
    """16 M incoherent rays on 1 M triangles: determinism, shard invariance, closest/shadow consistency."""
    s, xyz, idx, flags = big
    n = 1 << 24
    rays = scenes.rays_incoherent(n, seed=12345)
    h = s.trace_closest(rays)
    prim = helpers.prim_signed(h["prim"])
    hit = prim >= 0
    assert 0.25 < hit.mean() < 0.45
    assert np.all(h["t"][~hit] == 0) and np.all(h["t"][hit] > 0)
    # every hit point lies inside the tree bound
    b = s.bound().astype(np.float64)
    p = rays[hit, 0:3].astype(np.float64) + h["t"][hit, None].astype(np.float64) * rays[hit, 4:7].astype(np.float64)
    assert np.all(p >= b[:3] - 1e-4) and np.all(p <= b[3:] + 1e-4)
    # determinism: a second pass is byte-identical
    h2 = s.trace_closest(rays)
    assert h.tobytes() == h2.tobytes()
    # shard invariance: ragged shards traced separately give the same bytes as the whole batch
    cuts = [0, 1, 777, 5_000_001, n // 2 + 3, n]
    parts = [s.trace_closest(rays[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.concatenate(parts).tobytes() == h.tobytes()
    # closest/shadow consistency (all faces are visible shadow casters here, tmin = 0):
    # shadowed within T  <=>  the closest hit is nearer than T
    T = 0.25
    srays = rays.copy()
    srays[:, 7] = T
    sh = s.trace_shadow(srays) != rt.MISS
    assert np.array_equal(sh, hit & (h["t"] < T))


def test_host_and_device_entry_points_agree(built, big):
    import torch
    s = big[0]
    rays = scenes.rays_incoherent(1 << 20, seed=4)
    h = s.trace_closest(rays)
    d_rays = torch.from_numpy(rays).cuda()
    d_out = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    s.trace_closest_device(d_rays.data_ptr(), rays.shape[0], d_out.data_ptr(), st)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == h.tobytes()
    srays = scenes.rays_shadow(1 << 20, seed=5)
    sh = s.trace_shadow(srays)
    d_s = torch.from_numpy(srays).cuda()
    d_o = torch.empty(srays.shape[0], dtype=torch.int32, device="cuda")
    s.trace_shadow_device(d_s.data_ptr(), srays.shape[0], d_o.data_ptr(), st)
    torch.cuda.synchronize()
    assert d_o.cpu().numpy().view(np.uint32).tobytes() == sh.tobytes()
    # pinned host buffers take the direct-copy path and must give the same bytes
    pin_in = rt.PinnedBuffer((rays.shape[0], 8), np.float32)
    pin_out = rt.PinnedBuffer((rays.shape[0],), rt.HIT_DTYPE)
    pin_in.array[:] = rays
    s.trace_closest(pin_in.array, out=pin_out.array)
    assert pin_out.array.tobytes() == h.tobytes()
    pin_in.free(); pin_out.free()


def test_concurrent_host_threads(built, big):
    """Render workers call the accelerator concurrently (integrator_tiled.cc:232-248)."""
    s = big[0]
    batches = [scenes.rays_incoherent(300000 + 1000 * k, seed=50 + k) for k in range(6)]
    expect = [s.trace_closest(b) for b in batches]
    got = [None] * len(batches)
    errors = []

    def work(k):
        try:
            for _ in range(3):
                got[k] = s.trace_closest(batches[k])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(batches))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors
    for a, b in zip(expect, got):
        assert a.tobytes() == b.tobytes()


def test_single_ray_calls(built):
    """The per-ray compatibility path: batches of one."""
    xyz, idx, flags = scenes.cube_scene()
    s = make_scene(xyz, idx, flags)
    o = kdo.Oracle(xyz, idx, flags)
    rays = scenes.rays_camera(16, 9, eye=(8, -9, 6), look=(0, 0, 0.5))
    ref = o.brute_closest(rays)
    for i in range(rays.shape[0]):
        h = s.trace_closest(rays[i:i + 1])
        assert h["t"][0] == ref["t"][i] or abs(h["t"][0] - ref["t"][i]) <= 1e-5 * abs(ref["t"][i])
    assert s.trace_closest(rays[:0]).shape == (0,)
    s.close()


def test_update_face_flags_and_multi_mesh(built):
    xyz, idx, flags = scenes.objects(20000, n_spheres=8)
    new_flags = helpers.flag_mix(idx.shape[0], seed=5)
    a = make_scene(xyz, idx, flags)
    a.update_face_flags(new_flags)
    b = make_scene(xyz, idx, new_flags)
    closest, shadow = helpers.ray_zoo(a.bound(), n=100000, seed=8)
    assert a.trace_closest(closest).tobytes() == b.trace_closest(closest).tobytes()
    assert np.array_equal(a.trace_shadow(shadow) != rt.MISS, b.trace_shadow(shadow) != rt.MISS)
    # two uploads == one upload of the concatenation; face ids continue across meshes
    half = idx.shape[0] // 2
    c = rt.Scene(0)
    c.add_mesh(xyz, idx[:half], new_flags[:half])
    c.add_mesh(xyz, idx[half:], new_flags[half:])
    c.build()
    assert c.trace_closest(closest).tobytes() == b.trace_closest(closest).tobytes()
    a.close(); b.close(); c.close()


def test_empty_scene_and_errors(built):
    s = rt.Scene(0)
    with pytest.raises(rt.B200RTError):
        s.trace_closest(scenes.rays_incoherent(10))  # not built yet
    s.build()  # empty scene: zero bound, every ray misses (accelerator_kdtree_original.cc:96)
    assert np.array_equal(s.bound(), np.zeros(6, np.float32))
    h = s.trace_closest(scenes.rays_incoherent(1000))
    assert np.all(h["prim"] == rt.MISS) and np.all(h["t"] == 0)
    assert np.all(s.trace_shadow(scenes.rays_shadow(1000)) == rt.MISS)
    with pytest.raises(rt.B200RTError):
        s.trace_tshadow(scenes.rays_shadow(10), rt.TSHADOW_MAX + 1)
    with pytest.raises(rt.B200RTError):
        rt.Scene(99)
    s.close()


def test_build_parameters_do_not_change_hits(built):
    """Different trees, same answers (up to ties): the SURVEY 8a argument, checked on the GPU path."""
    xyz, idx, flags = scenes.objects(60000, n_spheres=16)
    rays = scenes.rays_incoherent(300000, seed=21)
    base = None
    for p in (None, rt.make_params(max_leaf_size=1), rt.make_params(max_leaf_size=8, cost_ratio=2.5, empty_bonus=0.1), rt.make_params(depth=12)):
        s = make_scene(xyz, idx, flags, p)
        h = s.trace_closest(rays)
        if base is None:
            base = h
        else:
            ref = dict(prim=helpers.prim_signed(base["prim"]), t=base["t"], u=base["u"], v=base["v"])
            helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], ref)
        s.close()


def test_restart_path_with_tiny_stack(built):
    """libb200rt_stack2.so is the same source built with a 2-entry short stack: nearly every ray overflows the ring
    and replays its descent (the exact kd-restart of kd_kernels.cuh), on a mixed scene and on the thin-slab cube grid.  Results must
    be byte-identical to the regular build's -- face ids included, i.e. the replay preserves the visiting order."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stress = os.path.join(root, "libyafaray_b200", "libb200rt_stack2.so")
    assert os.path.exists(stress), "build() must produce libb200rt_stack2.so"
    code = (
        "import sys, hashlib, numpy as np; sys.path.insert(0, %r)\n"
        "from libyafaray_b200 import rt, scenes\n"
        "from tests import helpers\n"
        "xyz, idx, fl = scenes.objects(150000, n_spheres=24); fl = helpers.flag_mix(idx.shape[0], 3)\n"
        "s = rt.Scene(0); s.add_mesh(xyz, idx, fl); s.build()\n"
        "c, sh = helpers.ray_zoo(s.bound(), n=300000, seed=17)\n"
        "h = hashlib.md5()\n"
        "h.update(s.trace_closest(c).tobytes()); h.update((s.trace_shadow(sh) != rt.MISS).tobytes())\n"
        "h.update(s.trace_tshadow(sh, 2)['shadowed'].tobytes())\n"
        "xyz, idx, fl = scenes.cube_grid(9)\n"  # 1-ulp slabs: zero-length intervals meet the replay
        "s2 = rt.Scene(0); s2.add_mesh(xyz, idx, fl); s2.build()\n"
        "c, sh = helpers.ray_zoo(s2.bound(), n=300000, seed=18, edge_cases=False)\n"
        "h.update(s2.trace_closest(c).tobytes()); h.update((s2.trace_shadow(sh) != rt.MISS).tobytes()); print(h.hexdigest())\n" % root)
    digests = []
    for lib in ("", stress):
        env = dict(os.environ)
        env.pop("B200RT_LIB", None)
        if lib:
            env["B200RT_LIB"] = lib
        out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append(out.stdout.strip().splitlines()[-1])
    assert digests[0] == digests[1]


def test_tree_space_rays_equal_wrapped_rays(built):
    """B200RT_RAYS_TREE_SPACE: feeding the rays the reference's wrappers would hand to the virtual queries
    (accelerator.h:103-120: origin += dir * tmin, t_max = tmax - 2 tmin) gives the same answers as letting the
    library apply the wrappers."""
    xyz, idx, _ = scenes.objects(40000, n_spheres=12)
    flags = helpers.flag_mix(idx.shape[0], seed=9)
    s = make_scene(xyz, idx, flags)
    closest, shadow = helpers.ray_zoo(s.bound(), n=100000, seed=31)
    pre = shadow.copy()
    tmin = shadow[:, 3:4]
    pre[:, 0:3] = shadow[:, 0:3] + shadow[:, 4:7] * tmin            # float32 multiply then add, as the wrapper does
    bounded = shadow[:, 7] >= 0
    pre[bounded, 7] = shadow[bounded, 7] - np.float32(2) * shadow[bounded, 3]
    a = s.trace_shadow(shadow)
    b = s.trace(rt.QUERY_SHADOW, pre, flags=rt.RAYS_TREE_SPACE)
    assert np.array_equal(a != rt.MISS, b != rt.MISS)
    ta = s.trace_tshadow(shadow, 3)
    tb = s.trace(rt.QUERY_TSHADOW, pre, flags=rt.RAYS_TREE_SPACE, max_depth=3)
    assert np.array_equal(ta["shadowed"], tb["shadowed"]) and np.array_equal(ta["n_transparent"], tb["n_transparent"])
    assert s.trace(rt.QUERY_CLOSEST, closest, flags=rt.RAYS_TREE_SPACE).tobytes() == s.trace_closest(closest).tobytes()
    with pytest.raises(rt.B200RTError):
        s.trace(7, closest)
    s.close()


@pytest.mark.parametrize("split", [False, True])
def test_trace_jobs_bundles_match_single_queries(built, split):
    """b200rt_trace_jobs[_begin/_end]: what one flush of the renderer's wavefront ray queue hands over -- closest, shadow and
    transparent-shadow batches in pinned memory, traced in place by ONE mixed-kind launch per scene (traceMixedKernel), plus
    a second scene and an unpinned job in the same call -- must be byte-identical to the single-kind queries."""
    xyz, idx, _ = scenes.objects(30000, n_spheres=10)
    flags = helpers.flag_mix(idx.shape[0], seed=4)
    s1 = make_scene(xyz, idx, flags)
    s2 = make_scene(*scenes.cube_scene())
    launches0 = rt.launch_count()
    for n in (1, 31, 700, 5000):
        closest, shadow = helpers.ray_zoo(s1.bound(), n=n, seed=50 + n)
        closest, shadow = closest[:n], shadow[:n]
        c2, _ = helpers.ray_zoo(s2.bound(), n=max(n // 2, 1), seed=60 + n)
        c2 = c2[: max(n // 2, 1)]
        bufs = {}
        def pin(name, src_or_shape, dtype=None):
            if dtype is None:
                b = rt.PinnedBuffer(src_or_shape.shape, np.float32); b.array[:] = src_or_shape
            else:
                b = rt.PinnedBuffer(src_or_shape, dtype)
            bufs[name] = b
            return b.array
        fl = rt.RAYS_TREE_SPACE | rt.BUFFERS_PINNED
        out_unpinned = np.empty(shadow.shape[0], np.uint32)
        jobs = [
            (s1, rt.QUERY_CLOSEST, fl, pin("rc", closest), pin("oc", (n,), rt.HIT_DTYPE), 0),
            (s1, rt.QUERY_SHADOW, fl, pin("rs", shadow), pin("os", (n,), np.uint32), 0),
            (s1, rt.QUERY_TSHADOW, fl, pin("rt", shadow), pin("ot", (n,), rt.TSHADOW_DTYPE), 3),
            (s1, rt.QUERY_TSHADOW, fl, pin("rt2", shadow), pin("ot2", (n,), rt.TSHADOW_DTYPE), 1),   # same scene and kind again: its own launch
            (s2, rt.QUERY_CLOSEST, fl, pin("rc2", c2), pin("oc2", (c2.shape[0],), rt.HIT_DTYPE), 0),
            (s1, rt.QUERY_SHADOW, rt.RAYS_TREE_SPACE, np.ascontiguousarray(shadow), out_unpinned, 0),  # pageable memory: staged path
        ]
        rt.trace_jobs(jobs, split=split)
        assert bufs["oc"].array.tobytes() == s1.trace(rt.QUERY_CLOSEST, closest, flags=rt.RAYS_TREE_SPACE).tobytes()
        assert bufs["os"].array.tobytes() == s1.trace(rt.QUERY_SHADOW, shadow, flags=rt.RAYS_TREE_SPACE).tobytes()
        for name, depth in (("ot", 3), ("ot2", 1)):
            ref = s1.trace(rt.QUERY_TSHADOW, shadow, flags=rt.RAYS_TREE_SPACE, max_depth=depth)
            got = bufs[name].array
            assert np.array_equal(got["shadowed"], ref["shadowed"]) and np.array_equal(got["n_transparent"], ref["n_transparent"])
        assert bufs["oc2"].array.tobytes() == s2.trace(rt.QUERY_CLOSEST, c2, flags=rt.RAYS_TREE_SPACE).tobytes()
        assert np.array_equal(out_unpinned, bufs["os"].array)
        for b in bufs.values():
            b.free()
    assert rt.launch_count() > launches0
    with pytest.raises(rt.B200RTError):
        rt.trace_jobs([(s1, 9, 0, np.zeros((4, 8), np.float32), np.zeros(4, np.uint32), 0)])
    s1.close(); s2.close()


def test_flush_combiner_merges_jobs_of_several_threads(built, monkeypatch):
    """B200RT_COMBINE=1 (off by default): the in-place jobs of several threads' b200rt_trace_jobs calls leave in
    shared launches of traceSegmentsKernel (kd_segments.cuh).  Eight threads, each tracing its own pinned closest / shadow /
    transparent-shadow batches in a loop: every answer byte-identical to the single-kind queries, fewer launches than jobs."""
    import threading
    monkeypatch.setenv("B200RT_COMBINE", "1")
    monkeypatch.setenv("B200RT_COMBINE_RAYS", "1500")
    xyz, idx, _ = scenes.objects(30000, n_spheres=0)
    flags = helpers.flag_mix(idx.shape[0], seed=4)
    s = make_scene(xyz, idx, flags)           # the combiner is made by b200rt_build, with the environment as it is now
    monkeypatch.delenv("B200RT_COMBINE")
    n_threads, rounds, n = 8, 25, 300
    fl = rt.RAYS_TREE_SPACE | rt.BUFFERS_PINNED
    work, expect = [], []
    for t in range(n_threads):
        closest, shadow = helpers.ray_zoo(s.bound(), n=n, seed=900 + t)
        closest, shadow = closest[:n], shadow[:n]
        bufs = [rt.PinnedBuffer(closest.shape, np.float32), rt.PinnedBuffer((n,), rt.HIT_DTYPE), rt.PinnedBuffer(shadow.shape, np.float32), rt.PinnedBuffer((n,), np.uint32),
                rt.PinnedBuffer(shadow.shape, np.float32), rt.PinnedBuffer((n,), rt.TSHADOW_DTYPE)]
        bufs[0].array[:] = closest; bufs[2].array[:] = shadow; bufs[4].array[:] = shadow
        work.append(bufs)
    failures = []
    launches0 = rt.launch_count()
    def worker(t):
        b = work[t]
        try:
            for _ in range(rounds):
                b[1].array[:] = 0; b[3].array[:] = 0
                rt.trace_jobs([(s, rt.QUERY_CLOSEST, fl, b[0].array, b[1].array, 0), (s, rt.QUERY_SHADOW, fl, b[2].array, b[3].array, 0),
                               (s, rt.QUERY_TSHADOW, fl, b[4].array, b[5].array, 2)], split=True)
        except Exception as e:  # noqa: BLE001
            failures.append(repr(e))
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads: th.start()
    for th in threads: th.join()
    launches = rt.launch_count() - launches0
    assert not failures, failures
    for t in range(n_threads):
        b = work[t]
        assert b[1].array.tobytes() == s.trace(rt.QUERY_CLOSEST, b[0].array, flags=rt.RAYS_TREE_SPACE).tobytes()
        assert b[3].array.tobytes() == s.trace(rt.QUERY_SHADOW, b[2].array, flags=rt.RAYS_TREE_SPACE).tobytes()
        ref = s.trace(rt.QUERY_TSHADOW, b[4].array, flags=rt.RAYS_TREE_SPACE, max_depth=2)
        assert np.array_equal(b[5].array["shadowed"], ref["shadowed"]) and np.array_equal(b[5].array["n_transparent"], ref["n_transparent"])
    n_jobs = n_threads * rounds * 3
    helpers.report("flush_combiner", jobs=n_jobs, launches=int(launches), rays_per_launch=float(n_jobs * n / max(1, launches)))
    assert 0 < launches < n_jobs, (launches, n_jobs)
    for bufs in work:
        for b in bufs:
            b.free()
    s.close()


REF_SCENES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_scenes", "*.bin")))


@pytest.mark.parametrize("path", REF_SCENES, ids=[os.path.basename(p)[:-4] for p in REF_SCENES])
def test_reference_test_scenes_primary_and_surface_rays(built, path):
    """The geometry of the reference's own tests/test00, test01, test03 clients as AcceleratorB200 extracts it (dumped with
    B200_DUMP_SCENE; axis-aligned boxes and planes -- the kind of scene whose kd slabs have zero thickness in t), traced with
    primary rays and with rays that START ON SURFACES like bounce / shadow rays do.  Checked against the oracle and, when it
    travelled, the live reference.  (tests/test02's 21 MB dump is not committed; scenes.cube_grid reproduces its slabs.)"""
    xyz, idx, flags = helpers.load_scene_dump(path)
    s = make_scene(xyz, idx, flags)
    o = kdo.Oracle(xyz, idx, flags)
    assert np.array_equal(s.bound(), o.bound())
    b = s.bound().astype(np.float64)
    ext = b[3:] - b[:3]
    primary = scenes.rays_incoherent(300000, seed=31, lo=b[:3] - 0.2 * ext, hi=b[3:] + 0.2 * ext)
    checkers = [("oracle", o)] + ([("live reference", yref.RefScene(xyz, idx, flags))] if yref.available() else [])
    for what, ref in checkers:
        r1 = ref.trace_closest(primary, threads=NCPU)
        surface = helpers.surface_rays(primary, r1["t"], seed=32)
        for rays, rr in ((primary, r1), (surface, ref.trace_closest(surface, threads=NCPU))):
            h = s.trace_closest(rays)
            helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], rr, min_agree=0.97)  # shared box edges tie
            assert np.array_equal((s.trace_shadow(rays) != rt.MISS).astype(np.uint8), ref.trace_shadow(rays, threads=NCPU)["shadowed"]), what
    s.close()


def _adversarial_scenes():
    out = {}
    x, i, f = scenes.objects(20000, n_spheres=10, seed=41)
    out["far_from_origin"] = ((x + np.array([1000.0, 2000.0, -500.0], np.float32)).astype(np.float32), i, f)  # coarse float grid: 6e-5 ulps
    out["cube_grid_far"] = helpers.cube_grid_far()                                                               # thin slabs AND coarse t
    x, i, f = scenes.heightfield(60)
    flat = x.copy(); flat[:, 2] = 0.25
    out["flat_with_duplicates"] = (np.concatenate([flat, flat]), np.concatenate([i, i + np.uint32(np.where(i == scenes.TRI, 0, flat.shape[0]))]).astype(np.uint32),
                                   np.concatenate([f, f]))                                                        # zero-extent bound on z, every face twice
    rng = np.random.default_rng(43)
    n = 4000
    a = rng.random((n, 3), dtype=np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    e = rng.normal(size=(n, 3)).astype(np.float32) * 1e-4
    sl = np.stack([a - d, a + d, a + e], axis=1).reshape(-1, 3).astype(np.float32)                                # needle triangles crossing the whole scene
    x, i, f = scenes.soup(20000, seed=44)
    si = np.concatenate([np.arange(3 * n, dtype=np.uint32).reshape(n, 3) + np.uint32(x.shape[0]), np.full((n, 1), scenes.TRI, np.uint32)], axis=1)
    out["needles_in_soup"] = (np.concatenate([x, sl]), np.concatenate([i, si]), np.concatenate([f, np.full(n, 3, np.uint8)]))
    return out


@pytest.mark.parametrize("name", ["far_from_origin", "cube_grid_far", "flat_with_duplicates", "needles_in_soup"])
def test_adversarial_scenes(built, name):
    """Geometry chosen to break a t-interval kd traversal: coordinates far from the origin (coarse float grid, slabs below the
    resolution of t), a flat scene (zero-extent tree bound on one axis) with every face duplicated (exact ties everywhere),
    needle triangles that cross the whole scene (referenced by thousands of leaves).  Primary and on-surface rays against the oracle."""
    xyz, idx, flags = _adversarial_scenes()[name]
    s = make_scene(xyz, idx, flags)
    o = kdo.Oracle(xyz, idx, flags)
    assert np.array_equal(s.bound(), o.bound())
    b = s.bound().astype(np.float64)
    ext = np.maximum(b[3:] - b[:3], 1e-3)
    primary = scenes.rays_incoherent(300000, seed=51, lo=b[:3] - 0.3 * ext, hi=b[3:] + 0.3 * ext)
    r1 = o.trace_closest(primary, threads=NCPU)
    surface = helpers.surface_rays(primary, r1["t"], seed=52, tmin=0.0005 * float(np.linalg.norm(ext)))
    floor = 0.45 if name == "flat_with_duplicates" else 0.9999  # duplicates: either copy may win
    for rays, rr in ((primary, r1), (surface, o.trace_closest(surface, threads=NCPU))):
        h = s.trace_closest(rays)
        shadow_differs = (s.trace_shadow(rays) != rt.MISS).astype(np.uint8) != o.trace_shadow(rays, threads=NCPU)["shadowed"]
        if name == "cube_grid_far":
            # At coordinates of 300 (ulp 3e-5) with faces planar up to one ulp no kd traversal is exact any more: the UNMODIFIED
            # reference on its own tree differs from a brute-force test of all primitives on ~1e-4 of these rays, hit/miss
            # included (pinned without any code of ours by tests/test_oracle.py::test_reference_itself_is_not_exact_on_cube_grid_far).
            # Two bars instead of the tie classification: (1) against brute force the kernel may be wrong no more often than
            # the reference is (2.5e-4, same bound as in that test); (2) on the SAME tree the kernel must agree with the
            # reference traversal restated over it (below, test_same_tree_exactness) .
            prim = helpers.prim_signed(h["prim"])
            brute = o.brute_closest(rays, threads=NCPU)
            agree, n_diff, n_bad = helpers.parity_counts(prim, h["t"], brute)
            helpers.report("cube_grid_far_vs_brute", rays=int(rays.shape[0]), ids_equal=agree, id_mismatches=n_diff, non_tie=n_bad,
                           shadow_differs_vs_other_tree=int(shadow_differs.sum()))
            same = prim == brute["prim"]
            assert np.array_equal(h["t"][same], brute["t"][same])
            assert n_bad <= 2.5e-4 * rays.shape[0] and shadow_differs.mean() <= 2.5e-4, (n_diff, n_bad, int(shadow_differs.sum()))
            assert agree >= 0.97
            continue
        helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], rr, min_agree=floor)
        assert not shadow_differs.any()
    s.close()


def test_rays_lying_in_split_planes(built):
    """Rays whose direction component is exactly zero and whose origin lies exactly on a kd split plane (axis-parallel light
    and camera rays in axis-aligned scenes): the kernel must give what the REFERENCE TRAVERSAL gives on the same tree
    (exported host tree -> oracle), hit/miss included.  Against another valid tree even the reference differs on such rays."""
    xyz, idx, flags = scenes.cube_grid(7)
    tree = rt.host_tree(xyz, idx)
    a, b = tree["a"], tree["b"]
    split = a.view(np.float32)
    interior = np.nonzero((b & 3) != 3)[0]
    rng = np.random.default_rng(2)
    lo, ext = tree["bound"][:3], tree["bound"][3:] - tree["bound"][:3]
    rays = np.zeros((20000, 8), np.float32)
    for k, node in enumerate(rng.choice(interior, size=rays.shape[0])):
        axis = int(b[node]) & 3
        o = (lo - 0.3 * ext + rng.random(3).astype(np.float32) * 1.6 * ext).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        o[axis], d[axis] = split[node], 0.0
        rays[k] = [o[0], o[1], o[2], 0.0, d[0], d[1], d[2], -1.0]
    s = make_scene(xyz, idx, flags)
    same_tree = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(tree), bound=tree["bound"])
    ref = same_tree.trace_closest(rays, threads=NCPU)
    h = s.trace_closest(rays)
    assert (ref["prim"] >= 0).mean() > 0.2
    helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], ref, min_agree=0.97)
    assert np.array_equal((s.trace_shadow(rays) != rt.MISS).astype(np.uint8), same_tree.trace_shadow(rays, threads=NCPU)["shadowed"])
    s.close()


SAME_TREE = dict(ZOO)
SAME_TREE["cube_grid_far"] = helpers.cube_grid_far()


@pytest.mark.parametrize("name", sorted(SAME_TREE))
def test_same_tree_exactness(built, name):
    """The kernel against the REFERENCE TRAVERSAL restated over the very tree the kernel walks (b200rt_host_tree_export ->
    oracle), on the whole zoo and WITH the edge-case rays (axis-parallel, exact diagonals, origins on the bound) that the
    cross-tree tests leave out on the cube grids: with the tree taken out of the comparison, hit/miss and t must agree on
    every ray -- what orthographic cameras and directional lights shoot through axis-aligned geometry included -- and
    ids may differ only where two primitives are hit at exactly the same t in different leaves or in another order (the kernel
    orders leaves by t-interval, the reference by stored entry/exit points).  The counts go to gpurun_out/parity_report.jsonl."""
    xyz, idx, flags = SAME_TREE[name]
    is_sphere = idx[:, 2] == scenes.SPHERE
    s = make_scene(xyz, idx, flags)
    tree = rt.host_tree(xyz, idx)
    o = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(tree), bound=tree["bound"])
    assert np.array_equal(s.bound(), o.bound())
    closest, shadow = helpers.ray_zoo(s.bound(), n=200000, seed=23, edge_cases=True)
    n_edge = scenes.rays_edge_cases(s.bound()[:3], s.bound()[3:], seed=1).shape[0]
    h = s.trace_closest(closest)
    ref = o.trace_closest(closest, threads=NCPU)
    prim = helpers.prim_signed(h["prim"])
    agree, n_diff, n_bad = helpers.parity_counts(prim, h["t"], ref, rtol=0.0)   # rtol 0: a "tie" must be the same t exactly
    e_agree, e_diff, e_bad = helpers.parity_counts(prim[-n_edge:], h["t"][-n_edge:], {k: v[-n_edge:] for k, v in ref.items() if k in ("prim", "t")}, rtol=0.0)
    sh = (s.trace_shadow(shadow) != rt.MISS).astype(np.uint8)
    rs = o.trace_shadow(shadow, threads=NCPU)["shadowed"]
    helpers.report("same_tree", scene=name, rays=int(closest.shape[0]), ids_equal=agree, id_mismatches=n_diff, not_exact_t_ties=n_bad,
                   edge_rays=int(n_edge), edge_id_mismatches=e_diff, edge_not_exact=e_bad, shadow_bool_mismatches=int((sh != rs).sum()),
                   spheres=bool(is_sphere.any()))
    same = prim == ref["prim"]
    assert np.array_equal(h["t"][same], ref["t"][same]) and np.array_equal(h["u"][same], ref["u"][same]) and np.array_equal(h["v"][same], ref["v"][same])
    if n_bad:
        # Where the kernel and the reference traversal disagree on the SAME tree beyond an exact-t tie, the kernel must be the one
        # that is right: its answer has to be what a brute-force test of every primitive gives.  (Seen on cube_grid_far only: the
        # reference's stored entry/exit points lose a 1-ulp slab at coordinates of 300, the kernel's t-intervals do not; the
        # unmodified reference shows the same leaks on its own tree, tests/test_oracle.py::test_reference_itself_is_not_exact_on_cube_grid_far.)
        bad = np.flatnonzero((prim != ref["prim"]) & ((h["t"] != ref["t"]) | ((prim >= 0) != (ref["prim"] >= 0))))
        brute = o.brute_closest(closest[bad], threads=NCPU)
        assert np.array_equal(prim[bad], brute["prim"]) and np.array_equal(h["t"][bad], brute["t"]), f"{n_bad} rays differ from the reference traversal on the same tree and from brute force"
        assert n_bad <= 2.5e-4 * closest.shape[0] and name == "cube_grid_far", (name, n_bad)
    assert agree >= (0.995 if "cube" in name else 0.99999), (agree, n_diff)
    assert np.array_equal(sh, rs)
    s.close()


def _sample_parity_on_exported_tree(xyz, idx, flags, n_rays, seed, label):
    s = make_scene(xyz, idx, flags)
    tree = rt.host_tree(xyz, idx)
    o = kdo.Oracle(xyz, idx, flags, tree=helpers.host_tree_as_oracle_tree(tree), bound=tree["bound"])
    b = s.bound()
    rays = scenes.rays_incoherent(n_rays, seed=seed, lo=b[:3], hi=b[3:])
    diag = float(np.linalg.norm(b[3:].astype(np.float64) - b[:3]))
    srays = scenes.rays_shadow(n_rays, seed=seed + 1, lo=b[:3], hi=b[3:], t_max=0.25 * diag)
    h = s.trace_closest(rays)
    ref = o.trace_closest(rays, threads=NCPU)
    prim = helpers.prim_signed(h["prim"])
    agree, n_diff, n_bad = helpers.parity_counts(prim, h["t"], ref)
    sh = (s.trace_shadow(srays) != rt.MISS).astype(np.uint8)
    rs = o.trace_shadow(srays, threads=NCPU)["shadowed"]
    helpers.report("regime_sample", scene=label, faces=int(idx.shape[0]), rays=int(n_rays), ids_equal=agree, id_mismatches=n_diff, non_tie=n_bad,
                   shadow_bool_mismatches=int((sh != rs).sum()), hit_fraction=float((prim >= 0).mean()), tree_nodes=int(tree["a"].shape[0]))
    helpers.check_closest_parity(prim, h["t"], h["u"], h["v"], ref)
    assert np.array_equal(sh, rs)
    return s, rays, h


def test_s10m_sample_against_oracle(built):
    """BASELINE.json configs[3]'s geometry (10 M-triangle object scene, 2 GB on the device, HBM-resident): a 400 k-ray sample
    against the reference traversal restated over the exported tree (a second 10 M-triangle build is not needed that way)."""
    xyz, idx, flags = scenes.objects(10_000_000)
    s, rays, h = _sample_parity_on_exported_tree(xyz, idx, flags, 400000, 71, "S10M-objects")
    assert s.stats()["n_faces"] > 9_900_000
    s.close()


def test_soup_1m_sample_against_oracle_and_live_reference(built):
    """The worst case for a kd-tree: 1 M random triangles (177 interior + 44 leaf visits and 34 primitive tests per ray)."""
    xyz, idx, flags = scenes.soup(1_000_000)
    s, rays, h = _sample_parity_on_exported_tree(xyz, idx, flags, 400000, 73, "S1M-soup")
    if yref.available():
        ref = yref.RefScene(xyz, idx, flags)
        r = ref.trace_closest(rays[:200000], threads=NCPU)
        helpers.check_closest_parity(helpers.prim_signed(h["prim"][:200000]), h["t"][:200000], h["u"][:200000], h["v"][:200000], r)
        ref.close()
    s.close()


# ---------------------------------------------------------------------------------------------- motion blur
MOTION_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "motion", "motion.npz")


def test_motion_blur_golden_vectors(built):
    """Bezier motion-blur faces + faces of moving instances + static faces at per-ray times, against what the unmodified reference
    returned (tests/golden/motion/motion.npz): ids up to ties, t/u/v bit-identical, shadow and transparent-shadow booleans."""
    g = np.load(MOTION_GOLDEN)
    mo = {k[len("motion_"):]: g[k] for k in g.files if k.startswith("motion_")}
    s = helpers.make_rt_motion_scene(rt, g["xyz"], g["idx"], g["flags"], mo)
    st = s.stats()
    assert st["n_bezier_faces"] == int((mo["kind"] == 1).sum()) and st["n_moving_faces"] == int((mo["kind"] == 2).sum())
    assert np.array_equal(s.bound(), g["bound"])
    h = s.trace(rt.QUERY_CLOSEST, g["closest_rays"], times=g["closest_times"])
    ref = dict(prim=g["closest_prim"], t=g["closest_t"], u=g["closest_u"], v=g["closest_v"])
    helpers.check_closest_parity(helpers.prim_signed(h["prim"]), h["t"], h["u"], h["v"], ref)
    sh = s.trace(rt.QUERY_SHADOW, g["shadow_rays"], times=g["shadow_times"])
    assert np.array_equal((sh != rt.MISS).astype(np.uint8), g["shadow_shadowed"])
    ts = s.trace(rt.QUERY_TSHADOW, g["shadow_rays"], max_depth=int(g["tshadow_depth"]), times=g["shadow_times"])
    assert np.array_equal(ts["shadowed"].astype(np.uint8), g["tshadow_shadowed"])
    # without times every ray is traced at time 0 (what the untimed entry points do)
    h0 = s.trace_closest(g["closest_rays"])
    assert h0.tobytes() == s.trace(rt.QUERY_CLOSEST, g["closest_rays"], times=np.zeros_like(g["closest_times"])).tobytes()
    assert h0.tobytes() != h.tobytes()
    s.close()


def test_motion_blur_against_oracle_and_live_reference(built):
    """A larger moving scene; batches big enough for the two-pass path (the ray time travels through the ray queue), the
    device entry point, and a jobs bundle with times."""
    import torch
    xyz, idx, _, mo = scenes.motion_scene(n_static=30000, n_bezier=20000, n_moving=12000, seed=21)
    flags = helpers.flag_mix(idx.shape[0], seed=22)
    s = helpers.make_rt_motion_scene(rt, xyz, idx, flags, mo)
    o = kdo.Oracle(xyz, idx, flags, motion=mo)
    assert np.array_equal(s.bound(), o.bound())
    b = s.bound()
    rays = scenes.rays_incoherent(300000, seed=23, lo=b[:3], hi=b[3:])
    times = scenes.ray_times(rays.shape[0], 24)
    h = s.trace(rt.QUERY_CLOSEST, rays, times=times)
    prim = helpers.prim_signed(h["prim"])
    ref = o.trace_closest(rays, threads=NCPU, times=times)
    helpers.check_closest_parity(prim, h["t"], h["u"], h["v"], ref)
    hit = prim >= 0
    assert set(np.unique(mo["kind"][prim[hit]])) == {0, 1, 2}
    srays = scenes.rays_shadow(300000, seed=25, lo=b[:3], hi=b[3:], t_max=0.4)
    sh = s.trace(rt.QUERY_SHADOW, srays, times=times)
    assert np.array_equal((sh != rt.MISS).astype(np.uint8), o.trace_shadow(srays, threads=NCPU, times=times)["shadowed"])
    if yref.available():
        live = yref.RefScene(xyz, idx, flags, motion=mo)
        assert np.array_equal(s.bound(), live.bound())
        helpers.check_closest_parity(prim, h["t"], h["u"], h["v"], live.trace_closest(rays, threads=NCPU, times=times))
        assert np.array_equal((sh != rt.MISS).astype(np.uint8), live.trace_shadow(srays, threads=NCPU, times=times)["shadowed"])
        live.close()
    # small batches take the single-kernel path: same bytes as the slice of the big batch
    assert s.trace(rt.QUERY_CLOSEST, rays[:1000], times=times[:1000]).tobytes() == h[:1000].tobytes()
    # device entry point
    d_r, d_t = torch.from_numpy(rays).cuda(), torch.from_numpy(times).cuda()
    d_o = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    rt._check(rt.lib().b200rt_trace_timed_device(s._h, rt.QUERY_CLOSEST, 0, d_r.data_ptr(), d_t.data_ptr(), rays.shape[0], d_o.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert d_o.cpu().numpy().tobytes() == h.tobytes()
    # a flush of the renderer's queue with ray times (pinned, in place, mixed kinds)
    n = 3000
    pr = rt.PinnedBuffer((n, 8), np.float32); pr.array[:] = rays[:n]
    ps = rt.PinnedBuffer((n, 8), np.float32); ps.array[:] = srays[:n]
    pt = rt.PinnedBuffer((n,), np.float32); pt.array[:] = times[:n]
    oc = rt.PinnedBuffer((n,), rt.HIT_DTYPE); os_ = rt.PinnedBuffer((n,), np.uint32)
    fl = rt.BUFFERS_PINNED
    rt.trace_jobs([(s, rt.QUERY_CLOSEST, fl, pr.array, oc.array, 0, pt.array), (s, rt.QUERY_SHADOW, fl, ps.array, os_.array, 0, pt.array)])
    assert oc.array.tobytes() == h[:n].tobytes() and np.array_equal(os_.array, sh[:n])
    for buf in (pr, ps, pt, oc, os_):
        buf.free()
    s.close()


def test_transparent_shadows_deeper_than_the_fixed_record(built):
    """A "shadow_depth" above the 8 casters b200rt_tshadow holds (the reference has no limit, accelerator.h:147-169):
    b200rt_trace_tshadow_deep with records of `capacity` casters, on 24 stacked transparent sheets over an opaque floor."""
    n_sheets = 24
    vs, fs = [], []
    for k in range(n_sheets + 1):
        z = 0.1 + 0.03 * k
        off = 4 * k
        vs.append(np.array([[0, 0, z], [1, 0, z], [1, 1, z], [0, 1, z]], np.float32) + np.float32(0.001 * k))
        fs.append([off, off + 1, off + 2, off + 3])
    xyz, idx = np.concatenate(vs), np.array(fs, np.uint32)
    flags = np.full(idx.shape[0], scenes.F_NORMAL | scenes.F_TRANSPARENT, np.uint8)
    flags[0] = scenes.F_NORMAL  # the lowest sheet is opaque
    s = make_scene(xyz, idx, flags)
    o = kdo.Oracle(xyz, idx, flags)
    rng = np.random.default_rng(3)
    n = 20000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:2] = rng.random((n, 2)) * 0.8 + 0.1
    rays[:, 2] = rng.random(n) * 1.2          # start between, above or below the sheets
    rays[:, 3] = 0.0005
    rays[:, 4:6] = rng.normal(scale=0.2, size=(n, 2))
    rays[:, 6] = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    rays[:, 7] = -1.0
    for depth in (3, 8, 12, 24, 40):
        ref = o.trace_tshadow(rays, depth, threads=NCPU, max_list=max(depth, 1))
        got = s.trace_tshadow_deep(rays, depth)
        assert np.array_equal(got["shadowed"].astype(np.uint8), ref["shadowed"]), depth
        lit = ref["shadowed"] == 0
        assert np.array_equal(got["n_transparent"][lit].astype(np.int32), ref["n_transparent"][lit]), depth
        a = np.sort(np.where(np.arange(max(depth, 1))[None, :] < got["n_transparent"][:, None], helpers.prim_signed(got["transparent"]["prim"]), -1), axis=1)
        b = np.sort(ref["list"].astype(np.int64), axis=1)
        assert np.array_equal(a[lit], b[lit]), depth
        if depth <= rt.TSHADOW_MAX:  # the fixed-size record gives the same answers
            fixed = s.trace_tshadow(rays, depth)
            assert np.array_equal(fixed["shadowed"], got["shadowed"]) and np.array_equal(fixed["n_transparent"], got["n_transparent"])
    assert (ref["n_transparent"][lit] > rt.TSHADOW_MAX).any(), "no ray passed more than 8 transparent sheets: the test would prove nothing"
    with pytest.raises(rt.B200RTError):
        s.trace_tshadow_deep(rays[:10], 9, capacity=8)
    s.close()
