"""Photon-map queries (include/b200pm.h): PhotonMap::updateTree / gather / findNearest of the reference
(src/photon/photon.cc:46-72, include/photon/pkdtree.h) on B200.

CPU (`-m "not gpu"`): the C restatement (oracle/pm_oracle.c) against the committed golden vectors (made from the unmodified
reference by tests/golden/pm/make_pm_golden.py) and against the live reference when oracle/_ref is there; the product's host-side
tree builder against both; the kernel's lookup code itself, compiled for the host (tests/native/pm_host_model.cu), against the
oracle; the C ABI's symbols and argument checks.
GPU (`-m gpu`): b200pm_* through the C ABI against the oracle and the golden vectors, bit for bit; properties at full size.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from libyafaray_b200 import pm, rt, scenes
from oracle import pmo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "pm", "pm_*.npz")))
KINDS = ["uniform", "surfaces", "clusters", "lattice"]
DEFAULT_TUNING = (3, 8, 16, 16)  # b200pm.cu Tuning: kernel (3 = default mix), round_steps, smem_k, patience


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_gather_equal(got, want, what):
    """got / want = (idx [n,k], d2 [n,k], n_found [n], radius_out [n]); everything bit for bit, order included."""
    g_idx, g_d2, g_n, g_r = got
    w_idx, w_d2, w_n, w_r = want
    assert np.array_equal(g_n, w_n), f"{what}: n_found differs on {np.count_nonzero(g_n != w_n)} points"
    valid = np.arange(w_idx.shape[1])[None, :] < w_n[:, None]
    assert np.array_equal(g_idx[valid], w_idx[valid]), f"{what}: photon ids / order differ"
    assert np.array_equal(bits(g_d2)[valid], bits(w_d2)[valid]), f"{what}: squared distances differ"
    assert np.array_equal(bits(g_r), bits(w_r)), f"{what}: final radii differ"


def product_result(found, n_found, radius_out):
    return found["photon"], found["dist_square"], n_found, radius_out


# ---------------------------------------------------------------------------------------------------------------- oracle

def test_golden_fixtures_exist():
    assert len(GOLDEN) == 4, GOLDEN


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_golden(built, path):
    g = np.load(path)
    o = pmo.OracleMap(g["pos"], g["dirs"])
    a, b = o.tree()
    assert np.array_equal(a, g["tree_a"]) and np.array_equal(b, g["tree_b"]), "node array differs from the reference's"
    for i, (k, r2) in enumerate(g["gathers"]):
        want = (g[f"g{i}_idx"], g[f"g{i}_d2"], g[f"g{i}_n"], g[f"g{i}_r"])
        assert_gather_equal(o.gather(g["points"], int(k), r2), want, f"gather k={int(k)} r2={r2}")
    assert_gather_equal(o.gather(g["points"], 12, 0.0, g["radii"]), (g["gr_idx"], g["gr_d2"], g["gr_n"], g["gr_r"]), "gather with per-point radii")
    for i, dist in enumerate(g["nearest"]):
        assert np.array_equal(o.nearest(g["points"], g["normals"], dist), g[f"n{i}"]), f"findNearest dist={dist}"


@pytest.mark.skipif(not pmo.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_live_reference(built, kind):
    for n in (1, 2, 3, 7, 1000, 30000):
        pos, dirs = scenes.photon_cloud(kind, n, seed=n + 1)
        points, normals = scenes.gather_points(pos, 2000, seed=n)
        ref, o = pmo.RefMap(pos, dirs, build_threads=4, query_threads=4), pmo.OracleMap(pos, dirs)
        for x, y in zip(ref.tree(), o.tree()):
            assert np.array_equal(x, y), (kind, n)
        for k, r2 in ((1, 1e-3), (5, 1e-2), (100, 0.05), (50, 1e30), (100, 1e-6), (2, 0.0)):
            assert_gather_equal(o.gather(points, k, r2), ref.gather(points, k, r2), f"{kind} n={n} k={k} r2={r2}")
        for dist in (1e-4, 1e-2, 1.0):
            assert np.array_equal(o.nearest(points, normals, dist), ref.nearest(points, normals, dist))


@pytest.mark.parametrize("kind", KINDS)
def test_pruning_never_changes_a_result(built, kind):
    """the split-plane test only skips photons the distance test would reject, so the results are a function
    of the leaf ORDER alone -- the lookup with the pruning switched off gives the same bits."""
    pos, dirs = scenes.photon_cloud(kind, 3000, seed=77)
    points, normals = scenes.gather_points(pos, 400, seed=78, jitter=0.02)
    o = pmo.OracleMap(pos, dirs)
    for k, r2 in ((1, 1e-3), (7, 5e-3), (40, 2e-2), (25, 1e30)):
        want = o.gather(points, k, r2)
        pmo.set_unpruned(True)
        try:
            got = o.gather(points, k, r2)
        finally:
            pmo.set_unpruned(False)
        assert_gather_equal(got, want, f"{kind} k={k} r2={r2}")
    for dist in (1e-3, 1.0):
        want = o.nearest(points, normals, dist)
        pmo.set_unpruned(True)
        try:
            got = o.nearest(points, normals, dist)
        finally:
            pmo.set_unpruned(False)
        assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------------ host side of the product

@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_host_builder_builds_the_reference_tree_golden(built, path):
    g = np.load(path)
    for threads in (1, 3, 0):
        a, b = pm.host_tree(g["pos"], threads)
        assert np.array_equal(a, g["tree_a"]) and np.array_equal(b, g["tree_b"]), threads


@pytest.mark.parametrize("kind", KINDS)
def test_host_builder_builds_the_reference_tree(built, kind):
    for n in (1, 2, 3, 5, 64, 4097, 200_000):
        pos, _ = scenes.photon_cloud(kind, n, seed=3 * n)
        want = pmo.OracleMap(pos).tree()
        for threads in (1, 8):
            a, b = pm.host_tree(pos, threads)
            assert np.array_equal(a, want[0]) and np.array_equal(b, want[1]), (kind, n, threads)
            assert len(a) == 2 * n - 1


def test_library_exports_every_photon_map_symbol(built):
    L = C.CDLL(rt.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "b200pm.h")).read()
    for name in pm.SYMBOLS:
        assert name + "(" in header, name
        assert hasattr(L, name), name
    import re
    declared = set(re.findall(r"\b(b200pm_[a-z_]+)\s*\(", header))
    assert declared == set(pm.SYMBOLS), declared ^ set(pm.SYMBOLS)


def test_lookup_kernels_contain_no_fused_multiply_add(built):
    """Bit parity of the squared distances rests on separate multiplies and adds (the reference is compiled without FMA
    contraction): the SASS of every photon-map kernel in the shipped library must not contain a single FFMA."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", rt.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    kernels, current = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            current = line.split("Function :")[1].strip()
            kernels[current] = 0
        elif current and "FFMA" in line:
            kernels[current] += 1
    pm_kernels = {name: n for name, n in kernels.items() if "pmLookup" in name}
    assert len(pm_kernels) >= 9, sorted(kernels)  # plain <0,1,2> + phased <0,1,2> x <single pop>
    assert all(n == 0 for n in pm_kernels.values()), pm_kernels


def test_argument_checks_without_a_device(built):
    pos = np.zeros((4, 3), np.float32)
    a, b = np.zeros(8, np.uint32), np.zeros(8, np.uint32)
    L = pm.lib()
    assert L.b200pm_host_tree_build(rt._p(pos), 0, 1, rt._p(a), rt._p(b)) == -1       # empty map: the reference logs an error and builds nothing
    assert L.b200pm_host_tree_build(None, 4, 1, rt._p(a), rt._p(b)) == -1
    assert b"photons" in L.b200rt_last_error()
    bad = np.array([[0, 0, 0], [1, np.nan, 0], [2, 2, 2], [3, 3, 3]], np.float32)
    assert L.b200pm_host_tree_build(rt._p(bad), 4, 1, rt._p(a), rt._p(b)) == -1 and b"finite" in L.b200rt_last_error()
    h = C.c_void_p(0)
    assert L.b200pm_create(0, None, None, 4, 1, C.byref(h)) == -1 and not h.value
    assert L.b200pm_gather(None, None, 0, 1, C.c_float(1.0), None, None, None, None) == -1
    if rt.device_count() == 0:
        # no CPU fallback: creating a map without a usable device fails loudly
        assert L.b200pm_create(0, rt._p(pos), None, 4, 1, C.byref(h)) < 0 and not h.value
        with pytest.raises(rt.B200RTError):
            pm.PhotonMap(pos)


# ---------------------------------------------------------------------------- the kernel's lookup code, compiled for the host

@pytest.fixture(scope="module")
def host_model(built):
    out = os.path.join(ROOT, "build", "pm_host_model.so")
    src = os.path.join(ROOT, "tests", "native", "pm_host_model.cu")
    hdr = os.path.join(ROOT, "libyafaray_b200", "csrc", "pm_kernels.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", out, src])
    L = C.CDLL(out)
    L.pm_model_lookup.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_uint32, C.c_uint32, C.c_float] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
    return L


def run_model(L, mode, nodes, dirs4, points, normals, k, r2, radii=None, round_steps=0, single_pop=0):
    n = len(points)
    found = np.zeros((n, k, 2), np.uint32)
    n_found, radius_out, nearest = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
    rc = L.pm_model_lookup(mode, rt._p(nodes), rt._p(dirs4), rt._p(points), rt._p(normals), n, k, C.c_float(r2), rt._p(radii), rt._p(found), rt._p(n_found), rt._p(radius_out),
                           rt._p(nearest), round_steps, single_pop)
    assert rc == 0
    return (found[:, :, 0], found[:, :, 1].view(np.float32), n_found, radius_out), nearest


@pytest.mark.parametrize("kind", KINDS)
def test_kernel_lookup_code_on_the_host_matches_oracle(built, host_model, kind):
    for n in (1, 2, 9, 5000):
        pos, dirs = scenes.photon_cloud(kind, n, seed=n + 5)
        points, normals = scenes.gather_points(pos, 700, seed=n + 6)
        o = pmo.OracleMap(pos, dirs)
        a, b = pm.host_tree(pos, 2)
        nodes = pm.pack_nodes(a, b, pos)
        dirs4 = np.zeros((n, 4), np.float32)
        dirs4[:, :3] = dirs
        for k, r2 in ((1, 1e-3), (2, 1e-2), (7, 1e-2), (64, 0.05), (33, 1e30)):
            want = o.gather(points, k, r2)
            # round_steps 0: the plain loop (pmLookupOne); otherwise the phased state machine (pmStep / pmResolve)
            for mode, round_steps, single_pop in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1), (0, 8, 1), (1, 8, 0), (1, 8, 1), (1, 1000, 1)):
                got, _ = run_model(host_model, mode, nodes, dirs4, points, normals, k, r2, round_steps=round_steps, single_pop=single_pop)
                assert_gather_equal(got, want, f"{kind} n={n} k={k} r2={r2} mode={mode} round_steps={round_steps} single_pop={single_pop}")
        radii = (np.random.default_rng(n).random(len(points)).astype(np.float32) * 0.05) ** 2
        for round_steps in (0, 8):
            got, _ = run_model(host_model, 0, nodes, dirs4, points, normals, 10, 0.0, radii, round_steps=round_steps, single_pop=1)
            assert_gather_equal(got, o.gather(points, 10, 0.0, radii), "per-point radii")
        for dist in (1e-4, 1e-2, 1.0):
            for round_steps in (0, 8):
                _, nearest = run_model(host_model, 2, nodes, dirs4, points, normals, 1, dist, round_steps=round_steps, single_pop=round_steps // 8)
                assert np.array_equal(nearest, o.nearest(points, normals, dist)), (kind, n, dist, round_steps)


# -------------------------------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_matches_golden(built, path):
    g = np.load(path)
    with pm.PhotonMap(g["pos"], g["dirs"]) as m:
        st = m.stats()
        assert st["n_photons"] == len(g["pos"]) and st["n_nodes"] == 2 * len(g["pos"]) - 1
        for i, (k, r2) in enumerate(g["gathers"]):
            want = (g[f"g{i}_idx"], g[f"g{i}_d2"], g[f"g{i}_n"], g[f"g{i}_r"])
            assert_gather_equal(product_result(*m.gather(g["points"], int(k), r2)), want, f"gather k={int(k)} r2={r2}")
        assert_gather_equal(product_result(*m.gather(g["points"], 12, 0.0, g["radii"])), (g["gr_idx"], g["gr_d2"], g["gr_n"], g["gr_r"]), "per-point radii")
        for i, dist in enumerate(g["nearest"]):
            assert np.array_equal(m.find_nearest(g["points"], g["normals"], dist), g[f"n{i}"]), f"findNearest dist={dist}"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_matches_oracle(built, kind):
    launches = rt.launch_count()
    for n in (1, 2, 3, 1000, 120_000):
        pos, dirs = scenes.photon_cloud(kind, n, seed=n + 21)
        points, normals = scenes.gather_points(pos, 6000, seed=n + 22)
        o = pmo.OracleMap(pos, dirs)
        with pm.PhotonMap(pos, dirs) as m:
            # k <= 256: heaps in shared memory; k = 300: heaps in the result array
            for k, r2 in ((1, 1e-3), (5, 1e-2), (100, 0.02), (256, 0.05), (300, 0.05), (50, 1e30), (100, 1e-7), (3, 0.0)):
                assert_gather_equal(product_result(*m.gather(points, k, r2)), o.gather(points, k, r2), f"{kind} n={n} k={k} r2={r2}")
            radii = (np.random.default_rng(n).random(len(points)).astype(np.float32) * 0.05) ** 2
            assert_gather_equal(product_result(*m.gather(points, 20, 0.0, radii)), o.gather(points, 20, 0.0, radii), "per-point radii")
            for dist in (1e-4, 1e-2, 1.0):
                assert np.array_equal(m.find_nearest(points, normals, dist), o.nearest(points, normals, dist)), (kind, n, dist)
    assert rt.launch_count() > launches, "no kernel of libb200rt.so was launched"


@pytest.mark.gpu
def test_gpu_kernel_variants_agree(built):
    """Every kernel variant behind b200pm_debug_set_tuning (plain loop, phased, phased + one pop per step; round length; heaps
    in shared memory or in the result array; make_heap patience) gives the reference's results, bit for bit."""
    pos, dirs = scenes.photon_cloud("surfaces", 150_000, seed=31)
    points, normals = scenes.gather_points(pos, 20_000, seed=32, jitter=0.004)
    o = pmo.OracleMap(pos, dirs)
    wants = {(k, r2): o.gather(points, k, r2) for k, r2 in ((1, 1e-3), (2, 1e-3), (3, 1e-3), (16, 2e-3), (60, 2e-3))}
    try:
        with pm.PhotonMap(pos, dirs) as m:
            for kernel, steps, smem_k, patience in ((0, 8, 0, 1), (0, 8, 256, 1), (1, 1, 0, 1), (1, 8, 16, 8), (2, 1, 0, 1), (2, 8, 0, 8), (2, 8, 256, 32), (2, 64, 16, 4), (2, 3, 0, 2)):
                pm.set_tuning(kernel, steps, smem_k, patience)
                for (k, r2), want in wants.items():
                    assert_gather_equal(product_result(*m.gather(points, k, r2)), want, f"kernel={kernel} steps={steps} smem_k={smem_k} patience={patience} k={k}")
    finally:
        pm.set_tuning(*DEFAULT_TUNING)


@pytest.mark.gpu
def test_gpu_device_buffer_entry_points(built):
    import torch

    pos, dirs = scenes.photon_cloud("surfaces", 50_000, seed=5)
    points, normals = scenes.gather_points(pos, 4096, seed=6)
    o = pmo.OracleMap(pos, dirs)
    with pm.PhotonMap(pos, dirs) as m:
        d_points, d_normals = torch.from_numpy(points).cuda(), torch.from_numpy(normals).cuda()
        found, n_found, radius_out = m.gather_device(d_points, 40, 0.01)
        nearest = m.find_nearest_device(d_points, d_normals, 0.01)
        torch.cuda.synchronize()
        f = found.cpu().numpy().view(np.uint32)
        got = (f[:, :, 0], f[:, :, 1].view(np.float32), n_found.cpu().numpy().view(np.uint32), radius_out.cpu().numpy())
        assert_gather_equal(got, o.gather(points, 40, 0.01), "device buffers")
        assert np.array_equal(nearest.cpu().numpy().view(np.uint32), o.nearest(points, normals, 0.01))
        # empty batch, missing directions
        assert m.gather(np.zeros((0, 3), np.float32), 4, 1.0)[1].shape == (0,)
    with pm.PhotonMap(pos) as m2, pytest.raises(rt.B200RTError):
        m2.find_nearest(points, normals, 0.01)


@pytest.mark.gpu
def test_gpu_full_size_properties(built):
    """1 M photons, 1 M gather points, k = 100 (the photon-mapping defaults: 'diffuse_photons' 1e6, 'diffuse_search' 100):
    size-independent properties over all points, the oracle on a sample, brute force on a smaller sample."""
    n, n_points, k, r2 = 1_000_000, 1_000_000, 100, 2.5e-4
    pos, dirs = scenes.photon_cloud("surfaces", n, seed=99)
    points, _ = scenes.gather_points(pos, n_points, seed=98, jitter=0.002)
    with pm.PhotonMap(pos, dirs) as m:
        found, n_found, radius_out = m.gather(points, k, r2)
    cnt = n_found.astype(np.int64)
    assert cnt.max() <= k and (cnt == k).mean() > 0.2 and (cnt < k).mean() > 0.02
    valid = np.arange(k)[None, :] < cnt[:, None]
    d2 = found["dist_square"]
    # every photon kept lies inside the search radius, at the distance reported (recomputed with the reference's operations)
    rows, cols = np.nonzero(valid)
    step = max(1, len(rows) // 4_000_000)
    rows, cols = rows[::step], cols[::step]
    v = pos[found["photon"][rows, cols]] - points[rows]
    recomputed = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    assert np.array_equal(bits(recomputed.astype(np.float32)), bits(d2[rows, cols]))
    assert (d2[rows, cols] < np.float32(r2)).all()
    full = cnt == k
    # a full result is a max-heap: the root is the farthest photon and is the radius handed back
    assert np.array_equal(bits(radius_out[full]), bits(d2[full, 0]))
    assert (d2[full].max(axis=1) == d2[full, 0]).all()
    parents = (np.arange(1, k) - 1) // 2
    assert (d2[full][:, parents] >= d2[full][:, 1:]).all()
    assert np.array_equal(bits(radius_out[~full]), np.full(int((~full).sum()), np.float32(r2)).view(np.uint32))
    # no photon twice
    srt = np.sort(np.where(valid, found["photon"], np.arange(k, dtype=np.uint32)[None, :] + np.uint32(0xF0000000)), axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all()
    # the oracle on a sample, bit for bit
    sample = np.random.default_rng(1).choice(n_points, 20_000, replace=False)
    o = pmo.OracleMap(pos, dirs)
    want = o.gather(points[sample], k, r2)
    assert_gather_equal((found["photon"][sample], d2[sample], n_found[sample], radius_out[sample]), want, "full size, sampled")
    # brute force on a smaller sample: the k nearest photons inside the radius, as a set (ties at the k-th distance aside)
    for i in sample[:100]:
        v = pos - points[i]
        all_d2 = ((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]).astype(np.float32)
        inside = np.nonzero(all_d2 < np.float32(r2))[0]
        if len(inside) <= k:
            assert set(inside) == set(found["photon"][i, :cnt[i]])
        else:
            kth = np.sort(all_d2[inside])[k - 1]
            assert cnt[i] == k and d2[i].max() == kth
            assert set(inside[all_d2[inside] < kth]) <= set(found["photon"][i])
