"""Shared helpers of the test-suite (scene zoo, conversions)."""
import numpy as np

from libyafaray_b200 import scenes


def flag_mix(n_faces, seed=1):
    """Per-face flags exercising every visibility combination and the transparent bit."""
    rng = np.random.RandomState(seed)
    fl = np.full(n_faces, scenes.F_NORMAL, np.uint8)
    r = rng.random_sample(n_faces)
    fl[r < 0.1] = 0
    fl[(r >= 0.1) & (r < 0.2)] = scenes.F_VISIBLE
    fl[(r >= 0.2) & (r < 0.3)] = scenes.F_SHADOW
    fl[(r >= 0.3) & (r < 0.5)] |= scenes.F_TRANSPARENT
    return fl


def scene_zoo(small=True):
    """name -> (xyz, idx, flags): triangles, quads, mixed, random soup, axis-aligned boxes with coplanar faces."""
    k = 1 if small else 4
    zoo = {
        "hf": scenes.heightfield(40 * k),
        "hf_quads": scenes.heightfield(30 * k, quads=True),
        "cubes": scenes.cube_scene(),
        "soup": scenes.soup(3000 * k * k),
        "objects": scenes.objects(6000 * k * k, n_spheres=12),
    }
    # spheres (SpherePrimitive) among mesh objects: mesh faces first, then the spheres (scenes.with_spheres)
    oxyz, oidx, ofl = scenes.objects(3000 * k * k, n_spheres=8, seed=21)
    zoo["spheres"] = scenes.with_spheres(oxyz, oidx, ofl, scenes.sphere_field(60 * k * k, seed=22, lo=oxyz.min(0), hi=oxyz.max(0)))
    # instanced axis-aligned cubes with faces planar up to one ulp (the reference's tests/test02): 1-ulp kd slabs
    zoo["cube_grid"] = scenes.cube_grid(5 if small else 9)
    out = {}
    for i, (name, (xyz, idx, fl)) in enumerate(zoo.items()):
        out[name] = (xyz, idx, fl)
        out[name + "_flags"] = (xyz, idx, flag_mix(idx.shape[0], seed=10 + i))
    return out


def ray_zoo(bound, n=20000, seed=3, edge_cases=True):
    """edge_cases=False leaves out the axis-parallel / exact-diagonal rays of scenes.rays_edge_cases: through a SYMMETRIC grid of
    axis-aligned cubes (cube_grid) they pass exactly through cube corners, where which of the eight cells around the corner a
    kd traversal walks through -- the reference's included -- depends on the rounding of three equal plane distances and on the
    tree, so two valid trees may disagree on such a ray (a corner leak, not a tie in t)."""
    lo, hi = bound[:3].astype(np.float64), bound[3:].astype(np.float64)
    ext = hi - lo
    diag = float(np.linalg.norm(ext))
    closest = np.concatenate([
        scenes.rays_incoherent(n, seed=seed, lo=lo - 0.1 * ext, hi=hi + 0.1 * ext),
        scenes.rays_incoherent(n // 4, seed=seed + 1, lo=lo, hi=hi, tmax=0.3 * diag, tmin=0.01 * diag),
        scenes.rays_edge_cases(lo, hi, seed=seed + 2) if edge_cases else np.zeros((0, 8), np.float32),
    ])
    shadow = np.concatenate([
        scenes.rays_shadow(n, seed=seed + 3, lo=lo, hi=hi, t_max=0.25 * diag),
        scenes.rays_incoherent(n // 4, seed=seed + 4, lo=lo, hi=hi, tmax=-1.0, tmin=0.0005),
        scenes.rays_edge_cases(lo, hi, seed=seed + 5) if edge_cases else np.zeros((0, 8), np.float32),
    ])
    return closest, shadow


def host_tree_as_oracle_tree(t):
    """libb200rt's exported host tree (include/b200rt.h) -> the oracle's kdo_tree arrays."""
    a, b = t["a"], t["b"]
    leaf = (b & 3) == 3
    split = np.where(leaf, 0, a).astype(np.uint32).view(np.float32)
    first = np.where(leaf, a, 0).astype(np.uint32)
    refs = t["refs"] if len(t["refs"]) else np.zeros(1, np.uint32)
    return dict(split=split, flags=b.astype(np.uint32), first_ref=first, refs=refs.astype(np.uint32))


def prim_signed(prim_u32):
    p = prim_u32.astype(np.int64)
    p[p == 0xFFFFFFFF] = -1
    return p


def check_closest_parity(got_prim, got_t, got_u, got_v, ref, min_agree=0.9999, rtol=1e-5):
    """North-star bar: same face id on >= 99.99 % of rays, t within 1e-5 relative; where ids agree the
    arithmetic is the reference's, so t/u/v must be BIT-identical; where they differ it must be a tie."""
    same = got_prim == ref["prim"]
    agree = float(np.mean(same)) if len(same) else 1.0
    assert agree >= min_agree, f"id agreement {agree:.6f} < {min_agree}"
    assert np.array_equal(got_t[same], ref["t"][same]), "t differs bitwise on rays with equal ids"
    assert np.array_equal(got_u[same], ref["u"][same]) and np.array_equal(got_v[same], ref["v"][same]), "uv differs bitwise"
    diff = ~same
    if diff.any():
        # a documented tie: both sides hit, and the two t agree to 1e-5 relative
        assert np.all(got_prim[diff] >= 0) and np.all(ref["prim"][diff] >= 0), "hit/miss disagreement"
        rel = np.abs(got_t[diff] - ref["t"][diff]) / np.maximum(np.abs(ref["t"][diff]), 1e-30)
        assert np.all(rel <= rtol), f"non-tie id mismatch, max rel t diff {rel.max()}"
    return agree


def make_rt_scene(rt, xyz, idx, flags=None, params=None, device=0):
    """rt.Scene from the flat arrays of the zoo: mesh faces through add_mesh, sphere marker faces (scenes.SPHERE) through
    add_spheres, in face order, so that face ids equal array rows."""
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
    s = rt.Scene(device, params)
    is_sphere = idx[:, 2] == scenes.SPHERE
    if not is_sphere.any():
        s.add_mesh(xyz, idx, flags)
    else:
        f = 0
        n = idx.shape[0]
        while f < n:
            e = f
            while e < n and is_sphere[e] == is_sphere[f]:
                e += 1
            fl = None if flags is None else flags[f:e]
            if is_sphere[f]:
                cr = np.concatenate([xyz[idx[f:e, 0]], xyz[idx[f:e, 1], :1]], axis=1)
                s.add_spheres(cr, fl)
            else:
                s.add_mesh(xyz, idx[f:e], fl)
            f = e
    s.build()
    return s


def load_scene_dump(path):
    """Geometry dumped by AcceleratorB200 (B200_DUMP_SCENE, integration/src/accelerator/accelerator_b200.cc): one record per
    b200rt_add_mesh call {uint64 n_verts, uint64 n_faces, float xyz[3 n_verts], uint32 idx[4 n_faces], uint8 flags[n_faces]}."""
    raw = open(path, "rb").read()
    off, xs, ids, fs, base = 0, [], [], [], 0
    while off < len(raw):
        nv, nf = (int(x) for x in np.frombuffer(raw, "<u8", 2, off)); off += 16
        x = np.frombuffer(raw, "<f4", 3 * nv, off).reshape(nv, 3); off += 12 * nv
        i = np.frombuffer(raw, "<u4", 4 * nf, off).reshape(nf, 4).copy(); off += 16 * nf
        f = np.frombuffer(raw, "u1", nf, off); off += nf
        tri = i[:, 3] == 0xFFFFFFFF
        i[:, :3] += base
        i[~tri, 3] += base
        xs.append(x); ids.append(i); fs.append(f); base += nv
    return np.concatenate(xs), np.concatenate(ids), np.concatenate(fs)


def surface_rays(rays, t, seed, tmin=0.0005):
    """Secondary rays as the integrators shoot them: origin = from + t * dir of a hit (float arithmetic), random direction."""
    rng = np.random.default_rng(seed)
    hit = t > 0
    p = (rays[hit, 0:3] + t[hit, None] * rays[hit, 4:7]).astype(np.float32)
    out = np.zeros((p.shape[0], 8), np.float32)
    out[:, 0:3] = p; out[:, 3] = tmin; out[:, 4:7] = rng.normal(size=p.shape).astype(np.float32); out[:, 7] = -1.0
    return out


def parity_counts(got_prim, got_t, ref, rtol=1e-5):
    """(ids equal incl. ties, id mismatches, of which not ties: hit/miss differs or t apart by more than rtol)."""
    same = got_prim == ref["prim"]
    diff = np.flatnonzero(~same)
    rel = np.abs(got_t[diff].astype(np.float64) - ref["t"][diff]) / np.maximum(np.abs(ref["t"][diff]), 1e-30)
    non_tie = (rel > rtol) | ((got_prim[diff] >= 0) != (ref["prim"][diff] >= 0))
    return float(same.mean()) if len(same) else 1.0, int(diff.size), int(non_tie.sum())


def report(kind, **fields):
    """Append one JSON line to gpurun_out/parity_report.jsonl (travels back from the GPU box; summarised under profiles/)."""
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(kind=kind, **fields)) + "\n")
    except OSError:
        pass


def cube_grid_far():
    """The thin-slab cube grid moved to coordinates of 300 (ulp 3e-5): 1-ulp slabs AND a coarse float grid."""
    x, i, f = scenes.cube_grid(5)
    return (x + np.array([300.0, 300.0, 300.0], np.float32)).astype(np.float32), i, f


def make_rt_motion_scene(rt, xyz, idx, flags, motion, params=None, device=0):
    """rt.Scene of a scenes.motion_scene(): runs of faces of one kind / time range / instance go to add_mesh, add_mesh_bezier and
    add_mesh_moving in face order, so that face ids equal array rows (as oracle.yref.RefScene builds the reference's scene)."""
    s = rt.Scene(device, params)
    kind, ft, fm = motion["kind"], motion["face_times"], motion["face_matrix"]
    n, f = idx.shape[0], 0
    while f < n:
        e = f
        while e < n and kind[e] == kind[f] and np.array_equal(ft[e], ft[f]) and (kind[f] != 2 or fm[e] == fm[f]):
            e += 1
        if kind[f] == 0:
            s.add_mesh(xyz, idx[f:e], flags[f:e])
        elif kind[f] == 1:
            s.add_mesh_bezier(xyz, motion["xyz1"], motion["xyz2"], idx[f:e], flags[f:e], (float(ft[f, 0]), float(ft[f, 1])))
        else:
            s.add_mesh_moving(xyz, idx[f:e], motion["matrices"][fm[f]], flags[f:e], (float(ft[f, 0]), float(ft[f, 1])))
        f = e
    s.build()
    return s
