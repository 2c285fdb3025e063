"""Deterministic synthetic scenes and ray batches.

Both arms (the CUDA path and the CPU oracle / reference) consume exactly the arrays produced here, so
"identical bytes on both sides" holds by construction.  numpy only; nothing here touches the GPU.

Mesh convention (the one `b200rt_upload_mesh` takes, include/b200rt.h):
  xyz   float32 [n_verts, 3]
  idx   uint32  [n_faces, 4]   idx[f, 3] == 0xFFFFFFFF marks a triangle, otherwise a quad
  flags uint8   [n_faces]      bit0 Visible, bit1 CastsShadows, bit2 transparent material
                               (the quad bit3 is derived from idx by the library)
Ray convention: float32 [n, 8] = ox oy oz tmin dx dy dz tmax  (tmax < 0 means "unbounded", as
`Ray::tmax_` does in the reference, include/accelerator/accelerator.h:91).
"""
from __future__ import annotations

import numpy as np

TRI = np.uint32(0xFFFFFFFF)
F_VISIBLE = 1
F_SHADOW = 2
F_TRANSPARENT = 4
F_NORMAL = F_VISIBLE | F_SHADOW


def _finish(xyz, idx, flags=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    if flags is None:
        flags = np.full(idx.shape[0], F_NORMAL, dtype=np.uint8)
    return xyz, idx, np.ascontiguousarray(flags, dtype=np.uint8)


def heightfield(cells: int = 707, seed: int = 12345, quads: bool = False):
    """S1M-hf: `cells` x `cells` grid over [0,1]^2, two triangles per cell (707 -> 999 698 triangles).

    z = 0.5 + 0.15 sin(17x) cos(13y) + 0.05 sin(71x + 3y) + 0.002 U(0,1)   
    """
    rng = np.random.RandomState(seed)
    n = cells + 1
    g = np.linspace(0.0, 1.0, n)
    x, y = np.meshgrid(g, g, indexing="xy")
    z = 0.5 + 0.15 * np.sin(17 * x) * np.cos(13 * y) + 0.05 * np.sin(71 * x + 3 * y) + 0.002 * rng.random_sample(x.shape)
    xyz = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    j, i = np.meshgrid(np.arange(cells), np.arange(cells), indexing="xy")
    v00 = (i * n + j).ravel()
    v01 = v00 + 1
    v10 = v00 + n
    v11 = v10 + 1
    if quads:
        idx = np.stack([v00, v01, v11, v10], axis=1)
    else:
        t0 = np.stack([v00, v01, v11, np.full_like(v00, TRI)], axis=1)
        t1 = np.stack([v00, v11, v10, np.full_like(v00, TRI)], axis=1)
        idx = np.stack([t0, t1], axis=1).reshape(-1, 4)
    return _finish(xyz, idx)


def soup(n_tris: int = 1_000_000, seed: int = 12345, jitter_scale: float = 0.25):
    """S1M-soup: random triangles, centres U[0,1]^3, vertex jitter +-jitter_scale/cbrt(N) (worst case)."""
    rng = np.random.RandomState(seed)
    c = rng.random_sample((n_tris, 1, 3))
    r = jitter_scale / np.cbrt(n_tris)
    v = c + (rng.random_sample((n_tris, 3, 3)) * 2.0 - 1.0) * r
    xyz = v.reshape(-1, 3)
    base = np.arange(n_tris, dtype=np.int64) * 3
    idx = np.stack([base, base + 1, base + 2, np.full_like(base, TRI)], axis=1)
    return _finish(xyz, idx)


def _uv_sphere(center, radius, n_lat, n_lon):
    th = np.linspace(0.0, np.pi, n_lat + 1)
    ph = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False)
    t, p = np.meshgrid(th, ph, indexing="ij")
    xyz = np.stack([np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)], axis=-1).reshape(-1, 3) * radius + center
    a = (np.arange(n_lat)[:, None] * n_lon + np.arange(n_lon)[None, :]).ravel()
    b = (np.arange(n_lat)[:, None] * n_lon + (np.arange(n_lon)[None, :] + 1) % n_lon).ravel()
    c = a + n_lon
    d = b + n_lon
    t0 = np.stack([a, c, d], axis=1)
    t1 = np.stack([a, d, b], axis=1)
    return xyz, np.concatenate([t0, t1], axis=0)


def objects(n_tris_target: int = 1_000_000, seed: int = 12345, n_spheres: int = 64, mixed_quads: bool = True):
    """S1M-obj: tessellated spheres over a tessellated ground plane inside [0,1]^3 (surface-like, occluding).

    The ground is made of quads when `mixed_quads` (the reference's test01 mixes quads and triangles).
    """
    rng = np.random.RandomState(seed)
    per = max(8, n_tris_target // (n_spheres + 1))
    n_lat = max(2, int(np.sqrt(per / 4.0)))
    n_lon = 2 * n_lat
    vs, fs, off = [], [], 0
    for _ in range(n_spheres):
        c = np.array([rng.uniform(0.1, 0.9), rng.uniform(0.1, 0.9), rng.uniform(0.15, 0.8)])
        r = rng.uniform(0.03, 0.09)
        xyz, tri = _uv_sphere(c, r, n_lat, n_lon)
        vs.append(xyz)
        fs.append(np.concatenate([tri + off, np.full((tri.shape[0], 1), int(TRI), dtype=np.int64)], axis=1))
        off += xyz.shape[0]
    cells = max(1, int(np.sqrt(per / (1.0 if mixed_quads else 2.0))))
    gx, gi, _ = heightfield(cells, seed + 1, quads=mixed_quads)
    gx = gx.copy()
    gx[:, 2] = 0.05 + 0.2 * (gx[:, 2] - 0.5)
    vs.append(gx)
    gi = gi.astype(np.int64)
    tri_mask = gi[:, 3] == int(TRI)
    gi[:, :3] += off
    gi[~tri_mask, 3] += off
    fs.append(gi)
    return _finish(np.concatenate(vs, axis=0), np.concatenate(fs, axis=0))


def cube_scene():
    """A tiny mixed quad/triangle scene in the spirit of tests/test01 (boxes on a two-triangle plane)."""
    vs, fs, off = [], [], 0
    quad_faces = np.array([[2, 0, 1, 3], [3, 7, 6, 2], [7, 5, 4, 6], [0, 4, 5, 1], [0, 2, 6, 4], [5, 7, 3, 1]])
    corners = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (0, 2)], dtype=np.float64)
    centers = [(-3.4, 2.4, 0.0), (4.27, 0.6, 0.0), (0.4, -1.2, 0.0), (-0.5, 3.0, 0.0), (2.0, 4.0, 0.5), (-4.0, -3.0, 0.25)]
    for k, c in enumerate(centers):
        v = corners * (0.5 + 0.15 * k) + np.array(c)
        vs.append(v)
        if k % 2 == 0:
            fs.append(quad_faces + off)
        else:
            t0 = np.concatenate([quad_faces[:, [0, 1, 2]], np.full((6, 1), int(TRI))], axis=1)
            t1 = np.concatenate([quad_faces[:, [0, 2, 3]], np.full((6, 1), int(TRI))], axis=1)
            t = np.concatenate([t0, t1], axis=0)
            t[:, :3] += off
            fs.append(t)
        off += 8
    plane = np.array([[-10, -10, 0], [10, -10, 0], [10, 10, 0], [-10, 10, 0]], dtype=np.float64)
    vs.append(plane)
    fs.append(np.array([[off, off + 1, off + 2, int(TRI)], [off, off + 2, off + 3, int(TRI)]]))
    return _finish(np.concatenate(vs, axis=0), np.concatenate(fs, axis=0))


# ----------------------------------------------------------------------------------------------- rays

def rays_incoherent(n: int, seed: int = 12345, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), tmax: float = -1.0, tmin: float = 0.0):
    """R-inc: origin uniform in the box [lo,hi], direction = normalised U[-.5,.5]^3."""
    rng = np.random.RandomState(seed)
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    r = np.empty((n, 8), dtype=np.float32)
    r[:, 0:3] = lo + rng.random_sample((n, 3)) * (hi - lo)
    d = rng.random_sample((n, 3)) - 0.5
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    r[:, 4:7] = d
    r[:, 3] = tmin
    r[:, 7] = tmax
    return r


def rays_shadow(n: int, seed: int = 12345, t_max: float = 0.25, **kw):
    """R-shadow (first set): same distribution as R-inc with a finite `t_max` and the usual shadow bias
    tmin = 0.0005 (Accelerator::shadowBias, include/accelerator/accelerator.h:86)."""
    return rays_incoherent(n, seed, tmax=t_max, tmin=0.0005, **kw)


def rays_camera(width: int, height: int, eye=(0.5, -1.2, 1.1), look=(0.5, 0.5, 0.45), up=(0, 0, 1), fov_deg: float = 50.0, seed: int | None = None):
    """R-coh: one perspective camera ray per pixel (pin-hole; direction normalised), optional sub-pixel jitter."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(look, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    px, py = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64), indexing="xy")
    if seed is not None:
        rng = np.random.RandomState(seed)
        px = px + rng.random_sample(px.shape)
        py = py + rng.random_sample(py.shape)
    else:
        px, py = px + 0.5, py + 0.5
    half = np.tan(np.radians(fov_deg) / 2)
    sx = (2 * px / width - 1) * half
    sy = (1 - 2 * py / height) * half * height / width
    d = fwd[None, None, :] + sx[..., None] * right + sy[..., None] * upv
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    r = np.empty((width * height, 8), dtype=np.float32)
    r[:, 0:3] = eye
    r[:, 3] = 0.0
    r[:, 4:7] = d.reshape(-1, 3)
    r[:, 7] = -1.0
    return r


def rays_edge_cases(bound_lo, bound_hi, seed: int = 7):
    """Hand-picked nasty rays: axis-parallel directions (zero components), origins on the bound, outside
    pointing away, tmax = 0, tiny tmax, un-normalised and huge directions."""
    rng = np.random.RandomState(seed)
    lo = np.asarray(bound_lo, dtype=np.float64)
    hi = np.asarray(bound_hi, dtype=np.float64)
    c = 0.5 * (lo + hi)
    ext = hi - lo
    rows = []
    for ax in range(3):
        for s in (-1.0, 1.0):
            d = np.zeros(3)
            d[ax] = s
            for _ in range(64):
                o = lo + rng.random_sample(3) * ext
                rows.append([*o, 0.0, *d, -1.0])
                o2 = o.copy()
                o2[ax] = c[ax] - s * ext[ax]
                rows.append([*o2, 0.0, *d, -1.0])
                rows.append([*o2, 0.0, *(-d), -1.0])            # outside, pointing away
                rows.append([*o2, 0.0, *(d * 1e-3), -1.0])      # un-normalised, short
                rows.append([*o2, 0.0, *(d * 1e4), -1.0])       # un-normalised, long
            for a2 in range(3):
                if a2 == ax:
                    continue
                d2 = d.copy()
                d2[a2] = 0.37
                for _ in range(32):
                    o = lo + rng.random_sample(3) * ext
                    rows.append([*o, 0.0, *d2, -1.0])
    for _ in range(256):
        o = lo + rng.random_sample(3) * ext
        d = rng.random_sample(3) - 0.5
        rows.append([*o, 0.0, *d, 0.0])                          # tmax == 0
        rows.append([*o, 0.0, *d, 1e-6])
        rows.append([*o, 0.01, *d, 0.05])
        rows.append([*lo, 0.0, *d, -1.0])                        # origin on the bound corner
        rows.append([*(hi + ext), 0.0, *(c - hi - ext), -1.0])   # from outside through the centre
    return np.asarray(rows, dtype=np.float32)


SPHERE = 0xFFFFFFFE  # face marker in idx[:, 2]: vertex idx[:, 0] = centre, x of vertex idx[:, 1] = radius (include/b200rt.h)


def with_spheres(xyz, idx, flags, spheres, sphere_flags=None):
    """Append spheres (rows cx, cy, cz, radius -- SpherePrimitive, src/geometry/primitive/primitive_sphere.cc) to a mesh in the
    flat encoding the oracle and the host-tree diagnostics read: two extra vertices and one marker face per sphere.
    The product path takes spheres through rt.Scene.add_spheres (b200rt_add_spheres) instead; face ids come out the same
    (mesh faces first, then the spheres in order)."""
    spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
    n = spheres.shape[0]
    base = xyz.shape[0]
    extra = np.zeros((2 * n, 3), dtype=np.float32)
    extra[0::2] = spheres[:, :3]
    extra[1::2, 0] = spheres[:, 3]
    faces = np.empty((n, 4), dtype=np.uint32)
    faces[:, 0] = base + 2 * np.arange(n)
    faces[:, 1] = base + 2 * np.arange(n) + 1
    faces[:, 2] = SPHERE
    faces[:, 3] = TRI
    if sphere_flags is None:
        sphere_flags = np.full(n, F_NORMAL, dtype=np.uint8)
    return (np.concatenate([xyz, extra]).astype(np.float32), np.concatenate([idx, faces]).astype(np.uint32),
            np.concatenate([flags, np.asarray(sphere_flags, dtype=np.uint8)]).astype(np.uint8))


def sphere_field(n: int = 200, seed: int = 5, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), r_lo: float = 0.01, r_hi: float = 0.06):
    """n random spheres inside [lo, hi]: rows cx, cy, cz, radius."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    out = np.empty((n, 4), dtype=np.float32)
    out[:, :3] = lo + rng.random((n, 3), dtype=np.float32) * (hi - lo)
    out[:, 3] = r_lo + rng.random(n, dtype=np.float32) * (r_hi - r_lo)
    return out


def cube_grid(n_per_axis: int = 9, scale: float = 0.06, distance: float = 0.35, floor: bool = True):
    """A grid of small axis-aligned cubes, the instanced "Cube" of the reference's tests/test02/test02.c:3546-3569: two of its
    vertices are off by 1e-6, so that after the instance transform (scale 0.06 + translation, in float) the faces are planar only
    up to one ulp.  A SAH builder cuts 1-ulp slabs around such faces, thinner than the resolution of t along most rays -- the
    case in which a t-interval traversal must still enter a slab whose entry and exit parameter coincide (kd_kernels.cuh)."""
    base = np.array([[1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1], [1, 0.999999, 1], [0.999999, -1, 1], [-1, -1, 1], [-1, 1, 1]], dtype=np.float32)
    faces = np.array([[0, 1, 2], [0, 2, 3], [4, 7, 6], [4, 6, 5], [0, 4, 5], [0, 5, 1], [1, 5, 6], [1, 6, 2], [2, 6, 7], [2, 7, 3], [4, 0, 3], [4, 3, 7]], dtype=np.uint32)
    s = np.float32(scale)
    xyz, idx = [], []
    start = np.float32(-0.5 * distance * (n_per_axis - 1))
    for i in range(n_per_axis):
        for j in range(n_per_axis):
            for k in range(n_per_axis):
                offset = np.array([start + np.float32(distance) * i, start + np.float32(distance) * j, start + np.float32(distance) * k], dtype=np.float32)
                idx.append(np.concatenate([faces + np.uint32(8 * len(xyz)), np.full((12, 1), TRI, np.uint32)], axis=1))
                xyz.append((base * s + offset).astype(np.float32))  # m00 * x + m03, one float multiply and one float add like Matrix4f * Point3f
    xyz, idx = np.concatenate(xyz), np.concatenate(idx)
    if floor:
        e = np.float32(1.2 * abs(float(start)) + 2 * scale)
        z = np.float32(float(start) - 2 * scale)
        base_v = xyz.shape[0]
        xyz = np.concatenate([xyz, np.array([[-e, -e, z], [e, -e, z], [e, e, z], [-e, e, z]], dtype=np.float32)])
        idx = np.concatenate([idx, np.array([[base_v, base_v + 1, base_v + 2, base_v + 3]], dtype=np.uint32)])
    return _finish(xyz, idx)


# ----------------------------------------------------------------------------------------------- motion blur

def _offset_faces(idx, off):
    out = idx.astype(np.int64).copy()
    tri = out[:, 3] == int(TRI)
    out[:, :3] += off
    out[~tri, 3] += off
    return out


def _rigid(angle, axis, translate, scale=1.0):
    """4x4 row-major obj_to_world: rotation about `axis` by `angle` (Rodrigues), uniform scale, translation."""
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    r = np.eye(3) + np.sin(angle) * k + (1 - np.cos(angle)) * (k @ k)
    m = np.eye(4)
    m[:3, :3] = scale * r
    m[:3, 3] = translate
    return m


def motion_scene(n_static: int = 3000, n_bezier: int = 2000, n_moving: int = 1500, seed: int = 3):
    """Static geometry + one Bezier motion-blur mesh + two moving instances.

    Returns (xyz, idx, flags, motion) in the conventions of this module, face rows in the order the reference numbers its
    primitives: objects first (static mesh, then the motion-blur mesh), instance primitives last.  motion = dict(
      kind        u8 [n_faces]        0 static, 1 face of the Bezier motion-blur mesh, 2 face of a moving instance
      xyz1, xyz2  f32 [n_verts, 3]    Bezier control positions of time steps 1 and 2 (step 0 is xyz), as a motion-blur MeshObject stores
                                      them -- it replaces the mid-time position p1 by the control point 2 p1 - (p0 + p2) / 2 when it is
                                      initialised (src/geometry/object/object_mesh.cc:268-277), which the accelerator then reads
      xyz1_user   f32 [n_verts, 3]    the mid-time positions themselves (what a client passes to yafaray_addVertexTimeStep)
      face_times  f32 [n_faces, 2]    time range (start, end) of the face's mesh / instance; (0, 0) for static faces
      face_matrix u32 [n_faces]       kind 2: index of the instance
      matrices    f32 [n_inst, 3, 16] row-major obj_to_world at the three time steps)
    For kind 2 faces xyz holds the BASE object's vertices; the world position is matrix(t) * vertex."""
    rng = np.random.RandomState(seed)
    sx, si, sf = objects(n_static, seed=seed, n_spheres=10)
    bx, bi, bf = objects(n_bezier, seed=seed + 1, n_spheres=6)
    bx = (bx * np.float32(0.6) + np.array([0.2, 0.2, 0.3], np.float32)).astype(np.float32)
    # the Bezier control positions: a drift plus a bend that differs per vertex, so that faces really deform
    drift1 = np.array([0.05, -0.03, 0.04]); drift2 = np.array([0.12, 0.02, -0.05])
    wob = rng.normal(scale=0.01, size=bx.shape)
    bx1 = (bx + drift1 + wob).astype(np.float32)
    bx2 = (bx + drift2 - 0.5 * wob + 0.05 * np.sin(6.0 * bx[:, [1, 2, 0]])).astype(np.float32)
    meshes = []
    for k in range(2):
        mx, mi, mf = objects(max(64, n_moving // 2), seed=seed + 2 + k, n_spheres=4)
        meshes.append(((mx * np.float32(0.35)).astype(np.float32), mi, mf))
    n_s, n_b = sx.shape[0], bx.shape[0]
    xyz = [sx, bx]; xyz1 = [sx, bx1]; xyz2 = [sx, bx2]
    idx = [si.astype(np.int64), _offset_faces(bi, n_s)]
    flags = [sf, bf]
    kind = [np.zeros(si.shape[0], np.uint8), np.ones(bi.shape[0], np.uint8)]
    times = [np.zeros((si.shape[0], 2), np.float32), np.tile(np.array([[0.1, 0.9]], np.float32), (bi.shape[0], 1))]
    fmat = [np.zeros(si.shape[0], np.uint32), np.zeros(bi.shape[0], np.uint32)]
    mats = []
    off = n_s + n_b
    ranges = [(0.0, 1.0), (0.25, 0.75)]
    for k, (mx, mi, mf) in enumerate(meshes):
        xyz.append(mx); xyz1.append(mx); xyz2.append(mx)
        idx.append(_offset_faces(mi, off)); off += mx.shape[0]
        flags.append(mf)
        kind.append(np.full(mi.shape[0], 2, np.uint8))
        times.append(np.tile(np.array([ranges[k]], np.float32), (mi.shape[0], 1)))
        fmat.append(np.full(mi.shape[0], k, np.uint32))
        base = np.array([0.1 + 0.5 * k, 0.55 - 0.4 * k, 0.3 + 0.2 * k])
        mats.append(np.stack([_rigid(0.0 + 0.3 * k, (0, 0, 1), base).reshape(16),
                              _rigid(0.4 + 0.3 * k, (0.2, 0.1, 1), base + np.array([0.08, 0.05, 0.06]), 1.1).reshape(16),
                              _rigid(0.9 + 0.3 * k, (0.3, -0.2, 1), base + np.array([0.1, 0.15, -0.04]), 0.9).reshape(16)]))
    x, i, f = _finish(np.concatenate(xyz), np.concatenate(idx), np.concatenate(flags))
    user1 = np.ascontiguousarray(np.concatenate(xyz1), np.float32)
    last = np.ascontiguousarray(np.concatenate(xyz2), np.float32)
    control = (np.float32(2) * user1 - (x + last) / np.float32(2)).astype(np.float32)  # math::bezierFindControlPoint, float arithmetic
    motion = dict(kind=np.concatenate(kind), xyz1=control, xyz1_user=user1, xyz2=np.ascontiguousarray(np.concatenate(xyz2), np.float32),
                  face_times=np.ascontiguousarray(np.concatenate(times), np.float32), face_matrix=np.concatenate(fmat),
                  matrices=np.ascontiguousarray(np.stack(mats), np.float32))
    return x, i, f, motion


def ray_times(n: int, seed: int = 1):
    """Ray::time_ values (include/geometry/ray.h:49): uniform in [0, 1] with the range ends and a few out-of-range values mixed in."""
    rng = np.random.RandomState(seed)
    t = rng.random_sample(n).astype(np.float32)
    special = np.array([0.0, 1.0, 0.1, 0.9, 0.25, 0.75, 0.5, -0.5, 1.5], np.float32)
    t[:: max(1, n // 200)] = special[np.arange(len(t[:: max(1, n // 200)])) % len(special)]
    return t


# ---- photon maps: synthetic photon sets and gather points ------------------------------------------

def photon_cloud(kind: str = "surfaces", n: int = 100_000, seed: int = 12345):
    """Synthetic photons: (pos [n,3] f32, dir [n,3] f32).

    uniform   -- uniform in the unit cube
    surfaces  -- on the six walls and the floor-parallel shelves of a box (what a diffuse photon map looks like: every photon
                 shares one coordinate with thousands of others, so the median splits meet exact ties)
    clusters  -- tight Gaussian blobs on a floor plus a thin uniform background (a caustic map)
    lattice   -- an integer lattice with every point stored several times (exact duplicates: the build's index tie-break
                 and the lookup's `<=` decide)
    """
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        pos = rng.random((n, 3), dtype=np.float32)
    elif kind == "surfaces":
        pos = rng.random((n, 3), dtype=np.float32)
        which = rng.integers(0, 8, n)
        axis = np.array([0, 0, 1, 1, 2, 2, 2, 2])[which]
        value = np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.25, 0.625], np.float32)[which]
        pos[np.arange(n), axis] = value
    elif kind == "clusters":
        n_bg = n // 10
        centres = rng.random((24, 3), dtype=np.float32)
        centres[:, 2] = 0.0
        which = rng.integers(0, len(centres), n - n_bg)
        blob = centres[which] + (rng.standard_normal((n - n_bg, 3)) * np.array([0.01, 0.01, 0.0])).astype(np.float32)
        pos = np.concatenate([blob.astype(np.float32), rng.random((n_bg, 3), dtype=np.float32)])
        pos = pos[rng.permutation(n)]
    elif kind == "lattice":
        side = max(2, int(round((n / 3) ** (1 / 3))))
        pts = rng.integers(0, side, (n, 3)).astype(np.float32) / np.float32(side)
        pos = pts
    else:
        raise ValueError(kind)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20).astype(np.float32)
    return np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(d, np.float32)


def gather_points(pos, n: int, seed: int = 777, jitter: float = 0.01, outside: float = 0.05):
    """Query points of a gather: most of them near photons (a shading point lies on the surface the photons landed on), a few
    exactly ON a photon, a few outside the map's bound.  Returns (points [n,3] f32, normals [n,3] f32)."""
    rng = np.random.default_rng(seed)
    base = pos[rng.integers(0, len(pos), n)]
    pts = base + (rng.standard_normal((n, 3)) * jitter).astype(np.float32)
    exact = rng.random(n) < 0.05
    pts[exact] = base[exact]
    far = rng.random(n) < outside
    pts[far] = (rng.random((int(far.sum()), 3)) * 3.0 - 1.0).astype(np.float32)
    nrm = rng.standard_normal((n, 3)).astype(np.float32)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20).astype(np.float32)
    return np.ascontiguousarray(pts, np.float32), np.ascontiguousarray(nrm, np.float32)
