"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference compiled into oracle/_ref/libyafref.so by
`make -C oracle ref`):   python tests/golden/make_golden.py

Each <name>.npz holds one small mesh (xyz, idx, flags), two ray batches, what the reference's own
Accelerator::intersect / isShadowed / isShadowedTransparentShadow returned for them (default accelerator:
"yafaray-kdtree-original", no params -- the configuration of the reference's tests/test01), the reference's
tree bound, and the reference's kd-tree in flat form so that the C restatement can be checked bit-for-bit,
ties included, where the reference library itself is absent (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from libyafaray_b200 import scenes  # noqa: E402
from oracle import yref  # noqa: E402
from tests.helpers import flag_mix, ray_zoo  # noqa: E402

TSHADOW_DEPTH = 3


def main():
    assert yref.available(), "build oracle/_ref first: make -C oracle ref"
    cases = {
        "hf": scenes.heightfield(24),
        "hf_quads": scenes.heightfield(20, quads=True),
        "cubes": scenes.cube_scene(),
        "soup": scenes.soup(1500, seed=4),
        "objects": scenes.objects(2500, n_spheres=6, seed=9),
    }
    # SpherePrimitive objects next to mesh faces (scenes.with_spheres encoding: marker faces after the mesh faces)
    sx, si, sf = scenes.objects(1200, n_spheres=4, seed=14)
    cases["spheres"] = scenes.with_spheres(sx, si, sf, scenes.sphere_field(40, seed=15, lo=sx.min(0), hi=sx.max(0)))
    for i, (name, (xyz, idx, _)) in enumerate(cases.items()):
        flags = flag_mix(idx.shape[0], seed=20 + i) if name != "hf" else np.full(idx.shape[0], 3, np.uint8)
        ref = yref.RefScene(xyz, idx, flags)
        bound = ref.bound()
        closest_rays, shadow_rays = ray_zoo(bound, n=3000, seed=30 + i)
        c = ref.trace_closest(closest_rays, threads=1)
        s = ref.trace_shadow(shadow_rays, threads=1)
        t = ref.trace_tshadow(shadow_rays, TSHADOW_DEPTH, threads=1)
        tree = ref.export_tree()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            xyz=xyz, idx=idx, flags=flags, bound=bound,
            closest_rays=closest_rays, closest_t=c["t"], closest_u=c["u"], closest_v=c["v"], closest_prim=c["prim"],
            shadow_rays=shadow_rays, shadow_shadowed=s["shadowed"], shadow_prim=s["prim"],
            tshadow_depth=np.int32(TSHADOW_DEPTH), tshadow_shadowed=t["shadowed"], tshadow_rgb=t["rgb"],
            tree_split=tree["split"], tree_flags=tree["flags"], tree_first_ref=tree["first_ref"], tree_refs=tree["refs"],
        )
        print(f"{name}: {idx.shape[0]} faces, {closest_rays.shape[0]} closest rays ({np.mean(c['prim'] >= 0):.2f} hit), "
              f"{shadow_rays.shape[0]} shadow rays ({s['shadowed'].mean():.2f} shadowed, tshadow {t['shadowed'].mean():.2f})")


def main_motion():
    """tests/golden/motion/motion.npz: static geometry + a Bezier motion-blur mesh + two moving instances (scenes.motion_scene),
    timed rays, and what the unmodified reference returned for them -- plus its kd-tree and bound."""
    os.makedirs(os.path.join(HERE, "motion"), exist_ok=True)
    xyz, idx, _, mo = scenes.motion_scene(n_static=1500, n_bezier=1200, n_moving=900, seed=5)
    flags = flag_mix(idx.shape[0], seed=41)
    ref = yref.RefScene(xyz, idx, flags, motion=mo)
    bound = ref.bound()
    closest_rays, shadow_rays = ray_zoo(bound, n=6000, seed=43)
    ct, st = scenes.ray_times(closest_rays.shape[0], 44), scenes.ray_times(shadow_rays.shape[0], 45)
    c = ref.trace_closest(closest_rays, threads=1, times=ct)
    s = ref.trace_shadow(shadow_rays, threads=1, times=st)
    t = ref.trace_tshadow(shadow_rays, TSHADOW_DEPTH, threads=1, times=st)
    tree = ref.export_tree()
    np.savez_compressed(
        os.path.join(HERE, "motion", "motion.npz"),
        xyz=xyz, idx=idx, flags=flags, bound=bound, **{"motion_" + k: v for k, v in mo.items()},
        closest_rays=closest_rays, closest_times=ct, closest_t=c["t"], closest_u=c["u"], closest_v=c["v"], closest_prim=c["prim"],
        shadow_rays=shadow_rays, shadow_times=st, shadow_shadowed=s["shadowed"], shadow_prim=s["prim"],
        tshadow_depth=np.int32(TSHADOW_DEPTH), tshadow_shadowed=t["shadowed"], tshadow_rgb=t["rgb"],
        tree_split=tree["split"], tree_flags=tree["flags"], tree_first_ref=tree["first_ref"], tree_refs=tree["refs"],
    )
    hit = c["prim"] >= 0
    print(f"motion: {idx.shape[0]} faces (kinds {np.bincount(mo['kind']).tolist()}), {closest_rays.shape[0]} closest rays ({hit.mean():.2f} hit, "
          f"kinds hit {np.bincount(mo['kind'][c['prim'][hit]]).tolist()}), {shadow_rays.shape[0]} shadow rays ({s['shadowed'].mean():.2f} shadowed)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "motion":
        main_motion()
    else:
        main()
        main_motion()
