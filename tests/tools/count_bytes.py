import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import time, numpy as np, json
from libyafaray_b200 import scenes, rt
from oracle import yref, kdo
from tests.helpers import host_tree_as_oracle_tree
xyz, idx, fl = scenes.heightfield(707)
t0=time.time(); s = yref.RefScene(xyz, idx, fl); print("ref build", time.time()-t0, s.build_seconds)
tr = s.export_tree()
print("ref nodes", len(tr['flags']), "refs", len(tr['refs']))
o = kdo.Oracle(xyz, idx, fl, tree=tr)
rays = scenes.rays_incoherent(2_000_000, seed=12345)
srays = scenes.rays_shadow(2_000_000, seed=12346, t_max=0.25)
c = o.trace_closest(rays, threads=8, counters=True)
sh = o.trace_shadow(srays, threads=8, counters=True)
print("closest", c['counters'].per_ray(), c['counters'].bytes_per_ray(48), "hit", (c['prim']>=0).mean())
print("shadow", sh['counters'].per_ray(), sh['counters'].bytes_per_ray(36), "shadowed", sh['shadowed'].mean())
t0=time.time(); r = s.trace_closest(rays, threads=8); print("ref closest Mrays/s", 2/r['seconds'])
r2 = s.trace_shadow(srays, threads=8); print("ref shadow Mrays/s", 2/r2['seconds'])
t0=time.time(); t = rt.host_tree(xyz, idx); print("my build", time.time()-t0, "nodes", len(t['a']), "refs", len(t['refs']))
om = kdo.Oracle(xyz, idx, fl, tree=host_tree_as_oracle_tree(t), bound=t['bound'])
cm = om.trace_closest(rays, threads=8, counters=True)
print("mine closest", cm['counters'].per_ray(), "mismatch", (cm['prim']!=c['prim']).sum(), (cm['t']!=c['t']).sum())
