"""Host kd-tree build times (libb200rt's own SAH builder, no device needed): python tools/build_time.py [--big]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libyafaray_b200 import rt, scenes  # noqa: E402

cases = [("S1M-hf", lambda: scenes.heightfield(707))]
if "--big" in sys.argv:
    cases.append(("S10M-objects", lambda: scenes.objects(10_000_000)))
for name, make in cases:
    xyz, idx, _ = make()
    t0 = time.perf_counter()
    tree = rt.host_tree(xyz, idx, rt.make_params(build_threads=0))
    print(json.dumps({"scene": name, "faces": int(idx.shape[0]), "threads": os.cpu_count(), "build_and_export_seconds": time.perf_counter() - t0,
                      "nodes": int(len(tree["a"]))}), flush=True)
