"""Acceleration-structure build time: bpt_build_accel wall clock (synchronised, geometry already
resident) for the atrium (262 144 triangles, merged + wide collapse) and the 2 M-triangle BLAS x 512 instances
(two-level), plus bpt_update_tlas. One JSON line per case; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split. Not the contract bench."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--cases", default="atrium,instanced")
args = ap.parse_args()
lib = pkg.load_library()


def run(name, scene, mode):
    ctx = capi.Context(lib, 256, 256)
    ctx.upload_scene(scene, mode); ctx.sync()
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter(); ctx.build_accel(mode); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
    out = dict(case=name, triangles=scene.num_triangles, instances=len(scene.instances), accel="merged" if mode == capi.ACCEL_MERGED else "two_level",
               build_ms_min=min(ts), build_ms_all=[round(t, 3) for t in ts])
    if mode == capi.ACCEL_TWO_LEVEL:
        tt = []
        for _ in range(args.reps):
            t0 = time.perf_counter(); ctx.update_tlas(); ctx.sync(); tt.append((time.perf_counter() - t0) * 1e3)
        out["update_tlas_ms_min"] = min(tt)
    out["mtris_per_s"] = scene.num_triangles / out["build_ms_min"] / 1e3
    print(json.dumps(out), flush=True)
    ctx.close()
    return out


res = []
cases = args.cases.split(",")
if "atrium" in cases:
    atr = scenes.atrium()
    res.append(run("atrium_merged", atr, capi.ACCEL_MERGED))
    res.append(run("atrium_two_level", atr, capi.ACCEL_TWO_LEVEL))
if "instanced" in cases:
    res.append(run("instanced_2Mx512", scenes.instanced(), capi.ACCEL_TWO_LEVEL))
os.makedirs(os.path.join(pkg.REPO_ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(pkg.REPO_ROOT, "gpurun_out", "build_times.json"), "w"), indent=1)
