"""Diagnostic: first per-bounce mismatch between the CUDA path and the oracle on the atrium."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes
from oracle import oracle_py as O

scene = scenes.atrium()
W, H, B = 240, 136, 8
gpu = capi.Context(pkg.load_library(), W, H); ref = O.OracleContext(W, H)
gpu.upload_scene(scene, capi.ACCEL_MERGED); ref.upload_scene(scene, capi.ACCEL_MERGED)
cam = engine.camera_matrices(scene.camera, W, H)
st = capi.Settings(max_bounces=B)
gpu.debug_capture(True); ref.debug_capture(True)
gpu.render(cam, 0, 1, st); ref.render(cam, 0, 1, st)
a, b = gpu.resolve(1), ref.resolve(1)
print("image mismatches:", (a != b).any(-1).sum(), "max abs", np.abs(a - b).max())
for bounce in range(1, B):
    qa, qb = gpu.read_queue(bounce, 0), ref.read_queue(bounce, 0)
    o = np.argsort(qa["pixels"], kind="stable")
    print("bounce", bounce, "n", len(qa["pixels"]), len(qb["pixels"]))
    if len(qa["pixels"]) != len(qb["pixels"]) or (qa["pixels"][o] != qb["pixels"]).any():
        print("  queue pixel sets differ"); 
        sa, sb = set(qa["pixels"].tolist()), set(qb["pixels"].tolist())
        print("  only gpu", sorted(sa - sb)[:10], "only ref", sorted(sb - sa)[:10])
        break
    bad = np.zeros(len(o), bool)
    for f in ("t", "u", "v", "instance", "primitive"):
        bad |= qa["hits"][f][o] != qb["hits"][f]
    print("  hit mismatches", bad.sum())
    if bad.any():
        for i in np.nonzero(bad)[0][:8]:
            print("   pixel", qb["pixels"][i], "gpu", qa["hits"][o][i], "ref", qb["hits"][i])
        break
    sa, sb = gpu.read_queue(bounce, 1), ref.read_queue(bounce, 1)
    print("  shadow", len(sa["pixels"]), len(sb["pixels"]))
