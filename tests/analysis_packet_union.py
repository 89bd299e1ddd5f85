"""Analysis script (not a test; lives under tests/ because it uses the CPU checker): how many nodes does a PACKET of 32 camera rays
(one 8x4 pixel tile) visit, against the sum of its rays' own traversals? Walks the binary BVH the checker builds for configs[1] with exact
per-ray box tests and each ray's final hit distance as its cull bound. Result quoted in DESIGN.md section 5: 37.9 nodes / 2.3 leaves per
ray, 44.6 / 5.5 per tile — the measurement behind k_trace_packet.   python tests/analysis_packet_union.py"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes
import oracle_py
W, H = 1920, 1080
scene = scenes.atrium()
ora = oracle_py.OracleContext(W, H)
ora.upload_scene(scene, capi.ACCEL_MERGED)
bvh = ora.read_bvh(0)
nodes = bvh["nodes"]; root = bvh["root"]; n = bvh["n"]
print("tris", n, "root", root)
cam = engine.camera_matrices(scene.camera, W, H)
iv = np.array(cam.matrix_inv_view, dtype=np.float64).reshape(4, 4).T   # column-major -> M[r][c]
ip = np.array(cam.matrix_inv_proj, dtype=np.float64).reshape(4, 4).T
def rays_for(px, py):
    u = (px + 0.5) / W; v = (py + 0.5) / H
    p = np.stack([2 * u - 1, 1 - 2 * v, np.ones_like(u), np.ones_like(u)], -1)
    dl = (p @ ip.T)[..., :3]; dl /= np.linalg.norm(dl, axis=-1, keepdims=True)
    d = dl @ iv[:3, :3].T; d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(iv[:3, 3], d.shape)
    return o, d
rng = np.random.default_rng(1)
c0lo = np.stack([nodes["c0_lo_x"], nodes["c0_lo_y"], nodes["c0_lo_z"]], -1).astype(np.float64)
c0hi = np.stack([nodes["c0_hi_x"], nodes["c0_hi_y"], nodes["c0_hi_z"]], -1).astype(np.float64)
c1lo = np.stack([nodes["c1_lo_x"], nodes["c1_lo_y"], nodes["c1_lo_z"]], -1).astype(np.float64)
c1hi = np.stack([nodes["c1_hi_x"], nodes["c1_hi_y"], nodes["c1_hi_z"]], -1).astype(np.float64)
ch0 = nodes["child0"]; ch1 = nodes["child1"]
tot_U = tot_N = tot_L = tot_LU = 0; tiles = 0
for t in range(300):
    tx = rng.integers(0, W // 8); ty = rng.integers(0, H // 4)
    px, py = np.meshgrid(np.arange(8) + tx * 8, np.arange(4) + ty * 4)
    o, d = rays_for(px.ravel().astype(np.float64), py.ravel().astype(np.float64))
    rays = np.zeros(32, capi.RAY); rays["origin"] = o; rays["direction"] = d; rays["tmin"] = 0.001; rays["tmax"] = 100.0
    hits = ora.trace_rays(rays)
    tb = np.where(hits["t"] >= 0, hits["t"], 100.0).astype(np.float64) * 1.00001
    idir = 1.0 / d; ood = o * idir
    def box(lo, hi, mask):
        t0 = lo[None, :] * idir - ood; t1 = hi[None, :] * idir - ood
        tn = np.maximum(np.minimum(t0, t1).max(-1), 0.001); tf = np.minimum(np.maximum(t0, t1).min(-1), tb)
        return mask & (tn <= tf)
    stack = [(root, np.ones(32, bool))]
    U = N = L = LU = 0
    while stack:
        node, mask = stack.pop()
        if node < 0:
            LU += 1; L += mask.sum(); continue
        U += 1; N += mask.sum()
        m0 = box(c0lo[node], c0hi[node], mask); m1 = box(c1lo[node], c1hi[node], mask)
        if m1.any(): stack.append((ch1[node], m1))
        if m0.any(): stack.append((ch0[node], m0))
    tot_U += U; tot_N += N; tot_L += L; tot_LU += LU; tiles += 1
print("tiles", tiles, "union nodes per tile", tot_U / tiles, "per-ray nodes", tot_N / tiles / 32, "union leaves per tile", tot_LU / tiles, "per-ray leaves", tot_L / tiles / 32)
