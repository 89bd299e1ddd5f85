"""Post-process pass (bloom + output, csrc/post.cu) timing: CUDA events around bpt_post_process_device, device-resident input,
against the HBM roofline. Algorithmic bytes = the compulsory traffic of the 6-launch plan: the FP32 sum buffer
is read twice (level 1, output), the output written once, every rgba16_sfloat target written once and read by its consumers:
60.75 B per pixel with bloom, 32 B per pixel without. One JSON line per case. Not the contract bench."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi

lib = pkg.load_library()
peaks = {}
try:
    peaks = json.load(open(os.path.join(pkg.REPO_ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass


def find_hbm(d):
    for k, v in d.items():
        if isinstance(v, dict):
            r = find_hbm(v)
            if r:
                return r
        elif isinstance(v, (int, float)) and "hbm" in k.lower() and "gb" in k.lower():
            return float(v)
    return None


HBM = find_hbm(peaks) or 6545.9
res = []
for (w, h) in ((1920, 1080), (3840, 2160)):
    ctx = capi.Context(lib, w, h)
    stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1)
    img = rng.random((h, w, 4), dtype=np.float32); img[rng.random((h, w)) < 0.02, :3] += 30.0
    ctx.upload_accum(img)
    out = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")            # > L2 (126 MB): flushed between timed iterations
    for bloom in (True, False):
        st = capi.PostSettings(bloom, 1.5, 0.5)
        for _ in range(5):
            ctx.post_process_device(st, 1, out.data_ptr())
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.post_process_device(st, 1, out.data_ptr()); e1.record(stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        bpp = 60.75 if bloom else 32.0
        gbs = w * h * bpp / ms / 1e6
        r = dict(case=f"post_{w}x{h}_{'bloom' if bloom else 'output_only'}", ms=ms, ms_min=min(ts), algorithmic_bytes_per_pixel=bpp,
                 achieved_gb_s=gbs, hbm_peak_gb_s=HBM, frac=gbs / HBM, launches=6 if bloom else 1, l2="flushed between iterations")
        print(json.dumps(r), flush=True); res.append(r)
    ctx.close()
os.makedirs(os.path.join(pkg.REPO_ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(pkg.REPO_ROOT, "gpurun_out", "post_times.json"), "w"), indent=1)
