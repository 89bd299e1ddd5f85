"""Multi-GPU check of the DDGI update's exchange step, run under torchrun with N >= 2 GPUs: every rank traces its
probe range (bpt_trace_probes_range), one NCCL all-gather assembles the per-ray results (sharding.allgather_probe_rays), every rank
blends; the atlases and rays must equal a single-GPU update bit for bit (all keys are global probe indices).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/check_probes_sharded_multigpu.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, scenes, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = pkg.load_library()
scene = scenes.atrium()
ctx = capi.Context(lib, 256, 256, device=local)
ctx.upload_scene(scene, capi.ACCEL_MERGED)
counts, rays_per_probe = (32, 32, 16), 256                        # BASELINE configs[4]
vol = scenes.probe_volume(scene, counts, rays_per_probe); tab = scenes.ddgi_sample_randoms()
n = counts[0] * counts[1] * counts[2]
first, count = sharding.probe_range(n, rank, world)
# the per-ray results stay on the device: trace into a torch tensor, NCCL all-gather device to device, blend from the gathered tensor
mine = torch.empty((count * rays_per_probe, 4), dtype=torch.float32, device="cuda")
ctx.trace_probes_range_into(vol, tab, 100, 2, first, count, mine.data_ptr())        # warm-up: kernels, NCCL communicator
sharding.allgather_probe_rays(mine, n, rays_per_probe)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
ctx.trace_probes_range_into(vol, tab, 0, 2, first, count, mine.data_ptr())
rays = sharding.allgather_probe_rays(mine, n, rays_per_probe)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
irr, vis = ctx.blend_probes_from_device(vol, tab, 0, rays.contiguous().data_ptr())
ok = True
if rank == 0:
    one = torch.empty((n * rays_per_probe, 4), dtype=torch.float32, device="cuda")
    t1 = time.perf_counter()
    ctx.trace_probes_range_into(vol, tab, 0, 2, 0, n, one.data_ptr()); torch.cuda.synchronize()
    dt1 = time.perf_counter() - t1
    full = one.cpu().numpy()
    firr, fvis = ctx.blend_probes(vol, tab, 0, full)
    same = np.array_equal(rays.cpu().numpy().view(np.uint32), full.view(np.uint32)) and np.array_equal(irr, firr) and np.array_equal(vis, fvis)
    print(f"DDGI update over {world} GPUs: rays + atlases bit-identical to one GPU: {same}; trace + gather {dt * 1e3:.1f} ms vs one GPU trace {dt1 * 1e3:.1f} ms "
          f"({n * rays_per_probe} rays x 2 bounces, results device-resident)")
    ok = same
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
