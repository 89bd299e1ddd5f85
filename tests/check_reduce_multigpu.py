"""Multi-GPU check of the one exchange step, run under torchrun with N >= 2 GPUs:
every rank renders its block of samples, bpt_reduce sums the FP32 buffers to rank 0 over NCCL, and rank 0's image must
equal a single-GPU render of all samples to ~1e-6 (FP32 sum order differs) — and the oracle's image within 1e-4.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/check_reduce_multigpu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = pkg.load_library()
W, H, per_rank = 160, 96, 3
scene = scenes.small_test_scene()
ctx = capi.Context(lib, W, H, device=local)
ctx.upload_scene(scene, capi.ACCEL_MERGED)
cam = engine.camera_matrices(scene.camera, W, H)
st = capi.Settings(max_bounces=6)
sharding.comm_init(ctx, torch.device("cuda", local))
ctx.render(cam, sharding.first_frame(per_rank, rank, world), per_rank, st)
ctx.reduce(0)
ctx.sync()
ok = True
if rank == 0:
    img = ctx.resolve(per_rank * world)
    one = capi.Context(lib, W, H, device=local)
    one.upload_scene(scene, capi.ACCEL_MERGED)
    one.render(cam, 0, per_rank * world, st)
    ref = one.resolve(per_rank * world)
    err = float(np.abs(img - ref).max()); scale = float(np.abs(ref).max())
    print(f"reduce over {world} GPUs vs one GPU: max abs diff {err:.3e} (image scale {scale:.3f})")
    ok = err <= 2e-6 * max(scale, 1.0)
    from oracle import oracle_py
    o = oracle_py.OracleContext(W, H); o.upload_scene(scene, capi.ACCEL_MERGED)
    o.render(cam, 0, per_rank * world, st)
    oerr = float(np.abs(img - o.resolve(per_rank * world)).max())
    print(f"vs oracle: max abs diff {oerr:.3e}")
    ok = ok and oerr <= 1e-4 * max(scale, 1.0)
    # the fp16 running average must refuse the reduce
    ctx.clear_accum(); ctx.render(cam, 0, 1, capi.Settings(max_bounces=3, state_precision=capi.STATE_REFERENCE_FP16))
    try:
        ctx.reduce(0); ok = False; print("reduce of an fp16 average was accepted")
    except capi.BptError as e:
        print("fp16 reduce refused:", e)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
