"""N>1 path on CPU: world_size-2 gloo run of the sample-index sharding + single reduce.
The per-rank "renderer" is the oracle here (no GPU); what is under test is the product's sharding /
reduce logic: the sharded result must equal the single-process render of the same samples."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, STEPS, BOUNCES = 40, 24, 3, 4


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bisemutum_engine_b200 import capi, scenes, sharding
    from oracle import oracle_py
    scene = scenes.small_test_scene()
    ctx = oracle_py.OracleContext(W, H, threads=2)
    ctx.upload_scene(scene, capi.ACCEL_MERGED)
    cam = oracle_py.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=BOUNCES)
    mine = sharding.frames_of_rank(STEPS, rank, world)
    for f in mine:
        ctx.render(cam, f, 1, st)
    sums = torch.from_numpy(ctx.resolve(1).copy())          # resolve(1) = the raw sums (alpha forced to 1)
    sharding.reduce_sums(sums, dst=0)
    if rank == 0:
        img = sharding.resolve(sums, STEPS * world)
        np.save(out_path, img.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(oracle, tmp_path):
    from bisemutum_engine_b200 import capi, scenes, sharding
    assert sorted(sharding.frames_of_rank(STEPS, 0, 2) + sharding.frames_of_rank(STEPS, 1, 2)) == list(range(2 * STEPS))
    out = str(tmp_path / "img.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    img = np.load(out)
    scene = scenes.small_test_scene()
    ctx = oracle.OracleContext(W, H)
    ctx.upload_scene(scene, capi.ACCEL_MERGED)
    ctx.render(oracle.camera_matrices(scene.camera, W, H), 0, 2 * STEPS, capi.Settings(max_bounces=BOUNCES))
    ref = ctx.resolve(2 * STEPS)
    # FP32 sums associate differently across ranks: equal to ~1e-7 relative, not bitwise
    np.testing.assert_allclose(img, ref, rtol=1e-5, atol=1e-7)
