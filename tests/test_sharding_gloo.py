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


def test_wave_schedule_of_a_finite_job():
    """bench.py's e2e arm: shrinking waves that add up to the job exactly; no wave larger than the library's; the last one is short."""
    sys.path.insert(0, ROOT)
    from bisemutum_engine_b200 import sharding
    assert sharding.wave_schedule(20, 32) == [16, 4]
    assert sharding.wave_schedule(128, 32) == [32, 32, 32, 26, 5, 1]
    assert sharding.wave_schedule(1, 32) == [1] and sharding.wave_schedule(0, 32) == []
    for frames in range(1, 300):
        for wave in (1, 2, 8, 32):
            sched = sharding.wave_schedule(frames, wave)
            assert sum(sched) == frames and all(1 <= w <= wave for w in sched)
            assert sched == sorted(sched, reverse=True)


def test_strong_scaling_split_tiles_the_job():
    """sharding.split_frames (bench.py --scaling strong; BASELINE configs[3]: 64 spp split across 8 GPUs): contiguous blocks that tile the job."""
    from bisemutum_engine_b200 import sharding
    for total in (8, 20, 64, 67, 256):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharding.split_frames(total, r, world, first=5) for r in range(world)]
            frames = [f for first, n in blocks for f in range(first, first + n)]
            assert frames == list(range(5, 5 + total))
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1


# ---- DDGI update sharded by probe index + all-gather of the per-ray results -------------------------------------
PROBES, RAYS = (3, 2, 3), 16


def _probe_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bisemutum_engine_b200 import capi, scenes, sharding
    from oracle import oracle_py
    scene = scenes.small_test_scene()
    ctx = oracle_py.OracleContext(8, 8, threads=2)
    ctx.upload_scene(scene, capi.ACCEL_MERGED)
    vol = scenes.probe_volume(scene, PROBES, RAYS, ray_length=50.0); tab = scenes.ddgi_sample_randoms()
    n = PROBES[0] * PROBES[1] * PROBES[2]
    first, count = sharding.probe_range(n, rank, world)
    mine = torch.from_numpy(ctx.trace_probes_range(vol, tab, 5, 2, first, count))
    rays = sharding.allgather_probe_rays(mine, n, RAYS)
    irr, vis = ctx.blend_probes(vol, tab, 5, rays.numpy())            # every rank blends all probes from the gathered rays
    np.savez(out_path + f".{rank}.npz", rays=rays.numpy(), irr=irr, vis=vis)
    dist.destroy_process_group()


def test_two_rank_probe_sharding_equals_single_process(oracle, tmp_path):
    from bisemutum_engine_b200 import capi, scenes, sharding
    n = PROBES[0] * PROBES[1] * PROBES[2]                              # 18 probes over 2 and 4 ranks: even and uneven blocks
    for world in (2, 4, 5):
        r = [sharding.probe_range(n, k, world) for k in range(world)]
        assert r[0][0] == 0 and sum(c for _, c in r) == n and all(r[k][0] + r[k][1] == r[k + 1][0] for k in range(world - 1))
    out = str(tmp_path / "probes")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_probe_worker, args=(2, port, out), nprocs=2, join=True)
    scene = scenes.small_test_scene()
    ctx = oracle.OracleContext(8, 8); ctx.upload_scene(scene, capi.ACCEL_MERGED)
    vol = scenes.probe_volume(scene, PROBES, RAYS, ray_length=50.0); tab = scenes.ddgi_sample_randoms()
    want = ctx.trace_probes(vol, tab, 5, 2)
    irr, vis = ctx.blend_probes(vol, tab, 5, want)
    for rank in (0, 1):
        got = np.load(out + f".{rank}.npz")
        np.testing.assert_array_equal(got["rays"].view(np.uint32), want.view(np.uint32))          # bit-identical: keys are global probe indices
        np.testing.assert_array_equal(got["irr"], irr); np.testing.assert_array_equal(got["vis"], vis)
    # a range in the middle of the volume equals the same rows of the full call; an out-of-range request is refused
    part = ctx.trace_probes_range(vol, tab, 5, 2, 7, 4)
    np.testing.assert_array_equal(part.view(np.uint32), want[7 * RAYS:11 * RAYS].view(np.uint32))
    import pytest
    with pytest.raises(capi.BptError):
        ctx.trace_probes_range(vol, tab, 5, 2, n - 1, 2)
