"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json): integer work (Morton codes, sort permutation, BVH topology, queue contents) is
bit-exact; per-sample radiance within 1e-4 relative (the single-light scenes are in fact bit-equal
because both sides follow one numeric contract — DESIGN.md "Numerics").
"""
import numpy as np
import pytest

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes

pytestmark = pytest.mark.gpu

MODES = [capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED]


@pytest.fixture(scope="module")
def lib():
    return pkg.load_library()


def make_pair(lib, oracle, scene, W, H, mode):
    gpu = capi.Context(lib, W, H)
    ref = oracle.OracleContext(W, H)
    gpu.upload_scene(scene, mode)
    ref.upload_scene(scene, mode)
    return gpu, ref


def assert_bvh_equal(a, b):
    assert a["n"] == b["n"] and a["root"] == b["root"]
    np.testing.assert_array_equal(a["morton"], b["morton"])
    np.testing.assert_array_equal(a["prims"], b["prims"])
    for f in ("child0", "child1", "parent"):
        np.testing.assert_array_equal(a["nodes"][f], b["nodes"][f])
    for f in a["nodes"].dtype.names[:12]:
        np.testing.assert_array_equal(a["nodes"][f], b["nodes"][f])     # value equality (-0 == +0)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("scene_fn", [scenes.small_test_scene, scenes.cornell_box])
def test_bvh_bit_exact(lib, oracle, scene_fn, mode):
    scene = scene_fn()
    gpu, ref = make_pair(lib, oracle, scene, 16, 16, mode)
    nb = len(scene.blas) if mode == capi.ACCEL_TWO_LEVEL else 1
    for b in range(nb):
        assert_bvh_equal(gpu.read_bvh(b), ref.read_bvh(b))
    if mode == capi.ACCEL_TWO_LEVEL:
        assert_bvh_equal(gpu.read_bvh(capi.BVH_TLAS), ref.read_bvh(capi.BVH_TLAS))


def random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.bounds
    rays = np.zeros(n, capi.RAY)
    rays["origin"] = rng.uniform(lo - 0.5, hi + 0.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"] = 0.001
    rays["tmax"] = rng.choice([100.0, 3.0, 0.5], n).astype(np.float32)
    return rays


@pytest.mark.parametrize("mode", MODES)
def test_trace_rays_parity(lib, oracle, mode):
    scene = scenes.small_test_scene()
    gpu, ref = make_pair(lib, oracle, scene, 16, 16, mode)
    rays = random_rays(scene, 20000, 11)
    a, b = gpu.trace_rays(rays, 5), ref.trace_rays(rays, 5)
    for f in ("t", "u", "v", "instance", "primitive"):
        np.testing.assert_array_equal(a[f], b[f])
    assert (a["t"] >= 0).mean() > 0.2
    np.testing.assert_array_equal(gpu.trace_shadow_rays(rays, 5), ref.trace_shadow_rays(rays, 5))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("scene_fn,W,H,bounces", [(scenes.small_test_scene, 96, 64, 6), (scenes.cornell_box, 128, 128, 5)])
def test_per_sample_radiance(lib, oracle, scene_fn, W, H, bounces, mode):
    scene = scene_fn()
    gpu, ref = make_pair(lib, oracle, scene, W, H, mode)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=bounces)
    for frame in (0, 7):
        gpu.clear_accum(); ref.clear_accum()
        gpu.render(cam, frame, 1, st); ref.render(cam, frame, 1, st)
        a, b = gpu.resolve(1), ref.resolve(1)
        assert np.isfinite(a).all()
        scale = np.maximum(np.abs(b), 1e-3)
        assert (np.abs(a - b) / scale).max() <= 1e-4          # BASELINE.json: 1e-4 relative per sample
        np.testing.assert_array_equal(a, b)                   # one light per vertex → identical add order


@pytest.mark.parametrize("mode", MODES)
def test_queue_contents_and_counters(lib, oracle, mode):
    scene = scenes.small_test_scene()
    W, H, B = 64, 48, 6
    gpu, ref = make_pair(lib, oracle, scene, W, H, mode)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=B)
    gpu.debug_capture(True); ref.debug_capture(True)
    gpu.render(cam, 3, 1, st); ref.render(cam, 3, 1, st)
    for bounce in range(1, B):
        qa, qb = gpu.read_queue(bounce, 0), ref.read_queue(bounce, 0)
        order = np.argsort(qa["pixels"], kind="stable")          # queue ORDER is free, CONTENT is not
        np.testing.assert_array_equal(qa["pixels"][order], qb["pixels"])
        for f in ("t", "u", "v", "instance", "primitive"):
            np.testing.assert_array_equal(qa["hits"][f][order], qb["hits"][f])
        sa, sb = gpu.read_queue(bounce, 1), ref.read_queue(bounce, 1)
        ka = np.sort(sa["pixels"].astype(np.uint64) << 32 | sa["lights"])
        kb = np.sort(sb["pixels"].astype(np.uint64) << 32 | sb["lights"])
        np.testing.assert_array_equal(ka, kb)
    ca, cb = gpu.counters(), ref.counters()
    assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays
    assert list(ca.extend_rays_per_bounce) == list(cb.extend_rays_per_bounce)
    assert list(ca.shadow_rays_per_bounce) == list(cb.shadow_rays_per_bounce)


def test_atrium_full_size_bvh_and_render(lib, oracle):
    """BASELINE configs[1] geometry at full size: bit-exact BVH; radiance parity at reduced resolution;
    at 1920x1080 the size-independent property sum(s0) + sum(s1) == sum(s0, s1) and counter sanity."""
    scene = scenes.atrium()
    assert scene.num_triangles == scenes.ATRIUM_TRIANGLES
    W, H = 240, 136
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    assert_bvh_equal(gpu.read_bvh(0), ref.read_bvh(0))
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=8)
    gpu.render(cam, 0, 1, st); ref.render(cam, 0, 1, st)
    a, b = gpu.resolve(1), ref.resolve(1)
    np.testing.assert_array_equal(a, b)
    ca, cb = gpu.counters(), ref.counters()
    assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays
    gpu.close()
    W, H = 1920, 1080
    big = capi.Context(lib, W, H)
    big.upload_scene(scene, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    big.render(cam, 0, 1, st); s0 = big.resolve(1)
    big.clear_accum(); big.render(cam, 1, 1, st); s1 = big.resolve(1)
    big.clear_accum(); big.render(cam, 0, 2, st); s01 = big.resolve(1)
    assert np.isfinite(s01).all() and s01[..., :3].mean() > 0
    # linearity of the FP32 sum buffer (exact up to the different association of the adds)
    np.testing.assert_allclose((s0 + s1)[..., :3], s01[..., :3], rtol=1e-5, atol=1e-7)
    c = big.counters()
    assert c.samples == 4 * W * H and c.extend_rays_per_bounce[1] == 4 * W * H
    assert all(c.extend_rays_per_bounce[i] >= c.extend_rays_per_bounce[i + 1] for i in range(1, 8))


def test_host_pass_accumulates_like_reference(lib, oracle):
    """PathTracingPass::render history rule (path_tracing.cpp:231-246): consecutive frames with an
    unchanged camera accumulate; a camera move resets. Image == oracle's mean of the same frames."""
    scene = scenes.small_test_scene()
    W, H = 64, 48
    r = engine.Renderer(W, H)
    r.set_scene(scene, capi.ACCEL_MERGED)
    n = 0
    for _ in range(3):
        n = r.frame(max_bounces=4)
    assert n == 3
    img = r.image(n)
    ref = oracle.OracleContext(W, H)
    ref.upload_scene(scene, capi.ACCEL_MERGED)
    ref.render(engine.camera_matrices(scene.camera, W, H), 0, 3, capi.Settings(max_bounces=4))
    np.testing.assert_array_equal(img, ref.resolve(3))
    cam2 = dict(scene.camera); cam2["position"] = (0.5, 2.2, 6.5)
    r.set_camera(cam2)
    assert r.frame(max_bounces=4) == 1          # history invalidated by the camera change
    r.close()
