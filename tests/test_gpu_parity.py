"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json): integer work (Morton codes, sort permutation, BVH topology, queue contents) is
bit-exact; per-sample radiance within 1e-4 relative (the single-light scenes are in fact bit-equal
because both sides follow one numeric contract — DESIGN.md "Numerics").
"""
import os

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


def _tiny_scene(num_tris, num_instances=1):
    b = scenes.SceneBuilder(f"tiny_{num_tris}x{num_instances}")
    m = b.add_material((0.8, 0.7, 0.6), roughness=0.4)
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.2], [2, 0, 0.1], [2, 1, 0.3]], np.float32)
    N = np.tile(np.array([[0, 0, 1]], np.float32), (6, 1)); T = np.tile(np.array([[1, 0, 0, 1]], np.float32), (6, 1))
    idx = np.array([[0, 1, 2], [1, 3, 2], [1, 4, 3], [4, 5, 3]], np.uint32)[:num_tris]
    mesh = b.add_mesh((P, N, T, P[:, :2].copy(), idx))
    for k in range(num_instances):
        b.add_drawable(mesh, m, scenes.translate(0.0, 1.3 * k, -0.4 * k))
    return b.finish(dir_lights=scenes.dir_light((0.2, 0.3, 1.0)), sky_faces=scenes.procedural_sky(8, (0.2, 0.3, 1.0)),
                    camera=dict(position=(0.8, 0.9, 4), front_dir=(0, 0, -1), up_dir=(0, 1, 0), yfov=40, near_z=0.01, far_z=100),
                    bounds=(np.array([-1.0, -1.0, -2.0]), np.array([3.0, 4.0, 1.0])))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tris,insts", [(1, 1), (2, 1), (1, 2), (3, 1), (4, 3)])
def test_tiny_scenes_and_edge_inputs(lib, oracle, mode, tris, insts):
    """The degenerate ends of every structure: a BVH that is a single leaf (no node, no wide node), two leaves (one node), a TLAS of one
    or two instances — through the single-block build, the wide collapse and both traversal kernels — and a 1 x 1 image."""
    scene = _tiny_scene(tris, insts)
    gpu, ref = make_pair(lib, oracle, scene, 24, 16, mode)
    nb = len(scene.blas) if mode == capi.ACCEL_TWO_LEVEL else 1
    for b in range(nb):
        assert_bvh_equal(gpu.read_bvh(b), ref.read_bvh(b))
    if mode == capi.ACCEL_TWO_LEVEL:
        assert_bvh_equal(gpu.read_bvh(capi.BVH_TLAS), ref.read_bvh(capi.BVH_TLAS))
    rays = random_rays(scene, 2000, 5)
    rays["origin"][:500] = (0.4, 0.3, 2.0); rays["direction"][:500] = (0, 0, -1)
    rays["origin"][:500, :2] += np.random.default_rng(1).uniform(-0.5, 2.0, (500, 2)).astype(np.float32)
    a, b_ = gpu.trace_rays(rays, 3), ref.trace_rays(rays, 3)
    for f in ("t", "u", "v", "instance", "primitive"):
        np.testing.assert_array_equal(a[f], b_[f], err_msg=f)
    assert (a["t"] >= 0).sum() > 20
    np.testing.assert_array_equal(gpu.trace_shadow_rays(rays, 3), ref.trace_shadow_rays(rays, 3))
    cam = engine.camera_matrices(scene.camera, 24, 16)
    st = capi.Settings(max_bounces=4)
    gpu.render(cam, 0, 3, st); ref.render(cam, 0, 3, st)
    np.testing.assert_array_equal(gpu.resolve(3), ref.resolve(3))
    assert gpu.counters().extend_rays == ref.counters().extend_rays and gpu.counters().shadow_rays == ref.counters().shadow_rays
    gpu.close(); ref.close()
    g1 = capi.Context(lib, 1, 1); r1 = oracle.OracleContext(1, 1)                   # a 1 x 1 image
    g1.upload_scene(scene, mode); r1.upload_scene(scene, mode)
    c1 = engine.camera_matrices(scene.camera, 1, 1)
    g1.render(c1, 7, 2, st); r1.render(c1, 7, 2, st)
    np.testing.assert_array_equal(g1.resolve(2), r1.resolve(2))
    np.testing.assert_array_equal(g1.post_process(capi.PostSettings(True, 0.1, 0.5), 2), oracle.post_process_image(r1.resolve(2), capi.PostSettings(True, 0.1, 0.5)))
    g1.close(); r1.close()


def test_error_paths_on_the_device(lib):
    ctx = capi.Context(lib, 8, 8)
    with pytest.raises(capi.BptError):
        ctx.build_accel(capi.ACCEL_MERGED)                                          # nothing uploaded
    with pytest.raises(capi.BptError):
        ctx.render(capi.Camera(), 0, 1, capi.Settings())                            # render before build
    with pytest.raises(capi.BptError):
        ctx.precompute_sky_ibl(capi.SkyIblDesc(specular_size=4, specular_levels=5))  # last mip would have no texel
    scene = _tiny_scene(2)
    ctx.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
    with pytest.raises(capi.BptError):
        ctx.render(engine.camera_matrices(scene.camera, 8, 8), 0, 1, capi.Settings(nee_mode=7))     # unknown mode switch
    with pytest.raises(capi.BptError):
        ctx.post_process(capi.PostSettings(True, float("nan"), 0.5), 1)
    ctx.render(engine.camera_matrices(scene.camera, 8, 8), 0, 1, capi.Settings(max_bounces=0))      # clamped to [2, 16] (path_tracing.cpp:290)
    c = ctx.counters()
    assert c.extend_rays_per_bounce[1] == 64 and c.extend_rays_per_bounce[2] == 0
    ctx.close()


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


def _tile_mask(W, H, stride, offset=0, tile=16):
    """Pixels the oracle renders under obpt_set_tile_sample(stride, offset): every stride-th 16x16 tile in row-major tile order."""
    tx = (W + tile - 1) // tile
    ys, xs = np.mgrid[0:H, 0:W]
    return ((ys // tile) * tx + xs // tile) % stride == offset


def test_config2_full_resolution_radiance(lib, oracle):
    """BASELINE configs[1] AT ITS OWN SIZE: the atrium at 1920x1080, depth 8, one sample of the full frame — every pixel of the CUDA
    image equals the oracle's (one shadow ray per vertex: the FP32 sums are bit-identical), and so do the ray counts per bounce."""
    scene = scenes.atrium()
    W, H = 1920, 1080
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=8)
    gpu.render(cam, 7, 1, st); ref.render(cam, 7, 1, st)
    a, b = gpu.resolve(1), ref.resolve(1)
    assert np.isfinite(b).all() and b[..., :3].mean() > 0.01
    np.testing.assert_array_equal(a, b)
    ca, cb = gpu.counters(), ref.counters()
    assert list(ca.extend_rays_per_bounce) == list(cb.extend_rays_per_bounce) and list(ca.shadow_rays_per_bounce) == list(cb.shadow_rays_per_bounce)
    assert ca.extend_rays_per_bounce[1] == W * H
    gpu.close(); ref.close()


def test_config3_full_resolution_radiance_on_sampled_tiles(lib, oracle):
    """BASELINE configs[2] at 1920x1080: 64 point/spot + 16 LTC rect lights, depth 3. The oracle renders every 8th 16x16 tile (the full
    frame is ~50 M shadow rays per sample); on those tiles the CUDA image of the FULL frame agrees to 1e-4 (up to 80 unordered
    shadow-ray atomics per pixel)."""
    luts = scenes.load_ltc_luts(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ltc_luts.npz"))
    scene = scenes.mixed_lights(luts)
    W, H = 1920, 1080
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=3)
    ref.set_tile_sample(8, 3)
    gpu.render(cam, 2, 1, st); ref.render(cam, 2, 1, st)
    a, b = gpu.resolve(1), ref.resolve(1)
    m = _tile_mask(W, H, 8, 3)
    assert m.mean() > 0.1 and (b[m][:, :3].max(axis=1) > 0).mean() > 0.5 and not b[~m][:, :3].any()
    err = np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-3)
    assert err.max() <= 1e-4, err.max()
    gpu.close(); ref.close()


def test_config4_full_resolution_radiance_on_sampled_tiles(lib, oracle):
    """BASELINE configs[3] at 3840x2160: 2 M-triangle mesh x 512 instances, two-level BVH, depth 8. Every 64th tile through the oracle;
    bit-identical radiance there (one directional light: one shadow ray per vertex)."""
    scene = scenes.instanced()
    W, H = 3840, 2160
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_TWO_LEVEL)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=8, ray_length=1000.0)
    ref.set_tile_sample(64, 5)
    gpu.render(cam, 1, 1, st); ref.render(cam, 1, 1, st)
    a, b = gpu.resolve(1), ref.resolve(1)
    m = _tile_mask(W, H, 64, 5)
    assert (b[m][:, :3].max(axis=1) > 0).mean() > 0.5
    np.testing.assert_array_equal(a[m], b[m])
    gpu.close(); ref.close()


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
    # OutputData.depth / .gbuffer of the pass (path_tracing.cpp:482-487) through the host mirror; asking for them must not
    # disturb the accumulation (the prefetched samples are re-traced)
    depth, g = r.primary_outputs(max_bounces=4)
    rdepth, rg = ref.render_primary(engine.camera_matrices(scene.camera, W, H), 3, capi.Settings(max_bounces=4))
    np.testing.assert_array_equal(depth, rdepth); np.testing.assert_array_equal(g["normal_roughness"], rg["normal_roughness"])
    assert r.frame(max_bounces=4) == 4
    ref.render(engine.camera_matrices(scene.camera, W, H), 3, 1, capi.Settings(max_bounces=4))
    np.testing.assert_array_equal(r.image(4), ref.resolve(4))
    # the step after the pass (basic.cpp:228-231): PostProcessPass::render on the accumulated colour, through the host mirror
    np.testing.assert_array_equal(r.post_process(True, 0.4, 0.5), oracle.post_process_image(ref.resolve(4), capi.PostSettings(True, 0.4, 0.5)))
    np.testing.assert_array_equal(r.post_process(False), oracle.post_process_image(ref.resolve(4), capi.PostSettings(False)))
    cam2 = dict(scene.camera); cam2["position"] = (0.5, 2.2, 6.5)
    r.set_camera(cam2)
    assert r.frame(max_bounces=4) == 1          # history invalidated by the camera change
    r.close()


def test_pass_writes_output_color_in_place(lib):
    """With a colour target bound (the device memory behind OutputData.color), the pass folds the frame's sample into the history and writes
    the rgba16_sfloat colour in ONE launch (bpt_accumulate_ahead_rgba16f): same bits as accumulate + bpt_resolve_device_rgba16f, frame by
    frame, and the FP32 history is the same too."""
    import torch
    scene = scenes.small_test_scene()
    W, H = 72, 40
    outs = []
    for fused in (False, True):
        r = engine.Renderer(W, H)
        r.set_scene(scene, capi.ACCEL_MERGED)
        r.set_prefetch(3)
        target = torch.zeros(H, W, 4, dtype=torch.float16, device="cuda")
        frames = []
        for f in range(5):                                   # two waves (3 + 2 of the next)
            if fused:
                r.set_color_target(target.data_ptr())
            n = r.frame(max_bounces=4)
            assert n == f + 1
            if not fused:
                r.ctx.resolve_device_rgba16f(n, target.data_ptr())
            r.ctx.sync()
            frames.append(target.cpu().numpy().view(np.uint16).copy())
        outs.append((frames, r.image(5)))
        r.set_color_target(0)
        assert r.frame(max_bounces=4) == 6                   # unbound again: plain accumulate
        r.close()
    for a, b in zip(outs[0][0], outs[1][0]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    assert outs[1][0][-1][..., 3].min() == 0x3c00 and outs[1][0][-1][..., :3].any()


def test_scene_change_invalidates_prefetched_samples(lib, oracle):
    """the pass traces a wave of samples ahead while the camera stands still. A light, sky or instance change between two
    frames must show in the very next frame, as in the reference (which shades every frame with the current parameters): the samples
    traced ahead are dropped and re-traced. An unchanged per-frame light upload (PathTracingPass::update_params) keeps them."""
    import copy
    scene = scenes.small_test_scene()
    W, H = 64, 48
    st = capi.Settings(max_bounces=4)
    cam = engine.camera_matrices(scene.camera, W, H)
    r = engine.Renderer(W, H)
    r.set_scene(scene, capi.ACCEL_MERGED)
    r.set_prefetch(8)
    ref = oracle.OracleContext(W, H)
    ref.upload_scene(scene, capi.ACCEL_MERGED)
    for f in range(2):
        r.ctx.upload_lights(scene)                           # same bytes every frame: nothing is dropped
        r.frame(max_bounces=4)
    assert r.ctx.pending_ahead()[0] == 6
    ref.render(cam, 0, 2, st)
    dimmed = copy.copy(scene)
    dimmed.dir_lights = scene.dir_lights.copy()
    dimmed.dir_lights["emission"] *= 0.25
    r.ctx.upload_lights(dimmed)                              # frame 2 onwards sees the dimmed light
    assert r.ctx.pending_ahead()[0] == 0
    ref.upload_lights(dimmed)
    for f in range(2):
        r.frame(max_bounces=4)
    ref.render(cam, 2, 2, st)
    np.testing.assert_array_equal(r.image(4), ref.resolve(4))
    r.ctx.update_sky_params(scene.sky_transform, np.float32(scene.sky_color) * 0.5)
    assert r.ctx.pending_ahead()[0] == 0
    ref.update_sky_params(scene.sky_transform, np.float32(scene.sky_color) * 0.5)
    r.frame(max_bounces=4); ref.render(cam, 4, 1, st)
    np.testing.assert_array_equal(r.image(5), ref.resolve(5))
    # a light-count change re-sizes the wave buffers: whatever was traced ahead is gone, never summed from a fresh allocation
    more = copy.copy(dimmed)
    more.dir_lights = np.concatenate([dimmed.dir_lights, scene.dir_lights])
    r.ctx.upload_lights(more); ref.upload_lights(more)
    assert r.ctx.pending_ahead()[0] == 0
    r.frame(max_bounces=4); ref.render(cam, 5, 1, st)
    a, b = r.image(6), ref.resolve(6)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6)   # two shadow rays per vertex: unordered atomics
    r.close(); ref.close()


def test_wave_budget_changes_the_footprint_not_the_image(lib, oracle):
    """bpt_set_wave_budget: fewer samples in flight per wave (a smaller footprint for a pass that shares the GPU) — the accumulated image is
    the same bit for bit, because samples are folded into the sum in frame order whatever the wave size."""
    scene = scenes.small_test_scene()
    W, H = 64, 48
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    imgs = []
    for budget in (0, W * H * 3, 1):                       # default (64 samples per wave here), 3 samples per wave, 1 sample per wave
        gpu = capi.Context(lib, W, H); gpu.upload_scene(scene, capi.ACCEL_MERGED)
        gpu.set_wave_budget(budget)
        gpu.render(cam, 0, 7, st)
        imgs.append(gpu.resolve(7)); gpu.close()
    np.testing.assert_array_equal(imgs[0], imgs[1]); np.testing.assert_array_equal(imgs[0], imgs[2])
    ref = oracle.OracleContext(W, H); ref.upload_scene(scene, capi.ACCEL_MERGED); ref.render(cam, 0, 7, st)
    np.testing.assert_array_equal(imgs[0], ref.resolve(7)); ref.close()


def test_renderer_plugin_runs_pt_and_post_process(lib, oracle):
    """The plugin level: an IRenderer registered by name and selected like `renderer = "..."` in project.toml runs
    BasicRenderer's path-tracing pipeline — update_params per frame, PathTracingPass::render + PostProcessPass::render per camera
    (basic.cpp:31-48,157-166,228-231) — as two render-graph passes; the back buffer equals oracle render + oracle post-process."""
    scene = scenes.small_test_scene()                       # one directional light, procedural sky
    W, H = 64, 48
    gpu = capi.Context(lib, W, H); gpu.upload_scene(scene, capi.ACCEL_MERGED)
    img, passes = engine.run_renderer(gpu, scene, W, H, 3, max_bounces=4, bloom=True, bloom_threshold=0.4)
    assert passes == 2
    ref = oracle.OracleContext(W, H); ref.upload_scene(scene, capi.ACCEL_MERGED)
    ref.render(engine.camera_matrices(scene.camera, W, H), 0, 3, capi.Settings(max_bounces=4))
    np.testing.assert_array_equal(img, oracle.post_process_image(ref.resolve(3), capi.PostSettings(True, 0.4, 0.5)))
    with pytest.raises(KeyError):
        engine.run_renderer(gpu, scene, W, H, 1, renderer="BasicRenderer")        # only what was registered can be selected
    gpu.close(); ref.close()


# ---- rect-light textures (lights.hlsl:425-447,495-511; mip chain: shaders/core/mipmap.hlsl) --------------------------
@pytest.mark.parametrize("fmt", [capi.TEXTURE_RGBA8_UNORM, capi.TEXTURE_RGBA8_SRGB, capi.TEXTURE_RGBA32_FLOAT])
def test_light_texture_mip_chain_bit_exact(lib, oracle, fmt):
    """k_light_tex_level0 + k_mip_downsample == the oracle's chain, texel for texel (odd sizes, 1-texel-wide levels, every format)."""
    gpu, ref = capi.Context(lib, 16, 16), oracle.OracleContext(16, 16)
    texs = [scenes.light_texture(w, h, fmt, levels=16, seed=w) for (w, h) in ((37, 22), (64, 64), (9, 5), (1, 7), (130, 3))]
    gpu.upload_light_textures(texs); ref.upload_light_textures(texs)
    for k in range(len(texs)):
        np.testing.assert_array_equal(gpu.read_light_texture(k), ref.read_light_texture(k))
    gpu.upload_light_textures([]); ref.upload_light_textures([])            # clearing is allowed
    gpu.close(); ref.close()


def test_textured_rect_lights_parity(lib, oracle):
    """Rect lights with textures (three formats, linear and nearest mip filters, two-sided and one-sided), through the whole pass."""
    luts = scenes.load_ltc_luts(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ltc_luts.npz"))
    scene = scenes.add_mixed_lights(scenes.small_test_scene(), 2, 4, luts, keep_dir_lights=True, light_range=12.0)
    scene.rect_lights["two_sided"][::2] = 1
    texs = [scenes.light_texture(37, 22, capi.TEXTURE_RGBA8_SRGB, mip_linear=1), scenes.light_texture(16, 16, capi.TEXTURE_RGBA8_UNORM, mip_linear=0, seed=8),
            scenes.light_texture(9, 5, capi.TEXTURE_RGBA32_FLOAT, levels=3, mip_linear=1, linear=0, seed=9)]
    plain = scenes.add_mixed_lights(scenes.small_test_scene(), 2, 4, luts, keep_dir_lights=True, light_range=12.0)
    plain.rect_lights["two_sided"][::2] = 1
    scenes.texture_rect_lights(scene, texs)
    W, H = 96, 64
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    for mode in (capi.ACCEL_MERGED, capi.ACCEL_TWO_LEVEL):
        gpu, ref = make_pair(lib, oracle, scene, W, H, mode)
        gpu.render(cam, 0, 2, st); ref.render(cam, 0, 2, st)
        a, b = gpu.resolve(2), ref.resolve(2)
        assert (np.abs(a - b) / np.maximum(np.abs(b), 1e-3)).max() <= 1e-4          # several lights per vertex: unordered shadow-ray atomics
        gpu.close(); ref.close()
    # the textures matter: the same lights untextured give a different image
    gpu2 = capi.Context(lib, W, H); gpu2.upload_scene(plain, capi.ACCEL_MERGED)
    gpu2.render(cam, 0, 2, st)
    assert np.abs(gpu2.resolve(2) - a).max() > 0.02
    gpu2.close()


# ---- BASELINE configs[2]: mixed lights (64 point/spot + 16 LTC rect) ---------------------------------
def test_mixed_lights_config3(lib, oracle):
    import os
    luts = scenes.load_ltc_luts(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ltc_luts.npz"))
    scene = scenes.mixed_lights(luts)
    assert len(scene.point_lights) == 64 and len(scene.rect_lights) == 16
    W, H = 160, 90
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=3)                                   # reference default (basic.hpp:78)
    gpu.debug_capture(True); ref.debug_capture(True)
    gpu.render(cam, 0, 1, st); ref.render(cam, 0, 1, st)
    a, b = gpu.resolve(1), ref.resolve(1)
    # up to 64 shadow rays per vertex land on one pixel through unordered atomics → 1e-4 relative, not bitwise
    assert (np.abs(a - b) / np.maximum(np.abs(b), 1e-3)).max() <= 1e-4
    for bounce in (1, 2):                                               # NEE-heavy shadow queue: contents bit-exact as a set
        sa, sb = gpu.read_queue(bounce, 1), ref.read_queue(bounce, 1)
        np.testing.assert_array_equal(np.sort(sa["pixels"].astype(np.uint64) << 32 | sa["lights"]),
                                      np.sort(sb["pixels"].astype(np.uint64) << 32 | sb["lights"]))
    ca, cb = gpu.counters(), ref.counters()
    assert ca.shadow_rays == cb.shadow_rays and ca.extend_rays == cb.extend_rays
    assert ca.shadow_rays > 5 * ca.extend_rays                          # many lights per vertex
    gpu.close()
    big = capi.Context(lib, 1920, 1080)                                 # full size: runs, finite, bounded ray counts
    big.upload_scene(scene, capi.ACCEL_MERGED)
    big.render(engine.camera_matrices(scene.camera, 1920, 1080), 0, 2, st)
    img = big.resolve(2)
    c = big.counters()
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
    assert c.shadow_rays <= 64 * c.extend_rays and c.extend_rays_per_bounce[1] == 2 * 1920 * 1080


# ---- BASELINE configs[3]: large mesh x many instances, two-level BVH ---------------------------------
@pytest.mark.parametrize("mode", MODES)
def test_instanced_small_parity(lib, oracle, mode):
    scene = scenes.instanced(48, 24, 3)
    W, H = 128, 72
    gpu, ref = make_pair(lib, oracle, scene, W, H, mode)
    if mode == capi.ACCEL_TWO_LEVEL:
        assert_bvh_equal(gpu.read_bvh(0), ref.read_bvh(0))
        assert_bvh_equal(gpu.read_bvh(capi.BVH_TLAS), ref.read_bvh(capi.BVH_TLAS))
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=8, ray_length=1000.0)
    gpu.render(cam, 0, 2, st); ref.render(cam, 0, 2, st)
    np.testing.assert_array_equal(gpu.resolve(2), ref.resolve(2))
    ca, cb = gpu.counters(), ref.counters()
    assert list(ca.extend_rays_per_bounce) == list(cb.extend_rays_per_bounce)


def test_instanced_full_size_config4(lib, oracle):
    """2 097 152-triangle mesh x 512 instances: bit-exact BLAS/TLAS, bit-exact hits of 20 000 random rays
    against the oracle, and a 3840x2160 sample with sane counters."""
    scene = scenes.instanced()
    assert scene.blas["num_triangles"][0] == 2097152 and len(scene.instances) == 513
    gpu, ref = make_pair(lib, oracle, scene, 64, 36, capi.ACCEL_TWO_LEVEL)
    assert_bvh_equal(gpu.read_bvh(0), ref.read_bvh(0))
    assert_bvh_equal(gpu.read_bvh(capi.BVH_TLAS), ref.read_bvh(capi.BVH_TLAS))
    rays = random_rays(scene, 20000, 17)
    rays["tmax"] = 1000.0
    a, b = gpu.trace_rays(rays, 1), ref.trace_rays(rays, 1)
    for f in ("t", "u", "v", "instance", "primitive"):
        np.testing.assert_array_equal(a[f], b[f])
    assert (a["t"] > 0).mean() > 0.3
    cam = engine.camera_matrices(scene.camera, 64, 36)
    st = capi.Settings(max_bounces=8, ray_length=1000.0)
    gpu.render(cam, 0, 1, st); ref.render(cam, 0, 1, st)
    np.testing.assert_array_equal(gpu.resolve(1), ref.resolve(1))
    with pytest.raises(capi.BptError):
        gpu.build_accel(capi.ACCEL_MERGED)                              # 1.07 G merged triangles: refused, not attempted
    gpu.close()
    W, H = 3840, 2160
    big = capi.Context(lib, W, H)
    big.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
    big.render(engine.camera_matrices(scene.camera, W, H), 0, 1, st)
    img = big.resolve(1)
    c = big.counters()
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
    assert c.extend_rays_per_bounce[1] == W * H and c.extend_rays_per_bounce[2] > 0.2 * W * H


# ---- BASELINE configs[4]: DDGI-style probe tracing through the same extend / shade kernels -----------
@pytest.mark.parametrize("mode", MODES)
def test_probe_tracing_parity(lib, oracle, mode):
    scene = scenes.small_test_scene()
    gpu, ref = make_pair(lib, oracle, scene, 32, 32, mode)             # wavefront capacity 64 * 1024 paths → exercises chunking
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (8, 6, 8), 256, ray_length=100.0)  # 98 304 probe rays
    for bounces in (1, 3):
        a, b = gpu.trace_probes(vol, table, 9, bounces), ref.trace_probes(vol, table, 9, bounces)
        np.testing.assert_array_equal(a[:, 3], b[:, 3])                # hit distances: bit-exact
        np.testing.assert_array_equal(a[:, :3], b[:, :3])              # one light → identical add order


def test_probe_range_is_the_sharding_unit(lib, oracle):
    """SURVEY §8e for the DDGI update: the probes of a volume split into per-rank ranges (sharding.probe_range) trace to exactly the
    rows of the full call — keys are global probe indices — so an all-gather of the ranges reproduces the single-GPU update."""
    from bisemutum_engine_b200 import sharding
    scene = scenes.small_test_scene()
    gpu, ref = make_pair(lib, oracle, scene, 32, 32, capi.ACCEL_TWO_LEVEL)
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (5, 3, 7), 64, ray_length=100.0)   # 105 probes: uneven over 4 ranks
    full = gpu.trace_probes(vol, table, 3, 2)
    np.testing.assert_array_equal(full, ref.trace_probes(vol, table, 3, 2))
    parts = [gpu.trace_probes_range(vol, table, 3, 2, *sharding.probe_range(105, r, 4)) for r in range(4)]
    np.testing.assert_array_equal(np.concatenate(parts).view(np.uint32), full.view(np.uint32))
    np.testing.assert_array_equal(parts[2], ref.trace_probes_range(vol, table, 3, 2, *sharding.probe_range(105, 2, 4)))
    assert gpu.trace_probes_range(vol, table, 3, 2, 10, 0).shape == (0, 4)
    # device-resident results (what a sharded update keeps between trace, all-gather and blend): same bits, same atlases
    import torch
    dev = torch.empty((105 * 64, 4), dtype=torch.float32, device="cuda")
    gpu.trace_probes_range_into(vol, table, 3, 2, 0, 105, dev.data_ptr())
    np.testing.assert_array_equal(dev.cpu().numpy().view(np.uint32), full.view(np.uint32))
    irr_d, vis_d = gpu.blend_probes_from_device(vol, table, 3, dev.data_ptr())
    irr_h, vis_h = gpu.blend_probes(vol, table, 3, full)
    np.testing.assert_array_equal(irr_d, irr_h); np.testing.assert_array_equal(vis_d, vis_h)
    with pytest.raises(capi.BptError):
        gpu.trace_probes_range(vol, table, 3, 2, 100, 6)


def test_probe_tracing_full_size_config5(lib, oracle):
    """32 x 32 x 16 probes x 256 rays = 4 194 304 rays per bounce over the atrium; a 1/64 slice of the probes is
    compared with the oracle bit-exactly (probe rays are independent), the full volume is checked for sanity."""
    scene = scenes.atrium()
    gpu = capi.Context(lib, 1920, 1080)
    gpu.upload_scene(scene, capi.ACCEL_MERGED)
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (32, 32, 16), 256)
    out = gpu.trace_probes(vol, table, 0, 2)
    assert out.shape == (4194304, 4) and np.isfinite(out).all()
    assert 0.5 < (out[:, 3] > 0).mean() <= 1.0 and out[:, :3].mean() > 0
    c = gpu.counters()
    assert c.extend_rays_per_bounce[1] == 4194304 and 0 < c.extend_rays_per_bounce[2] < 4194304
    ref = oracle.OracleContext(8, 8)
    ref.upload_scene(scene, capi.ACCEL_MERGED)
    sub = scenes.probe_volume(scene, (32, 8, 1), 256)                   # the first 256 probes of the z = 0 slab ... same base/extent
    sub.extent[:] = [vol.extent[0], vol.extent[1] * 7 / 31, 0.0]
    b = ref.trace_probes(sub, table, 0, 2)
    a = gpu.trace_probes(sub, table, 0, 2)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("switch", ["russian_roulette", "pixel_jitter", "nee_none"])
def test_mode_switches(lib, oracle, switch):
    scene = scenes.small_test_scene()
    W, H = 96, 64
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=8, nee_mode=capi.NEE_NONE) if switch == "nee_none" else capi.Settings(max_bounces=8, **{switch: 1})
    gpu.render(cam, 4, 3, st); ref.render(cam, 4, 3, st)
    np.testing.assert_array_equal(gpu.resolve(3), ref.resolve(3))
    ca, cb = gpu.counters(), ref.counters()
    assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays
    with pytest.raises(capi.BptError):
        gpu.render(cam, 0, 1, capi.Settings(state_precision=1))             # a change of accumulation rule needs clear_accum
    with pytest.raises(capi.BptError):
        gpu.render(cam, 0, 1, capi.Settings(state_precision=7))             # unknown value: rejected, not ignored


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("scene_fn,W,H,bounces", [(scenes.small_test_scene, 96, 64, 6), (scenes.cornell_box, 100, 75, 5)])
def test_reference_fp16_state(lib, oracle, scene_fn, W, H, bounces, mode):
    """state_precision = reference_fp16 (SURVEY §8a storage note; rows a1/a8/a16/a17 literally): half ray directions and
    throughput, the packed G-buffer through its texture formats, the half additive blit per bounce and the running half
    lerp of pt_accumulate. One light => one shadow ray per vertex => the image is bit-identical to the oracle."""
    scene = scene_fn()
    gpu, ref = make_pair(lib, oracle, scene, W, H, mode)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=bounces, state_precision=capi.STATE_REFERENCE_FP16)
    gpu.render(cam, 0, 3, st); ref.render(cam, 0, 3, st)
    gpu.render(cam, 3, 2, st); ref.render(cam, 3, 2, st)                   # the sample count carries over between calls
    a, b = gpu.resolve(5), ref.resolve(5)
    np.testing.assert_array_equal(a[..., :3].view(np.uint32), b[..., :3].view(np.uint32))
    np.testing.assert_array_equal(a[..., :3], a[..., :3].astype(np.float16).astype(np.float32))   # an rgba16_sfloat image
    ca, cb = gpu.counters(), ref.counters()
    assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays
    gpu.clear_accum(); ref.clear_accum()                                    # back to the FP32 rule
    gpu.render(cam, 0, 2, capi.Settings(max_bounces=bounces)); ref.render(cam, 0, 2, capi.Settings(max_bounces=bounces))
    f32 = gpu.resolve(2)
    np.testing.assert_array_equal(f32, ref.resolve(2))
    assert (f32[..., :3] != a[..., :3]).any()


@pytest.mark.parametrize("mode", MODES)
def test_primary_outputs_and_rtao(lib, oracle, mode):
    """bpt_render_primary (OutputData.depth / .gbuffer) and bpt_trace_ao (RTAO through the connect kernel, cull-non-opaque
    any-hit rays) against the oracle: bit-exact, on the reference's own example scene (alpha-tested + translucent drawables)."""
    import os
    scene = scenes.scene_basic(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene_basic.npz"))
    W, H = 160, 96
    gpu, ref = make_pair(lib, oracle, scene, W, H, mode)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    depth, g = gpu.render_primary(cam, 2, st)
    rdepth, rg = ref.render_primary(cam, 2, st)
    np.testing.assert_array_equal(depth.view(np.uint32), rdepth.view(np.uint32))
    for f in capi.GBUFFER_TEXEL.names:
        np.testing.assert_array_equal(g[f].view(np.uint32), rg[f].view(np.uint32), err_msg=f)
    assert 0.3 < (depth > 0).mean() < 1.0
    for half, frame in ((False, 5), (True, 4), (True, 7)):
        a = gpu.trace_ao(cam, frame, depth, g["normal_roughness"], 0.5, 0.5, half)
        b = ref.trace_ao(cam, frame, rdepth, rg["normal_roughness"], 0.5, 0.5, half)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
        assert (a[..., 1] == 1).any() and (a[..., 0] < 1).any()
    ca, cb = gpu.counters(), ref.counters()
    assert ca.extend_rays == cb.extend_rays == W * H and ca.shadow_rays == cb.shadow_rays > 0
    # the render path is unaffected by the extra passes (they reuse its buffers)
    gpu.render(cam, 0, 2, st); ref.render(cam, 0, 2, st)
    np.testing.assert_array_equal(gpu.resolve(2), ref.resolve(2))


def test_ddgi_volume_lighting_and_feedback(lib, oracle):
    """bpt_set_ddgi_volume / bpt_ddgi_lighting (calc_ddgi_volume_lighting) and the previous-update feedback inside
    bpt_trace_probes: a two-update DDGI loop (trace -> blend -> bind -> trace with feedback -> blend) matches the oracle bit for bit."""
    scene = scenes.small_test_scene()
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (6, 4, 4), 128, ray_length=100.0)
    gpu, ref = make_pair(lib, oracle, scene, 16, 16, capi.ACCEL_MERGED)
    r0g, r0r = gpu.trace_probes(vol, table, 0, 1), ref.trace_probes(vol, table, 0, 1)
    np.testing.assert_array_equal(r0g, r0r)
    irr_g, vis_g = gpu.blend_probes(vol, table, 0, r0g); irr_r, vis_r = ref.blend_probes(vol, table, 0, r0r)
    np.testing.assert_array_equal(irr_g, irr_r); np.testing.assert_array_equal(vis_g, vis_r)
    gpu.set_ddgi_volume(vol, irr_g, vis_g); ref.set_ddgi_volume(vol, irr_r, vis_r)
    rng = np.random.default_rng(9)
    lo = np.float32(vol.base_position[:]); ext = np.float32(vol.extent[:])
    n = 20000
    pos = (lo + rng.uniform(-0.1, 1.1, (n, 3)) * ext).astype(np.float32)
    nrm = rng.normal(size=(n, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    view = rng.normal(size=(n, 3)); view = (view / np.linalg.norm(view, axis=1, keepdims=True)).astype(np.float32)
    a, b = gpu.ddgi_lighting(pos, nrm, view), ref.ddgi_lighting(pos, nrm, view)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    assert (a[:, 3] == 1).mean() > 0.3 and a[:, :3].max() > 0
    for bounces in (1, 3):
        r1g, r1r = gpu.trace_probes(vol, table, 1, bounces), ref.trace_probes(vol, table, 1, bounces)
        np.testing.assert_array_equal(r1g.view(np.uint32), r1r.view(np.uint32))
    assert (r1g[:, :3] >= 0).all()
    irr1_g, _ = gpu.blend_probes(vol, table, 1, r1g, irr_g, vis_g); irr1_r, _ = ref.blend_probes(vol, table, 1, r1r, irr_r, vis_r)
    np.testing.assert_array_equal(irr1_g, irr1_r)
    gpu.set_ddgi_volume(None); ref.set_ddgi_volume(None)
    np.testing.assert_array_equal(gpu.trace_probes(vol, table, 1, 1), ref.trace_probes(vol, table, 1, 1))
    with pytest.raises(capi.BptError):
        gpu.ddgi_lighting(pos[:4], nrm[:4], view[:4])                                        # nothing bound


def test_project_loader_upload(lib, oracle, tmp_path):
    """host/project.cpp end to end on the GPU: a project directory in the reference's on-disk formats (written by tests/_mini_project.py:
    v1 + v2 .biasset meshes, a texture, three material snippets incl. the alpha-tested cage, dir + point light) is loaded and uploaded by
    the C++ host library, and renders bit-identically to the oracle fed with the same arrays through the Python path."""
    import _mini_project
    _mini_project.write(str(tmp_path))
    p = engine.Project(str(tmp_path))
    W, H = p.info.target_width, p.info.target_height
    st = capi.Settings(max_bounces=p.info.max_bounces, ray_length=p.info.ray_length)
    cam = engine.camera_matrices(p.camera(), W, H)
    for mode in MODES:
        gpu = capi.Context(lib, W, H)
        p.upload(gpu, mode)
        ref = oracle.OracleContext(W, H); ref.upload_scene(p.scene_data(), mode)
        gpu.render(cam, 0, 2, st); ref.render(cam, 0, 2, st)
        a, b = gpu.resolve(2), ref.resolve(2)
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6)                 # two lights: shadow-ray terms add through unordered atomics
        ca, cb = gpu.counters(), ref.counters()
        assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays and a[..., :3].mean() > 0.02
        gpu.close()
    p.close()


def test_native_host_renders_a_project(lib, oracle, tmp_path):
    """The whole drop-in path in C++ only (host/render_project.cpp): project directory -> loader -> C ABI -> renderer selected by name ->
    frames -> post-process -> PFM. The image equals the oracle's render + post-process of the same project."""
    import subprocess
    import _mini_project
    _mini_project.write(str(tmp_path))
    exe = os.path.join(pkg.PACKAGE_DIR, "host", "render_project")
    assert os.path.exists(exe), "build it first: make -C bisemutum-engine_b200/host (there is no fallback)"
    out = str(tmp_path / "out.pfm")
    r = subprocess.run([exe, str(tmp_path), out, "3", "--merged", "--bloom", "0.4", "0.5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "CudaPathTracingRenderer: 3 frames" in r.stdout
    with open(out, "rb") as f:
        assert f.readline() == b"PF\n"
        w, h = (int(v) for v in f.readline().split())
        assert f.readline().strip() == b"-1.0"
        img = np.frombuffer(f.read(), "<f4").reshape(h, w, 3)[::-1]
    p = engine.Project(str(tmp_path))
    assert (w, h) == (p.info.target_width, p.info.target_height)
    ref = oracle.OracleContext(w, h); ref.upload_scene(p.scene_data(), capi.ACCEL_MERGED)
    ref.render(engine.camera_matrices(p.camera(), w, h), 0, 3, capi.Settings(max_bounces=p.info.max_bounces, ray_length=p.info.ray_length))
    want = oracle.post_process_image(ref.resolve(3), capi.PostSettings(True, 0.4, 0.5))[..., :3]
    np.testing.assert_allclose(img, want, rtol=2e-3, atol=1e-5)     # two lights (unordered atomics) through the half stores of the bloom chain
    bad = subprocess.run([exe, str(tmp_path), out, "1", "--renderer", "BasicRenderer"], capture_output=True, text=True)
    assert bad.returncode == 1 and "not registered" in bad.stderr
    p.close()


def test_library_reduce_single_rank(lib, oracle):
    """bpt_comm_unique_id / bpt_comm_init / bpt_reduce with a one-rank communicator: NCCL loads, the reduce is the identity, the
    fp16 running average refuses it. (The N >= 2 check is tests/check_reduce_multigpu.py under torchrun; the CPU suite covers the sharding
    arithmetic with gloo.)"""
    scene = scenes.small_test_scene()
    W, H = 64, 48
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    with pytest.raises(capi.BptError):
        gpu.reduce(0)                                                        # no communicator yet
    gpu.comm_init(lib.comm_unique_id(), 0, 1)
    gpu.render(cam, 0, 2, st); ref.render(cam, 0, 2, st)
    gpu.reduce(0); gpu.sync()
    np.testing.assert_array_equal(gpu.resolve(2), ref.resolve(2))
    gpu.clear_accum(); gpu.render(cam, 0, 1, capi.Settings(max_bounces=4, state_precision=capi.STATE_REFERENCE_FP16))
    with pytest.raises(capi.BptError):
        gpu.reduce(0)


def test_reference_fp16_host_pass(lib, oracle):
    """The frame-at-a-time host pass (render_ahead / accumulate_ahead) follows the same running lerp."""
    scene = scenes.small_test_scene()
    W, H = 64, 48
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=5, state_precision=capi.STATE_REFERENCE_FP16)
    n = gpu.render_ahead(cam, 10, 4, st)
    assert n >= 1
    for _ in range(n):
        gpu.accumulate_ahead(1)
    ref.render(cam, 10, n, st)
    np.testing.assert_array_equal(gpu.resolve(n)[..., :3].view(np.uint32), ref.resolve(n)[..., :3].view(np.uint32))


def test_rect_shadow_switch(lib, oracle):
    """rect_shadow = mrp_ray: one shadow ray per rect light towards its most representative point."""
    import os
    luts = scenes.load_ltc_luts(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ltc_luts.npz"))
    scene = scenes.add_mixed_lights(scenes.small_test_scene(), 4, 3, luts, keep_dir_lights=True, light_range=12.0)
    W, H = 80, 56
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_TWO_LEVEL)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4, rect_shadow=1)
    gpu.render(cam, 1, 2, st); ref.render(cam, 1, 2, st)
    a, b = gpu.resolve(2), ref.resolve(2)
    assert (np.abs(a - b) / np.maximum(np.abs(b), 1e-3)).max() <= 1e-4
    ca, cb = gpu.counters(), ref.counters()
    assert ca.shadow_rays == cb.shadow_rays and ca.extend_rays == cb.extend_rays
    base = oracle.OracleContext(W, H); base.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
    base.render(cam, 1, 2, capi.Settings(max_bounces=4))
    assert cb.shadow_rays > base.counters().shadow_rays                     # the rect lights now cast rays
    assert (b[..., :3] <= base.resolve(2)[..., :3] + 1e-5).all()            # shadowing never adds light


def test_update_tlas_and_resize(lib, oracle):
    """The reference rebuilds its TLAS every frame (render_graph.cpp:818-825): move instances, bpt_update_tlas,
    compare with the oracle doing the same; then bpt_resize drops the history and renders at the new extent."""
    scene = scenes.instanced(24, 12, 2)
    W, H = 64, 40
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_TWO_LEVEL)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4, ray_length=1000.0)
    rng = np.random.default_rng(5)
    moved = scene.instances.copy()
    for i in range(len(moved) - 1):
        moved[i]["transform"][:, 3] += rng.uniform(-1.0, 1.0, 3).astype(np.float32)
    for ctx in (gpu, ref):
        ctx.upload_instances(moved)
        ctx.update_tlas()
    assert_bvh_equal(gpu.read_bvh(capi.BVH_TLAS), ref.read_bvh(capi.BVH_TLAS))
    gpu.render(cam, 0, 1, st); ref.render(cam, 0, 1, st)
    a = gpu.resolve(1)
    np.testing.assert_array_equal(a, ref.resolve(1))
    # moving instances changed the picture
    still = oracle.OracleContext(W, H); still.upload_scene(scene, capi.ACCEL_TWO_LEVEL); still.render(cam, 0, 1, st)
    assert (a != still.resolve(1)).any()
    W2, H2 = 48, 48
    for ctx in (gpu, ref):
        ctx._call("resize", W2, H2); ctx.width, ctx.height = W2, H2
    cam2 = engine.camera_matrices(scene.camera, W2, H2)
    gpu.render(cam2, 3, 1, st); ref.render(cam2, 3, 1, st)
    np.testing.assert_array_equal(gpu.resolve(1), ref.resolve(1))           # history was dropped: exactly one sample
    with pytest.raises(capi.BptError):
        merged = capi.Context(lib, 8, 8); merged.upload_scene(scene, capi.ACCEL_MERGED); merged.update_tlas()


def test_probe_blending_parity(lib, oracle):
    """SURVEY §8f rank 1 on the device: trace → blend → blend with history, bit-exact against the oracle."""
    scene = scenes.small_test_scene()
    gpu, ref = make_pair(lib, oracle, scene, 32, 32, capi.ACCEL_MERGED)
    table = scenes.ddgi_sample_randoms()
    vol = scenes.probe_volume(scene, (8, 8, 8), 64, ray_length=100.0)       # the reference's 8^3 probes x 64 rays
    ra, rb = gpu.trace_probes(vol, table, 0, 1), ref.trace_probes(vol, table, 0, 1)
    np.testing.assert_array_equal(ra, rb)
    ia, va = gpu.blend_probes(vol, table, 0, ra)
    ib, vb = ref.blend_probes(vol, table, 0, rb)
    np.testing.assert_array_equal(ia, ib); np.testing.assert_array_equal(va, vb)
    assert ia.shape == (64, 512, 4) and va.shape == (128, 1024, 2)          # ddgi.hpp:70-78 atlas sizes
    r1 = gpu.trace_probes(vol, table, 1, 1)
    ia2, va2 = gpu.blend_probes(vol, table, 1, r1, ia, va)
    ib2, vb2 = ref.blend_probes(vol, table, 1, r1, ib, vb)
    np.testing.assert_array_equal(ia2, ib2); np.testing.assert_array_equal(va2, vb2)


def test_converged_image_relmse(lib, oracle):
    """BASELINE.json gate: the converged image reaches relative MSE <= 1e-3 against a 4096-spp oracle render
    (configs[0] Cornell box, reduced to 64x64 so the double-precision oracle finishes in seconds)."""
    scene = scenes.cornell_box(tess=8)
    W, H, N = 64, 64, 4096
    gpu, ref = make_pair(lib, oracle, scene, W, H, capi.ACCEL_MERGED)
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=5)
    truth = ref.render_converged(cam, 0, N, st)[..., :3].astype(np.float64)       # 4096 spp, float64 accumulation
    gpu.render(cam, 0, N, st)
    img = gpu.resolve(N)[..., :3].astype(np.float64)

    def relmse(a, b):
        return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-4)))
    assert relmse(img, truth) <= 1e-3
    assert relmse(img, truth) <= 1e-9          # in fact: same samples, FP32 vs FP64 summation only
    # an independent set of 4096 samples (other frame indices) differs only by Monte-Carlo noise
    gpu.clear_accum(); gpu.render(cam, 100000, N, st)
    other = gpu.resolve(N)[..., :3].astype(np.float64)
    assert relmse(other, truth) < 0.5 and abs(other.mean() / truth.mean() - 1) < 0.05
