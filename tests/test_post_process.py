"""The step after the path: PostProcessPass::render = bloom + output pass
(bisemutum/src/renderer/pass/post_process.cpp:92-273, shaders/renderer/post_process/*.hlsl).

CPU: the oracle (oracle/oracle_post.cpp, pass by pass) against an independent float64 numpy restatement and against the
properties the shaders imply; the CUDA source's per-pixel functions (csrc/bpt_post.cuh, host build) == oracle bit for bit.
GPU: the fused kernels of csrc/post.cu through the C ABI == oracle bit for bit, at odd sizes and at 1920x1080.
"""
import numpy as np
import pytest

import _hostcheck as HC
import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi

OFFSETS = np.array([-3.23076923, -1.38461538, 0.0, 1.38461538, 3.23076923])
WEIGHTS = np.array([0.07027027, 0.31621622, 0.22702703, 0.31621622, 0.07027027])


def hdr_image(h, w, seed, bright=0.02):
    """Mostly dim image with a few very bright texels and blobs (what bloom is for)."""
    rng = np.random.default_rng(seed)
    img = rng.random((h, w, 4), dtype=np.float32) * np.float32(0.8)
    hot = rng.random((h, w)) < bright
    img[hot, :3] += rng.random((int(hot.sum()), 3), dtype=np.float32) * np.float32(40.0)
    img[..., 3] = 1.0
    return img


# ---- independent restatement: float64, no half stores, coordinates computed exactly -----------------------------------
def _bilinear(img, xs, ys):
    """img: (h, w, 3) float64; xs, ys: texel-space sample positions (already minus 0.5), clamp to edge."""
    h, w = img.shape[:2]
    x0 = np.floor(xs); y0 = np.floor(ys)
    fx = (xs - x0)[..., None]; fy = (ys - y0)[..., None]
    xi0 = np.clip(x0.astype(int), 0, w - 1); xi1 = np.clip(x0.astype(int) + 1, 0, w - 1)
    yi0 = np.clip(y0.astype(int), 0, h - 1); yi1 = np.clip(y0.astype(int) + 1, 0, h - 1)
    top = img[yi0, xi0] * (1 - fx) + img[yi0, xi1] * fx
    bot = img[yi1, xi0] * (1 - fx) + img[yi1, xi1] * fx
    return top * (1 - fy) + bot * fy


def _filter(src, dw, dh, vertical):
    sh, sw = src.shape[:2]
    y, x = np.mgrid[0:dh, 0:dw]
    u = (x + 0.5) / dw; v = (y + 0.5) / dh
    out = np.zeros((dh, dw, 3))
    for o, wgt in zip(OFFSETS, WEIGHTS):
        uu, vv = (u, v + o / dh) if vertical else (u + o / dw, v)
        out += wgt * _bilinear(src, uu * sw - 0.5, vv * sh - 0.5)
    return out


def _combine(c1, c2):
    h, w = c1.shape[:2]
    y, x = np.mgrid[0:h, 0:w]
    return c1 + _bilinear(c2, (x + 0.5) / w * c2.shape[1] - 0.5, (y + 0.5) / h * c2.shape[0] - 0.5)


def numpy_bloom(img, threshold, softness):
    c = img[..., :3].astype(np.float64)
    h, w = c.shape[:2]
    soft_t = softness * (threshold * 0.9 + 0.1)
    bx = threshold; by = threshold * soft_t; bw = 0.25 / (by + 0.00001); by -= threshold
    lum = c @ np.array([0.212671, 0.715160, 0.072169])
    soft = np.clip(lum + by, 0.0, bx) ** 2 * bw
    pre = c * (np.maximum(soft, lum - bx) / np.maximum(lum, 0.0001))[..., None]
    t, src = [], pre
    for i in range(3):
        dw, dh = max(w >> (i + 1), 1), max(h >> (i + 1), 1)
        hp = _filter(src, dw, dh, False); t.append(hp)
        src = _filter(hp, dw, dh, True); t.append(src)
    c2 = _combine(t[3], t[5]); c1 = _combine(t[1], c2)
    return _combine(c, c1)


@pytest.mark.parametrize("size", [(48, 64), (45, 77), (135, 240)])
def test_oracle_matches_float64_restatement(oracle, size):
    h, w = size
    img = hdr_image(h, w, 3)
    got = oracle.post_process_image(img, capi.PostSettings(True, 1.5, 0.5))
    want = numpy_bloom(img, 1.5, 0.5)
    assert np.all(got[..., 3] == 1.0)
    # differences: half stores of the 10 intermediate targets (2^-11 relative each) and FP32 coordinates
    np.testing.assert_allclose(got[..., :3], want, rtol=6e-3, atol=2e-3)
    assert np.abs(got[..., :3] - img[..., :3]).max() > 0.5                      # the bloom really added light around the hot texels


def test_oracle_properties(oracle):
    img = hdr_image(40, 56, 5)
    off = oracle.post_process_image(img, capi.PostSettings(False))
    np.testing.assert_array_equal(off[..., :3], img[..., :3]); assert np.all(off[..., 3] == 1.0)      # post_process.hlsl: (xyz, 1)
    # nothing above the (soft) threshold -> bloom adds exactly 0; the output is the colour through one rgba16_sfloat store
    dim = (np.minimum(img, np.float32(1.0)) * np.float32(0.3)).astype(np.float32)
    got = oracle.post_process_image(dim, capi.PostSettings(True, 2.0, 0.0))
    np.testing.assert_array_equal(got[..., :3], dim[..., :3].astype(np.float16).astype(np.float32))
    # constant bright image: the filter weights sum to 1 and the three levels add up -> colour * (1 + 3 * pre-weight)
    const = np.zeros((64, 64, 4), np.float32); const[..., :3] = (4.0, 2.0, 1.0); const[..., 3] = 1
    got = oracle.post_process_image(const, capi.PostSettings(True, 1.5, 0.5))
    lum = 4.0 * 0.212671 + 2.0 * 0.715160 + 1.0 * 0.072169
    wgt = (lum - 1.5) / lum
    np.testing.assert_allclose(got[32, 32, :3], np.array([4.0, 2.0, 1.0]) * (1 + 3 * wgt), rtol=3e-3)
    # a NaN / Inf texel is dropped by the filter passes (bloom_filter.hlsl:21,32) and only survives in its own pixel
    bad = hdr_image(40, 56, 6); bad[10, 10, 0] = np.inf; bad[20, 30, 1] = np.nan
    got = oracle.post_process_image(bad, capi.PostSettings(True, 1.5, 0.5))
    mask = np.ones((40, 56), bool); mask[10, 10] = False; mask[20, 30] = False
    assert np.isfinite(got[mask]).all()


CASES = [((37, 23), 11, capi.PostSettings(True, 1.5, 0.5)), ((64, 48), 12, capi.PostSettings(True, 0.7, 1.0)),
         ((75, 130), 13, capi.PostSettings(True, 3.0, 0.0)), ((9, 5), 14, capi.PostSettings(True, 1.0, 0.25)),
         ((33, 64), 15, capi.PostSettings(False))]


def _case_input(size, seed, with_specials=True):
    h, w = size
    img = hdr_image(h, w, seed, bright=0.05)
    if with_specials and h > 8 and w > 8:
        img[3, 4, 1] = np.inf; img[5, 2, 2] = np.nan; img[7, 7, :3] = -2.0; img[1, 1, :3] = 70000.0     # inf / NaN / negative / beyond half range
    spp = 7
    return (img * np.float32(spp)).astype(np.float32), spp            # sums, as the accumulation buffer holds them


@pytest.mark.parametrize("size,seed,st", CASES)
def test_cuda_source_bit_exact_on_host(oracle, size, seed, st):
    """csrc/bpt_post.cuh (host build, driven like post.cu drives it) == oracle_post.cpp, bit for bit, NaN for NaN."""
    sums, spp = _case_input(size, seed)
    resolved = sums * (np.float32(1.0) / np.float32(spp)); resolved[..., 3] = 1.0
    want = oracle.post_process_image(resolved, st)
    got = HC.post_process(sums, spp, st)
    np.testing.assert_array_equal(np.nan_to_num(got, nan=-777.0), np.nan_to_num(want, nan=-777.0))


@pytest.mark.gpu
@pytest.mark.parametrize("size,seed,st", CASES + [((1080, 1920), 16, capi.PostSettings(True, 1.5, 0.5))])
def test_gpu_post_process_bit_exact(oracle, size, seed, st):
    """bpt_post_process (csrc/post.cu: fused level kernels with the horizontal pass in shared memory) == oracle, bit for bit."""
    h, w = size
    sums, spp = _case_input(size, seed)
    resolved = sums * (np.float32(1.0) / np.float32(spp)); resolved[..., 3] = 1.0
    want = oracle.post_process_image(resolved, st)
    ctx = capi.Context(pkg.load_library(), w, h)
    ctx.upload_accum(sums)
    before = ctx.counters().kernel_launches
    got = ctx.post_process(st, spp)
    assert ctx.counters().kernel_launches - before == (6 if st.bloom else 1)          # 11 reference passes -> 6 launches
    np.testing.assert_array_equal(np.nan_to_num(got, nan=-777.0), np.nan_to_num(want, nan=-777.0))
    ctx.close()


@pytest.mark.gpu
def test_gpu_post_process_after_render(oracle):
    """The pass runs on what the path tracer accumulated: render -> post-process on both sides."""
    from bisemutum_engine_b200 import engine, scenes
    W, H = 96, 64
    scene = scenes.small_test_scene()
    cam = engine.camera_matrices(scene.camera, W, H)
    st = capi.Settings(max_bounces=4)
    gpu = capi.Context(pkg.load_library(), W, H); ref = oracle.OracleContext(W, H)
    for c in (gpu, ref):
        c.upload_scene(scene, capi.ACCEL_MERGED); c.render(cam, 0, 3, st)
    ps = capi.PostSettings(True, 0.4, 0.5)
    want = oracle.post_process_image(ref.resolve(3), ps)
    got = gpu.post_process(ps, 3)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(capi.BptError):
        gpu.post_process(capi.PostSettings(True, 1.5, 1.5), 3)                        # softness outside [0, 1]
    gpu.close(); ref.close()
