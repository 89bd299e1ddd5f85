"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must reproduce them too (GPU)."""
import importlib.util
import os

import numpy as np
import pytest

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden_v1.npz"))


def check(golden, got, multi_light):
    for k, v in got.items():
        if multi_light and ".image." in k:
            # several shadow rays per vertex add to a pixel through unordered atomics: 1e-4 relative (BASELINE.json)
            np.testing.assert_allclose(v, golden[k], rtol=1e-4, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_array_equal(v, golden[k], err_msg=k)


@pytest.mark.parametrize("case", ["small", "cornell", "mixed", "scene_basic"])
def test_oracle_reproduces_golden(oracle, golden, case):
    for name, scene, bounces in make_golden.cases():
        if name == case:
            check(golden, make_golden.compute(name, scene, bounces, oracle.OracleContext), False)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["small", "cornell", "mixed", "scene_basic"])
def test_cuda_reproduces_golden(golden, case):
    lib = pkg.load_library()
    for name, scene, bounces in make_golden.cases():
        if name == case:
            check(golden, make_golden.compute(name, scene, bounces, lambda w, h: capi.Context(lib, w, h)), case == "mixed")


# ---- golden_v2.npz: post-process, sky IBL, ray-traced reflections -------------------------------------------------------------
@pytest.fixture(scope="module")
def golden2():
    return np.load(os.path.join(GOLDEN_DIR, "golden_v2.npz"))


def test_oracle_reproduces_golden_v2(oracle, golden2):
    got = make_golden.compute_v2(oracle.OracleContext, lambda ctx, sums, spp, st: oracle.post_process_image(sums * (np.float32(1.0) / np.float32(spp)), st))
    assert set(got) == set(golden2.files)
    for k, v in got.items():
        np.testing.assert_array_equal(v, golden2[k], err_msg=k)


@pytest.mark.gpu
def test_cuda_reproduces_golden_v2(golden2):
    lib = pkg.load_library()

    def post(ctx, sums, spp, st):
        ctx.upload_accum(sums)                       # (the render's own sums, handed back: post-process reads the accumulation buffer)
        return ctx.post_process(st, spp)
    got = make_golden.compute_v2(lambda w, h: capi.Context(lib, w, h), post)
    for k, v in got.items():
        if k.endswith(".color"):                     # several lights -> unordered shadow-ray atomics: 1e-4 (BASELINE.json)
            np.testing.assert_allclose(v, golden2[k], rtol=1e-4, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_array_equal(v, golden2[k], err_msg=k)
