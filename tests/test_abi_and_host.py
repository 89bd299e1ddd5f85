"""C-ABI surface and host-side logic (no GPU): every symbol include/bpt/bpt.h declares is exported by
libbpt.so, the library refuses to work without a device (no CPU fallback), and the C++ host mirror
(camera matrices, frustum planes, culling list, light packing) agrees bit-exactly with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi, engine, scenes


def test_libbpt_exports_every_declared_symbol():
    header = open(os.path.join(pkg.REPO_ROOT, "include", "bpt", "bpt.h")).read()
    declared = sorted(set(re.findall(r"BPT_API\s+[\w\s\*]+?\b(bpt_\w+)\s*\(", header)))
    assert len(declared) >= 30
    lib = C.CDLL(pkg.LIBBPT_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.BPT_EXPORTS) == declared          # the Python binding covers the whole header
    lib.bpt_version.restype = C.c_char_p
    assert b"sm_100a" in lib.bpt_version()


def test_libbpt_host_exports_every_declared_symbol():
    """include/bpt/bpt_host.h is the C view of the host mirror; c_exports.cpp includes it, so the compiler has already checked the signatures."""
    header = open(os.path.join(pkg.REPO_ROOT, "include", "bpt", "bpt_host.h")).read()
    declared = sorted(set(re.findall(r"BPT_HOST_API\s+[\w\s\*]+?\b(bpt_host_\w+)\s*\(", header)))
    assert len(declared) >= 25
    lib = engine.host_library()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    defined = set(re.findall(r"^HOST_API\s+[\w\s\*]+?\b(bpt_host_\w+)\s*\(", open(os.path.join(pkg.PACKAGE_DIR, "host", "c_exports.cpp")).read(), re.M))
    assert defined == set(declared)                      # nothing exported that the header does not declare
    assert '#include "../../include/bpt/bpt_host.h"' in open(os.path.join(pkg.PACKAGE_DIR, "host", "c_exports.cpp")).read()
    assert C.sizeof(engine.ProjectInfo) == 9 * 4 + C.sizeof(engine.HostCameraDesc) + 12 + C.sizeof(capi.AoSettings)


def test_no_cpu_fallback():
    """Without a CUDA device bpt_create must fail loudly (BPT_ERR_NO_DEVICE), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BptError) as e:
        capi.Context(pkg.load_library(), 8, 8)
    assert e.value.status in (5, 2)
    src = "".join(open(os.path.join(pkg.PACKAGE_DIR, "csrc", f)).read() for f in os.listdir(os.path.join(pkg.PACKAGE_DIR, "csrc")) if f.endswith((".cu", ".cuh")))
    import re
    # the product never includes, links or calls the oracle; comments may CITE an oracle file as the definition a kernel reproduces
    assert "oracle" not in re.sub(r"oracle/oracle_\w+\.cpp", "", src) and "obpt_" not in src
    assert not [l for l in src.splitlines() if l.lstrip().startswith("#include") and "oracle" in l]
    for f in ("capi.py", "engine.py", "scenes.py", "__init__.py"):
        assert "oracle_py" not in open(os.path.join(pkg.PACKAGE_DIR, f)).read()


def test_struct_layouts_match_header():
    sizes = {"bpt_drawable_sbt_data": 36, "bpt_blas_desc": 16, "bpt_instance_desc": 64, "bpt_material": 64, "bpt_dir_light_data": 64,
             "bpt_point_light_data": 64, "bpt_rect_light_data": 112, "bpt_bvh_node": 64, "bpt_ray": 32, "bpt_hit": 20}
    src = '#include "bpt/bpt.h"\n#include <stdio.h>\nint main(){' + "".join(f'printf("{k} %zu\\n", sizeof({k}));' for k in sizes) + "return 0;}"
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(pkg.REPO_ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    assert {k: int(v) for k, v in got.items()} == sizes
    assert C.sizeof(capi.Camera) == 192 and C.sizeof(capi.Settings) == 32 and C.sizeof(capi.Counters) == 32 + 256


CAMS = [
    dict(position=(0.0, 1.0, 4.6), front_dir=(0.0, 0.0, -1.0), up_dir=(0.0, 1.0, 0.0), yfov=30.0, near_z=0.001, far_z=1e5),
    dict(position=(-16.5, 3.2, 0.6), front_dir=(1.0, 0.12, -0.03), up_dir=(0.0, 1.0, 0.0), yfov=30.0, near_z=0.001, far_z=1e5),
    dict(position=(3.0, 7.0, -2.0), front_dir=(-0.4, -0.8, 0.3), up_dir=(0.0, 1.0, 0.0), yfov=55.0, near_z=0.1, far_z=500.0),
    dict(position=(1.0, 2.0, 3.0), front_dir=(0.2, -0.1, -1.0), up_dir=(0.0, 1.0, 0.0), yfov=40.0, near_z=0.1, far_z=100.0, orthographic=True),
]


@pytest.mark.parametrize("cam", CAMS)
@pytest.mark.parametrize("size", [(1920, 1080), (512, 512), (333, 127)])
def test_camera_matrices_host_vs_oracle(oracle, cam, size):
    W, H = size
    a, b = engine.camera_matrices(cam, W, H), oracle.camera_matrices(cam, W, H)
    for f in ("matrix_inv_view", "matrix_inv_proj", "matrix_proj_view"):
        np.testing.assert_array_equal(np.array(getattr(a, f)), np.array(getattr(b, f)))
    iv = np.array(a.matrix_inv_view, np.float64).reshape(4, 4).T
    np.testing.assert_allclose(iv[:3, 3], cam["position"], rtol=1e-5, atol=1e-5)      # camera.hlsl:7-9
    assert np.isfinite(np.array(a.matrix_inv_proj)).all()
    # inverse really is the inverse: proj_view * inv_view * inv_proj ≈ I
    pv = np.array(a.matrix_proj_view, np.float64).reshape(4, 4).T
    ip = np.array(a.matrix_inv_proj, np.float64).reshape(4, 4).T
    np.testing.assert_allclose(pv @ iv @ ip, np.eye(4), atol=2e-3)


@pytest.mark.parametrize("cam", CAMS)
def test_frustum_planes_and_culling_list(oracle, cam):
    W, H = 1920, 1080
    pa, pb = engine.frustum_planes(cam, W, H), oracle.frustum_planes(cam, W, H)
    np.testing.assert_array_equal(pa, pb)
    front = np.asarray(cam["front_dir"], np.float64); front /= np.linalg.norm(front)
    inside = np.asarray(cam["position"], np.float64) + front * 5.0
    assert (pa[:, :3].astype(np.float64) @ inside + pa[:, 3] > 0).all()              # normals point inwards
    rng = np.random.default_rng(4)
    c = rng.uniform(-30, 30, (500, 3)); e = rng.uniform(0.05, 6, (500, 3))
    aabbs = np.concatenate([c - e, c + e], 1).astype(np.float32)
    va, vb = engine.cull_aabbs(pa, aabbs), oracle.cull_aabbs(pb, aabbs)
    np.testing.assert_array_equal(va, vb)                                            # visibility list: bit-exact
    assert 0 < va.sum() < len(va)
    # conservative: a box whose centre is inside all planes is never culled
    cin = (pa[:, :3].astype(np.float64) @ c.T + pa[:, 3:4] > 0).all(0)
    assert va[cin].all()


def test_drawable_culling_on_the_atrium(oracle):
    scene = scenes.atrium()
    boxes = engine.drawable_world_aabbs(scene)
    planes = engine.frustum_planes(scene.camera, 1920, 1080)
    vis = engine.cull_aabbs(planes, boxes)
    np.testing.assert_array_equal(vis, oracle.cull_aabbs(oracle.frustum_planes(scene.camera, 1920, 1080), boxes))
    assert 5 < vis.sum() < len(vis)      # the camera at one end sees part of the hall


def test_light_packing_matches_reference_rules():
    h = engine.host_library()
    out = np.zeros(1, capi.POINT_LIGHT); n = C.c_int()
    col = np.array([1.0, 0.5, 0.25], np.float32); tr = np.array([1, 2, 3], np.float32)
    rot = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32).reshape(9)      # +Y → -X
    h.bpt_host_pack_point_light(col.ctypes.data_as(C.c_void_p), 2.0, 30.0, 1, 30.0, 60.0, tr.ctypes.data_as(C.c_void_p),
                                rot.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(n))
    assert n.value == 1
    np.testing.assert_array_equal(out["emission"][0], col * 2)
    np.testing.assert_array_equal(out["direction"][0], [-1, 0, 0])
    assert out["range_sqr_inv"][0] == np.float32(1) / np.float32(900)
    assert abs(out["cos_inner"][0] - np.cos(np.radians(30))) < 1e-6 and abs(out["cos_outer"][0] - 0.5) < 1e-6
    h.bpt_host_pack_point_light(col.ctypes.data_as(C.c_void_p), 0.0, 30.0, 0, 30.0, 60.0, tr.ctypes.data_as(C.c_void_p),
                                rot.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(n))
    assert n.value == 0                                                             # zero emission is skipped (lights.cpp:130-132)
    r = np.zeros(1, capi.RECT_LIGHT)
    ident = np.eye(3, dtype=np.float32).reshape(9)
    h.bpt_host_pack_rect_light(col.ctypes.data_as(C.c_void_p), 1.0, 2.0, 1.0, 0, tr.ctypes.data_as(C.c_void_p),
                               ident.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), C.byref(n))
    np.testing.assert_array_equal(r["position0"][0], [2, 2.5, 3]); np.testing.assert_array_equal(r["position2"][0], [0, 1.5, 3])
    np.testing.assert_array_equal(r["normal"][0], [0, 0, 1]); assert r["inv_width_sqr"][0] == 0.25 and r["texture_index"][0] == -1
