"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. It reuses the POD layouts of the product's C ABI (bisemutum_engine_b200.capi) because
the oracle mirrors that ABI one-to-one under the `obpt_` prefix.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import bisemutum_engine_b200 as pkg
from bisemutum_engine_b200 import capi

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "extend_rays", "extend_nodes", "extend_tris", "extend_instances",
        "shadow_rays", "shadow_nodes", "shadow_tris", "shadow_instances",
        "shaded_vertices", "miss_vertices", "samples",
        "extend_wide_nodes", "extend_leaf_boxes", "shadow_wide_nodes", "shadow_leaf_boxes", "extend_wide_rays", "shadow_wide_rays")]


class CameraDesc(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("front_dir", C.c_float * 3), ("up_dir", C.c_float * 3),
                ("yfov", C.c_float), ("near_z", C.c_float), ("far_z", C.c_float), ("aspect", C.c_float),
                ("orthographic", C.c_uint32)]


def build(force: bool = False) -> str:
    """Compiles the oracle with its Makefile if the .so is missing or stale."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp", ".h"))]
    srcs.append(os.path.join(pkg.REPO_ROOT, "include", "bpt", "bpt.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def library() -> capi.Library:
    global _lib
    if _lib is None:
        build()
        _lib = capi.Library(LIB_PATH, "obpt_", {
            "set_threads": [C.c_void_p, C.c_uint32],
            "get_stats": [C.c_void_p, C.POINTER(Stats)],
            "set_tile_sample": [C.c_void_p, C.c_uint32, C.c_uint32],
            "render_converged": [C.c_void_p, C.POINTER(capi.Camera), C.c_uint32, C.c_uint32, C.POINTER(capi.Settings), C.c_void_p],
        })
        L = _lib.lib
        L.obpt_get_threads.argtypes, L.obpt_get_threads.restype = [C.c_void_p], C.c_uint32
        L.obpt_camera_matrices.argtypes, L.obpt_camera_matrices.restype = [C.POINTER(CameraDesc), C.c_void_p, C.c_void_p, C.POINTER(capi.Camera)], None
        L.obpt_frustum_planes.argtypes, L.obpt_frustum_planes.restype = [C.POINTER(CameraDesc), C.c_void_p], None
        L.obpt_cull_aabbs.argtypes, L.obpt_cull_aabbs.restype = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p], None
        L.obpt_transform_aabb.argtypes, L.obpt_transform_aabb.restype = [C.c_void_p, C.c_void_p, C.c_void_p], None
        L.obpt_rng_tea.argtypes, L.obpt_rng_tea.restype = [C.c_uint32, C.c_uint32], C.c_uint32
        L.obpt_rng_lcg.argtypes, L.obpt_rng_lcg.restype = [C.POINTER(C.c_uint32)], C.c_uint32
        L.obpt_sincos_2pi.argtypes, L.obpt_sincos_2pi.restype = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)], None
        L.obpt_atan2.argtypes, L.obpt_atan2.restype = [C.c_float, C.c_float], C.c_float
        L.obpt_acos.argtypes, L.obpt_acos.restype = [C.c_float], C.c_float
        L.obpt_ggx_vndf_sample.argtypes, L.obpt_ggx_vndf_sample.restype = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p], None
        L.obpt_surface_eval_lit.argtypes = [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_void_p]
        L.obpt_surface_eval_lit.restype = None
        L.obpt_store_half.argtypes, L.obpt_store_half.restype = [C.c_float], C.c_float
        L.obpt_gbuffer_roundtrip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.obpt_gbuffer_roundtrip.restype = None
        L.obpt_set_wide_from_bounce.argtypes, L.obpt_set_wide_from_bounce.restype = [C.c_void_p, C.c_uint32], C.c_int
        L.obpt_wide_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.obpt_wide_stats.restype = C.c_int
        L.obpt_morton63.argtypes, L.obpt_morton63.restype = [C.c_void_p, C.c_void_p, C.c_void_p], C.c_uint64
        L.obpt_post_process_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(capi.PostSettings), C.c_void_p]
        L.obpt_post_process_image.restype = C.c_int
    return _lib


class OracleContext(capi.Context):
    def __init__(self, width, height, threads: int = 0):
        super().__init__(library(), width, height)
        if threads:
            self._call("set_threads", threads)

    @property
    def threads(self) -> int:
        return self.L.lib.obpt_get_threads(self._h)

    def set_tile_sample(self, stride: int, offset: int = 0):
        self._call("set_tile_sample", stride, offset)

    def stats(self) -> Stats:
        s = Stats()
        self._call("get_stats", C.byref(s))
        return s

    def set_wide_from_bounce(self, bounce: int):
        """Bounces >= bounce walk the 4-wide quantised tree (what the CUDA kernels do with 2); 0 = binary everywhere."""
        st = library().lib.obpt_set_wide_from_bounce(self._h, bounce)
        if st != 0:
            raise capi.BptError(st, "obpt_set_wide_from_bounce", "")

    def wide_stats(self, rays, width=4, quantised=True, order=0):
        """Closest hits through a `width`-ary collapse of the merged BVH + work counters (oracle_wide.cpp).
        Returns (t, prim, dict(rays, nodes, tris, boxes, leaf_boxes))."""
        t = np.zeros(len(rays), np.float32); prim = np.zeros(len(rays), np.uint32); cnt = np.zeros(5, np.uint64)
        st = library().lib.obpt_wide_stats(self._h, width, 1 if quantised else 0, order, rays.ctypes.data_as(C.c_void_p), len(rays),
                                          t.ctypes.data_as(C.c_void_p), prim.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
        if st != 0:
            raise capi.BptError(st, "obpt_wide_stats", "")
        return t, prim, dict(zip(("rays", "nodes", "tris", "boxes", "leaf_boxes"), (int(x) for x in cnt)))

    def render_converged(self, camera, frame_first, num_samples, settings) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.float32)
        self._call("render_converged", C.byref(camera), frame_first, num_samples, C.byref(settings), out.ctypes.data_as(C.c_void_p))
        return out


def post_process_image(image: np.ndarray, settings: capi.PostSettings) -> np.ndarray:
    """PostProcessPass::render (post_process.cpp:92-273) on an H x W x 4 float32 image: oracle_post.cpp."""
    image = np.ascontiguousarray(image, np.float32)
    h, w = image.shape[:2]
    out = np.zeros_like(image)
    st = library().lib.obpt_post_process_image(image.ctypes.data_as(C.c_void_p), w, h, C.byref(settings), out.ctypes.data_as(C.c_void_p))
    if st != 0:
        raise capi.BptError(st, "obpt_post_process_image", "")
    return out


def camera_desc(cam: dict, aspect: float) -> CameraDesc:
    d = CameraDesc()
    d.position[:] = cam["position"]; d.front_dir[:] = cam["front_dir"]; d.up_dir[:] = cam["up_dir"]
    d.yfov, d.near_z, d.far_z, d.aspect = cam["yfov"], cam["near_z"], cam["far_z"], aspect
    d.orthographic = 1 if cam.get("orthographic") else 0
    return d


def camera_matrices(cam: dict, width: int, height: int) -> capi.Camera:
    """Oracle restatement of Camera::update_shader_params (camera.cpp:73-118)."""
    out = capi.Camera()
    d = camera_desc(cam, float(np.float32(width) / np.float32(height)))
    library().lib.obpt_camera_matrices(C.byref(d), None, None, C.byref(out))
    return out


def frustum_planes(cam: dict, width: int, height: int) -> np.ndarray:
    planes = np.zeros((6, 4), np.float32)
    d = camera_desc(cam, float(np.float32(width) / np.float32(height)))
    library().lib.obpt_frustum_planes(C.byref(d), planes.ctypes.data_as(C.c_void_p))
    return planes


def cull_aabbs(planes: np.ndarray, aabbs: np.ndarray) -> np.ndarray:
    aabbs = np.ascontiguousarray(aabbs, np.float32)
    vis = np.zeros(len(aabbs), np.uint8)
    library().lib.obpt_cull_aabbs(np.ascontiguousarray(planes, np.float32).ctypes.data_as(C.c_void_p),
                                  aabbs.ctypes.data_as(C.c_void_p), len(aabbs), vis.ctypes.data_as(C.c_void_p))
    return vis
