"""Python view of the C++ host mirror (host/engine.hpp → libbpt_host.so) on top of libbpt.so.

`Renderer` is what a user of the reference's `BasicRenderer` in path-tracing mode gets: create it
for a target extent, hand it a scene, then call `frame()` once per engine frame — each frame is one
sample per pixel accumulated into the history, exactly like `PathTracingPass::render`
(bisemutum/src/renderer/pass/path_tracing.cpp:224-488).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import LIBBPT_HOST_PATH, capi, load_library


class HostCameraDesc(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("front_dir", C.c_float * 3), ("up_dir", C.c_float * 3),
                ("yfov", C.c_float), ("near_z", C.c_float), ("far_z", C.c_float),
                ("width", C.c_uint32), ("height", C.c_uint32), ("orthographic", C.c_uint32)]


_host = None


def host_library():
    """libbpt_host.so (C++ host mirror). Loading it also loads libbpt.so (its dependency)."""
    global _host
    if _host is None:
        if not os.path.exists(LIBBPT_HOST_PATH):
            raise FileNotFoundError(f"{LIBBPT_HOST_PATH} is missing — run __graft_entry__.build()")
        load_library()     # make sure libbpt.so resolves from the in-tree path first
        h = C.CDLL(LIBBPT_HOST_PATH)
        h.bpt_host_camera_matrices.argtypes = [C.POINTER(HostCameraDesc), C.POINTER(capi.Camera), C.c_void_p, C.c_void_p]
        h.bpt_host_camera_matrices.restype = None
        h.bpt_host_frustum_planes.argtypes = [C.POINTER(HostCameraDesc), C.c_void_p]
        h.bpt_host_cull.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        h.bpt_host_transform_aabb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        h.bpt_host_pack_point_light.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        h.bpt_host_pack_rect_light.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        h.bpt_host_pack_dir_light.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        h.bpt_host_pass_create.argtypes, h.bpt_host_pass_create.restype = [C.c_void_p, C.POINTER(HostCameraDesc)], C.c_void_p
        h.bpt_host_pass_destroy.argtypes = [C.c_void_p]
        h.bpt_host_pass_set_camera.argtypes = [C.c_void_p, C.POINTER(HostCameraDesc)]
        h.bpt_host_pass_set_frame.argtypes = [C.c_void_p, C.c_uint64]
        h.bpt_host_pass_set_prefetch.argtypes = [C.c_void_p, C.c_uint32]
        h.bpt_host_pass_read_primary.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
        h.bpt_host_pass_frame.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        h.bpt_host_pass_frame.restype = C.c_int
        _host = h
    return _host


def camera_desc(cam: dict, width: int, height: int) -> HostCameraDesc:
    d = HostCameraDesc()
    d.position[:] = cam["position"]
    d.front_dir[:] = cam["front_dir"]
    d.up_dir[:] = cam["up_dir"]
    d.yfov, d.near_z, d.far_z = cam.get("yfov", 30.0), cam.get("near_z", 0.01), cam.get("far_z", 10000.0)
    d.width, d.height = width, height
    d.orthographic = 1 if cam.get("orthographic") else 0
    return d


def camera_matrices(cam: dict, width: int, height: int) -> capi.Camera:
    """gfx::Camera::update_shader_params (host mirror of camera.cpp:73-118)."""
    out = capi.Camera()
    d = camera_desc(cam, width, height)
    host_library().bpt_host_camera_matrices(C.byref(d), C.byref(out), None, None)
    return out


def frustum_planes(cam: dict, width: int, height: int) -> np.ndarray:
    planes = np.zeros((6, 4), np.float32)
    d = camera_desc(cam, width, height)
    host_library().bpt_host_frustum_planes(C.byref(d), planes.ctypes.data_as(C.c_void_p))
    return planes


def cull_aabbs(planes: np.ndarray, aabbs: np.ndarray) -> np.ndarray:
    """Visibility list of RenderGraph::add_rendered_object_list's frustum test (render_graph.cpp:391-461)."""
    aabbs = np.ascontiguousarray(aabbs, np.float32)
    planes = np.ascontiguousarray(planes, np.float32)
    vis = np.zeros(len(aabbs), np.uint8)
    host_library().bpt_host_cull(planes.ctypes.data_as(C.c_void_p), aabbs.ctypes.data_as(C.c_void_p), len(aabbs), vis.ctypes.data_as(C.c_void_p))
    return vis


def drawable_world_aabbs(scene) -> np.ndarray:
    """World AABB per drawable = Transform::transform_bounding_box(mesh bbox) (drawable.cpp:5-7)."""
    out = np.zeros((len(scene.instances), 6), np.float32)
    pos = scene.positions
    h = host_library()
    for i, inst in enumerate(scene.instances):
        b = scene.blas[int(inst["blas"])]
        idx = scene.indices[int(b["index_offset"]): int(b["index_offset"]) + 3 * int(b["num_triangles"])]
        p = pos[int(b["position_offset"]):].reshape(-1, 3)[idx]
        local = np.concatenate([p.min(0), p.max(0)]).astype(np.float32)
        m = np.ascontiguousarray(inst["transform"], np.float32)
        h.bpt_host_transform_aabb(m.ctypes.data_as(C.c_void_p), local.ctypes.data_as(C.c_void_p), out[i].ctypes.data_as(C.c_void_p))
    return out


class Renderer:
    """BasicRenderer in path-tracing mode, reduced to what the pass needs."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.lib = load_library()
        self.ctx = capi.Context(self.lib, width, height, device)
        self.width, self.height = width, height
        self._pass = None
        self.settings = capi.Settings()
        self.scene = None

    def set_scene(self, scene, accel_mode: int = capi.ACCEL_MERGED):
        self.scene = scene
        self.ctx.upload_scene(scene, accel_mode)
        d = camera_desc(scene.camera, self.width, self.height)
        h = host_library()
        if self._pass:
            h.bpt_host_pass_destroy(self._pass)
        self._pass = h.bpt_host_pass_create(self.ctx._h, C.byref(d))

    def set_camera(self, cam: dict):
        d = camera_desc(cam, self.width, self.height)
        host_library().bpt_host_pass_set_camera(self._pass, C.byref(d))

    def set_prefetch(self, frames: int):
        """Samples traced ahead per wave while the history stays valid (PathTracingPass::set_prefetch_frames)."""
        host_library().bpt_host_pass_set_prefetch(self._pass, frames)

    def set_frame(self, frame: int):
        host_library().bpt_host_pass_set_frame(self._pass, frame)

    def frame(self, ray_length: float = 100.0, max_bounces: int = 3, accumulate: bool = True) -> int:
        """One engine frame (= 1 spp). Returns the number of frames accumulated in the history."""
        n = C.c_uint64()
        st = host_library().bpt_host_pass_frame(self._pass, ray_length, max_bounces, 1 if accumulate else 0, C.byref(n))
        if st != 0:
            raise capi.BptError(st, "PathTracingPass::render", self.lib.fn("last_error")(self.ctx._h).decode())
        return n.value

    def primary_outputs(self, ray_length: float = 100.0, max_bounces: int = 3):
        """OutputData.depth / .gbuffer of the current frame: (H, W) float32 depth, (H, W) capi.GBUFFER_TEXEL."""
        depth = np.zeros((self.height, self.width), np.float32)
        g = np.zeros((self.height, self.width), capi.GBUFFER_TEXEL)
        st = host_library().bpt_host_pass_read_primary(self._pass, ray_length, max_bounces, depth.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
        if st != 0:
            raise capi.BptError(st, "PathTracingPass::read_primary_outputs", self.lib.fn("last_error")(self.ctx._h).decode())
        return depth, g

    def image(self, accumulated_frames: int) -> np.ndarray:
        return self.ctx.resolve(accumulated_frames)

    def close(self):
        if self._pass:
            host_library().bpt_host_pass_destroy(self._pass)
            self._pass = None
        self.ctx.close()
