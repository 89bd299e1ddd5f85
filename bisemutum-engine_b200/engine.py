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
        h.bpt_host_pass_set_color_target.argtypes = [C.c_void_p, C.c_void_p]
        h.bpt_host_pass_reset_history.argtypes = [C.c_void_p]
        h.bpt_host_pass_read_primary.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
        h.bpt_host_pass_frame.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        h.bpt_host_pass_frame.restype = C.c_int
        h.bpt_host_pass_post_process.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        h.bpt_host_pass_post_process.restype = C.c_int
        h.bpt_host_renderer_run.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(HostCameraDesc), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.c_void_p, C.c_void_p, C.c_float, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_void_p, C.POINTER(C.c_uint32)]
        h.bpt_host_renderer_run.restype = C.c_int
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


class ProjectInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_drawables", "num_blas", "num_materials", "num_textures", "num_dir_lights", "num_point_lights",
                                          "num_rect_lights", "target_width", "target_height")] + \
               [("camera", HostCameraDesc), ("ray_length", C.c_float), ("max_bounces", C.c_uint32), ("accumulate", C.c_uint32),
                ("ambient_occlusion", capi.AoSettings)]


class Project:
    """A bisemutum project directory (project.toml, asset_metadata.toml, scene.toml, materials/*.toml, *.biasset) loaded by the C++
    host library (host/project.cpp) — the headless stand-in for the reference's asset manager + ECS deserialisation."""
    _ARRAYS = {"positions": (0, np.float32), "normals": (1, np.float32), "tangents": (2, np.float32), "texcoords": (3, np.float32),
               "indices": (4, np.uint32), "blas": (5, capi.BLAS_DESC), "drawables": (6, capi.DRAWABLE_SBT), "instances": (7, capi.INSTANCE_DESC),
               "materials": (8, capi.MATERIAL), "dir_lights": (9, capi.DIR_LIGHT), "point_lights": (10, capi.POINT_LIGHT), "rect_lights": (11, capi.RECT_LIGHT)}

    def __init__(self, directory: str, _loader: str = "bpt_host_project_load"):
        h = host_library()
        h.bpt_host_project_load.argtypes, h.bpt_host_project_load.restype = [C.c_char_p, C.c_char_p, C.c_uint64], C.c_void_p
        h.bpt_host_project_import_gltf.argtypes, h.bpt_host_project_import_gltf.restype = [C.c_char_p, C.c_char_p, C.c_uint64], C.c_void_p
        h.bpt_host_project_free.argtypes = [C.c_void_p]
        h.bpt_host_project_get_info.argtypes = [C.c_void_p, C.POINTER(ProjectInfo)]
        h.bpt_host_project_array.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        h.bpt_host_project_array.restype = C.c_void_p
        h.bpt_host_project_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        h.bpt_host_project_error.argtypes, h.bpt_host_project_error.restype = [C.c_void_p], C.c_char_p
        err = C.create_string_buffer(512)
        self._h = getattr(h, _loader)(directory.encode(), err, 512)
        if not self._h:
            raise RuntimeError(f"project {directory}: {err.value.decode()}")
        self.info = ProjectInfo()
        h.bpt_host_project_get_info(self._h, C.byref(self.info))

    @classmethod
    def from_gltf(cls, path: str) -> "Project":
        """A .gltf / .glb file through the C++ importer (host/gltf.cpp = the reference's `menu_action_import_model_gltf`,
        import_model.cpp:27-430): meshes, MikkTSpace tangents, metallic-roughness materials, textures and the node hierarchy.
        Camera and lights are not part of the import (the reference ignores glTF cameras / lights too)."""
        return cls(path, _loader="bpt_host_project_import_gltf")

    def array(self, name: str) -> np.ndarray:
        which, dt = self._ARRAYS[name]
        n = C.c_uint64()
        p = host_library().bpt_host_project_array(self._h, which, C.byref(n), None, None, None)
        return np.frombuffer((C.c_uint8 * n.value).from_address(p), dtype=dt).copy() if n.value else np.zeros(0, dt)

    def texture(self, k: int):
        """(texels (H, W, 4) uint8 for the 8-bit formats, rhi format id)."""
        n, w, hh, f = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        p = host_library().bpt_host_project_array(self._h, 16 + k, C.byref(n), C.byref(w), C.byref(hh), C.byref(f))
        raw = np.frombuffer((C.c_uint8 * n.value).from_address(p), dtype=np.uint8).copy()
        return (raw.reshape(hh.value, w.value, 4) if f.value in (37, 43) else raw.view(np.float32).reshape(hh.value, w.value, 4)), f.value

    def camera(self) -> dict:
        c = self.info.camera
        return dict(position=tuple(c.position), front_dir=tuple(c.front_dir), up_dir=tuple(c.up_dir), yfov=c.yfov, near_z=c.near_z, far_z=c.far_z,
                    orthographic=bool(c.orthographic))

    def scene_data(self):
        """The loaded project as a scenes.SceneData (for contexts that upload through the Python path, e.g. the CPU oracle in tests)."""
        from . import scenes
        texs = []
        for k in range(self.info.num_textures):
            texels, fmt = self.texture(k)
            smp = (C.c_uint32 * 4)()
            host_library().bpt_host_project_texture_sampler(C.c_void_p(self._h), k, smp)
            texs.append({"texels": np.ascontiguousarray(texels), "width": texels.shape[1], "height": texels.shape[0],
                         "format": {43: capi.TEXTURE_RGBA8_SRGB, 109: capi.TEXTURE_RGBA32_FLOAT}.get(fmt, capi.TEXTURE_RGBA8_UNORM),
                         "address_u": capi.ADDRESS_REPEAT if smp[2] == 0 else capi.ADDRESS_CLAMP,
                         "address_v": capi.ADDRESS_REPEAT if smp[3] == 0 else capi.ADDRESS_CLAMP, "linear": 1 if smp[0] == 1 else 0})   # as upload_project maps them
        nd = self.info.num_drawables
        va = np.full(nd, capi.VA_POSITION | capi.VA_NORMAL | capi.VA_TANGENT | capi.VA_TEXCOORD, np.uint32)
        return scenes.SceneData(name="project", positions=self.array("positions"), normals=self.array("normals"), tangents=self.array("tangents"),
                                texcoords=self.array("texcoords"), indices=self.array("indices"), drawables=self.array("drawables"), drawable_va=va,
                                blas=self.array("blas"), instances=self.array("instances"), materials=self.array("materials"), textures=texs,
                                dir_lights=self.array("dir_lights"), point_lights=self.array("point_lights"), rect_lights=self.array("rect_lights"),
                                camera=self.camera())

    def upload(self, ctx: capi.Context, accel_mode: int = capi.ACCEL_TWO_LEVEL):
        st = host_library().bpt_host_project_upload(self._h, ctx._h, accel_mode)
        if st != 0:
            raise capi.BptError(st, "bpt_host_project_upload", host_library().bpt_host_project_error(self._h).decode())

    def close(self):
        if self._h:
            host_library().bpt_host_project_free(self._h)
            self._h = None


def run_renderer(ctx: "capi.Context", scene, width: int, height: int, frames: int, renderer: str = "CudaPathTracingRenderer", ray_length: float = 100.0,
                 max_bounces: int = 3, bloom: bool = False, bloom_threshold: float = 1.5, bloom_threshold_softness: float = 0.5):
    """The plugin path: GraphicsManager::register_renderer<CudaPathTracingRenderer>() + set_renderer(name), then `frames` engine frames of
    prepare_renderer_per_frame_data / render_camera / RenderGraph::execute on a context whose geometry, materials and accel are uploaded.
    Returns (back buffer (H, W, 4) float32, render-graph passes per frame); raises KeyError for an unregistered renderer name."""
    h = host_library()
    d = camera_desc(scene.camera, width, height)
    out = np.zeros((height, width, 4), np.float32)
    passes = C.c_uint32(0)
    dirs = np.ascontiguousarray(scene.dir_lights)
    faces = None if scene.sky_faces is None else np.ascontiguousarray(scene.sky_faces, np.float32)
    xf = np.ascontiguousarray(scene.sky_transform, np.float32); col = np.ascontiguousarray(scene.sky_color, np.float32)
    st = h.bpt_host_renderer_run(ctx._h, renderer.encode(), C.byref(d), frames, dirs.ctypes.data_as(C.c_void_p), len(dirs),
                                 faces.ctypes.data_as(C.c_void_p) if faces is not None else None, 0 if faces is None else faces.shape[1],
                                 xf.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), ray_length, max_bounces, 1 if bloom else 0,
                                 bloom_threshold, bloom_threshold_softness, out.ctypes.data_as(C.c_void_p), C.byref(passes))
    if st == -1:
        raise KeyError(f"renderer {renderer!r} is not registered")
    if st != 0:
        raise capi.BptError(st, "CudaPathTracingRenderer::render_camera", ctx.L.fn("last_error")(ctx._h).decode())
    return out, passes.value


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

    def reset_history(self):
        """The next frame() starts a new accumulation: the image is cleared and samples traced ahead are dropped (PathTracingPass::reset_history)."""
        host_library().bpt_host_pass_reset_history(self._pass)

    def set_color_target(self, device_ptr: int):
        """Binds the device memory behind OutputData.color (rgba16_sfloat): every frame() then writes its accumulated colour there (0: unbound)."""
        host_library().bpt_host_pass_set_color_target(self._pass, C.c_void_p(device_ptr) if device_ptr else None)

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

    def post_process(self, bloom: bool = False, bloom_threshold: float = 1.5, bloom_threshold_softness: float = 0.5) -> np.ndarray:
        """PostProcessPass::render on the camera's accumulated colour (basic.cpp:228-231): the back-buffer image, (H, W, 4) float32."""
        out = np.zeros((self.height, self.width, 4), np.float32)
        st = host_library().bpt_host_pass_post_process(self._pass, 1 if bloom else 0, bloom_threshold, bloom_threshold_softness, out.ctypes.data_as(C.c_void_p))
        if st != 0:
            raise capi.BptError(st, "PostProcessPass::render", self.lib.fn("last_error")(self.ctx._h).decode())
        return out

    def close(self):
        if self._pass:
            host_library().bpt_host_pass_destroy(self._pass)
            self._pass = None
        self.ctx.close()
