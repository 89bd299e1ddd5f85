"""ctypes / numpy view of include/bpt/bpt.h.

The same POD layouts are used by the CUDA library (`libbpt.so`, prefix ``bpt_``) and, in the
tests only, by the CPU oracle (`liboracle.so`, prefix ``obpt_``) — that is what makes the parity
tests read "same calls, same inputs, two implementations".

Nothing here computes: it marshals host buffers across the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

# ---------------------------------------------------------------------------------------------
# POD layouts (numpy dtypes, byte-identical to the C structs; sizes asserted below)
# ---------------------------------------------------------------------------------------------
f32, u32, i32, u64 = np.float32, np.uint32, np.int32, np.uint64

DRAWABLE_SBT = np.dtype([(n, u32) for n in (
    "drawable_index", "position_offset", "normal_offset", "tangent_offset", "color_offset",
    "texcoord_offset", "texcoord2_offset", "index_offset", "material_offset")])
BLAS_DESC = np.dtype([("position_offset", u32), ("index_offset", u32), ("num_triangles", u32), ("reserved", u32)])
INSTANCE_DESC = np.dtype([("transform", f32, (3, 4)), ("instance_id_and_mask", u32),
                          ("sbt_offset_and_flags", u32), ("blas", u64)])
MATERIAL = np.dtype([("base_color", f32, 4), ("emission", f32, 3), ("roughness", f32), ("metallic", f32),
                     ("normal_map_scale", f32), ("occlusion_strength", f32), ("flags", u32),
                     ("base_color_tex", i32), ("metallic_roughness_tex", i32), ("normal_map_tex", i32),
                     ("occlusion_tex", i32)])
DIR_LIGHT = np.dtype([("emission", f32, 3), ("sm_index", i32), ("direction", f32, 3), ("shadow_strength", f32),
                      ("cascade_shadow_radius_sqr", f32, 4), ("shadow_depth_bias", f32),
                      ("shadow_normal_bias", f32), ("_pad1", f32, 2)])
POINT_LIGHT = np.dtype([("emission", f32, 3), ("range_sqr_inv", f32), ("position", f32, 3), ("cos_inner", f32),
                        ("direction", f32, 3), ("cos_outer", f32), ("sm_index", i32), ("shadow_strength", f32),
                        ("shadow_depth_bias", f32), ("shadow_normal_bias", f32)])
RECT_LIGHT = np.dtype([("emission", f32, 3), ("texture_index", i32), ("center_position", f32, 3), ("two_sided", u32),
                       ("position0", f32, 3), ("inv_width_sqr", f32), ("position1", f32, 3), ("inv_height_sqr", f32),
                       ("position2", f32, 3), ("inv_texel_size", f32), ("position3", f32, 3), ("_pad1", f32),
                       ("normal", f32, 3), ("_pad2", f32)])
BVH_NODE = np.dtype([(n, f32) for n in (
    "c0_lo_x", "c0_hi_x", "c0_lo_y", "c0_hi_y", "c1_lo_x", "c1_hi_x", "c1_lo_y", "c1_hi_y",
    "c0_lo_z", "c0_hi_z", "c1_lo_z", "c1_hi_z")] + [("child0", i32), ("child1", i32), ("parent", i32), ("reserved", u32)])
RAY = np.dtype([("origin", f32, 3), ("tmin", f32), ("direction", f32, 3), ("tmax", f32)])
HIT = np.dtype([("t", f32), ("u", f32), ("v", f32), ("instance", u32), ("primitive", u32)])

assert DRAWABLE_SBT.itemsize == 36 and BLAS_DESC.itemsize == 16 and INSTANCE_DESC.itemsize == 64
assert MATERIAL.itemsize == 64 and DIR_LIGHT.itemsize == 64 and POINT_LIGHT.itemsize == 64
assert RECT_LIGHT.itemsize == 112 and BVH_NODE.itemsize == 64 and RAY.itemsize == 32 and HIT.itemsize == 20

# enums
VA_POSITION, VA_NORMAL, VA_TANGENT, VA_COLOR, VA_TEXCOORD, VA_TEXCOORD2 = 1, 2, 4, 8, 16, 32
INSTANCE_FORCE_OPAQUE, INSTANCE_FORCE_NON_OPAQUE = 4, 8
MATERIAL_KIND_GLTF_PBR, MATERIAL_KIND_ASSIMP_DIFFUSE, MATERIAL_KIND_DEFAULT = 0, 1, 2
MATERIAL_KIND_CONSTANT_COLOR, MATERIAL_KIND_CHECKERBOARD, MATERIAL_KIND_TEXTURED, MATERIAL_KIND_TRANSPARENT, MATERIAL_KIND_CAGE = 3, 4, 5, 6, 7
MATERIAL_KIND_VERTEX_COLOR = 8
SURFACE_MODEL_UNLIT, SURFACE_MODEL_LIT = 0, 1
BLEND_OPAQUE, BLEND_ALPHA_TEST, BLEND_TRANSLUCENT = 0, 1, 2
MATERIAL_FLAG_TWO_SIDED = 1
TEXTURE_RGBA8_UNORM, TEXTURE_RGBA32_FLOAT, TEXTURE_RGBA8_SRGB = 0, 1, 2
ADDRESS_REPEAT, ADDRESS_CLAMP = 0, 1
ACCEL_TWO_LEVEL, ACCEL_MERGED = 0, 1
NEE_SHADOW_RAY, NEE_NONE = 0, 1
STATE_FP32, STATE_REFERENCE_FP16 = 0, 1
BVH_TLAS = 0xFFFFFFFF

STATUS_NAMES = {0: "BPT_OK", 1: "BPT_ERR_INVALID", 2: "BPT_ERR_CUDA", 3: "BPT_ERR_OOM", 4: "BPT_ERR_STATE",
                5: "BPT_ERR_NO_DEVICE", 6: "BPT_ERR_UNSUPPORTED"}


def material_flags(kind=MATERIAL_KIND_GLTF_PBR, blend=BLEND_OPAQUE, model=SURFACE_MODEL_LIT, two_sided=False) -> int:
    return (1 if two_sided else 0) | (kind << 8) | (blend << 16) | (model << 24)


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("max_lights_per_vertex", C.c_uint32)]


class GeometryStreams(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("num_position_floats", C.c_uint64),
                ("normals", C.c_void_p), ("num_normal_floats", C.c_uint64),
                ("tangents", C.c_void_p), ("num_tangent_floats", C.c_uint64),
                ("colors", C.c_void_p), ("num_color_floats", C.c_uint64),
                ("texcoords", C.c_void_p), ("num_texcoord_floats", C.c_uint64),
                ("texcoords2", C.c_void_p), ("num_texcoord2_floats", C.c_uint64),
                ("indices", C.c_void_p), ("num_indices", C.c_uint64)]


class TextureDesc(C.Structure):
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32),
                ("address_mode_u", C.c_uint32), ("address_mode_v", C.c_uint32), ("filter_linear", C.c_uint32)]


class ReblurSettings(C.Structure):
    """bpt_reblur_settings, defaults = the constants of ReblurPass::render (reblur.cpp:318-319,383)."""
    _fields_ = [("virtual_history", C.c_uint32), ("blur_radius", C.c_float), ("anti_flickering_strength", C.c_float), ("_pad", C.c_uint32)]

    def __init__(self, virtual_history=1, blur_radius=0.9, anti_flickering_strength=3.5):
        super().__init__(virtual_history, blur_radius, anti_flickering_strength, 0)


class ReblurInputs(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("noised", C.c_void_p), ("hit_positions", C.c_void_p), ("depth", C.c_void_p),
                ("normal_roughness", C.c_void_p), ("velocity", C.c_void_p), ("history_validation", C.c_void_p)]


class LightTextureDesc(C.Structure):
    """bpt_light_texture_desc: a rect-light texture (RectLightComponent::texture) and how its sampler filters it."""
    _fields_ = [("texels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("levels", C.c_uint32),
                ("address_mode_u", C.c_uint32), ("address_mode_v", C.c_uint32), ("filter_linear", C.c_uint32), ("mip_linear", C.c_uint32)]


class LtcLuts(C.Structure):
    _fields_ = [("matrix_lut0", C.c_void_p), ("matrix_lut1", C.c_void_p), ("matrix_lut2", C.c_void_p), ("norm_lut", C.c_void_p)]


class Camera(C.Structure):
    _fields_ = [("matrix_inv_view", C.c_float * 16), ("matrix_inv_proj", C.c_float * 16), ("matrix_proj_view", C.c_float * 16)]


class Settings(C.Structure):
    _fields_ = [("ray_length", C.c_float), ("max_bounces", C.c_uint32), ("accumulate", C.c_uint32), ("nee_mode", C.c_uint32),
                ("rect_shadow", C.c_uint32), ("russian_roulette", C.c_uint32), ("pixel_jitter", C.c_uint32),
                ("state_precision", C.c_uint32)]

    def __init__(self, ray_length=100.0, max_bounces=3, accumulate=1, nee_mode=NEE_SHADOW_RAY, **kw):
        super().__init__(ray_length=ray_length, max_bounces=max_bounces, accumulate=accumulate, nee_mode=nee_mode, **kw)


class Counters(C.Structure):
    _fields_ = [("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("samples", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("extend_rays_per_bounce", C.c_uint64 * 16), ("shadow_rays_per_bounce", C.c_uint64 * 16)]


class KernelTimes(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("raygen_ms", "extend_ms", "shade_ms", "connect_ms", "other_ms")] + \
               [(n, C.c_uint64) for n in ("raygen_launches", "extend_launches", "shade_launches", "connect_launches", "other_launches")]


class ProbeVolume(C.Structure):
    _fields_ = [("base_position", C.c_float * 3), ("_pad0", C.c_float), ("frame_x", C.c_float * 3), ("_pad1", C.c_float),
                ("frame_y", C.c_float * 3), ("_pad2", C.c_float), ("frame_z", C.c_float * 3), ("_pad3", C.c_float),
                ("extent", C.c_float * 3), ("ray_length", C.c_float), ("probe_counts", C.c_uint32 * 3), ("rays_per_probe", C.c_uint32)]


class ProbeBlend(C.Structure):
    _fields_ = [("irradiance_size", C.c_uint32), ("visibility_size", C.c_uint32), ("alpha", C.c_float), ("history_valid", C.c_uint32)]


class AoSettings(C.Structure):
    """bpt_ao_settings = BasicRenderer::AmbientOcclusionSettings (renderer/basic.hpp:40-50)."""
    _fields_ = [("range", C.c_float), ("strength", C.c_float), ("half_resolution", C.c_uint32)]


class ReflectionSettings(C.Structure):
    """bpt_reflection_settings = BasicRenderer::ReflectionSettings (renderer/basic.hpp:52-65), mode = raytraced."""
    _fields_ = [("range", C.c_float), ("strength", C.c_float), ("max_roughness", C.c_float), ("fade_roughness", C.c_float), ("half_resolution", C.c_uint32),
                ("ibl", C.c_uint32)]

    def __init__(self, range_=16.0, strength=1.0, max_roughness=0.3, fade_roughness=0.1, half_resolution=True, ibl=False):
        super().__init__(range_, strength, max_roughness, fade_roughness, 1 if half_resolution else 0, 1 if ibl else 0)


class SkyIblDesc(C.Structure):
    """bpt_sky_ibl_desc: texture sizes of SkyboxContext (src/renderer/context/skybox.cpp:11-29) + Skybox::diffuse/specular_strength."""
    _fields_ = [("diffuse_size", C.c_uint32), ("specular_size", C.c_uint32), ("specular_levels", C.c_uint32), ("brdf_lut_size", C.c_uint32),
                ("diffuse_strength", C.c_float), ("specular_strength", C.c_float)]

    def __init__(self, diffuse_size=256, specular_size=256, specular_levels=5, brdf_lut_size=128, diffuse_strength=1.0, specular_strength=1.0):
        super().__init__(diffuse_size, specular_size, specular_levels, brdf_lut_size, diffuse_strength, specular_strength)

    def specular_texels(self) -> int:
        return sum(6 * (self.specular_size >> l) ** 2 for l in range(self.specular_levels))


class PostSettings(C.Structure):
    """bpt_post_settings = the bloom fields of PostProcessVolume (include/bisemutum/renderer/post_process_volume.hpp:13-15)."""
    _fields_ = [("bloom", C.c_uint32), ("bloom_threshold", C.c_float), ("bloom_threshold_softness", C.c_float), ("_pad", C.c_uint32)]

    def __init__(self, bloom=False, bloom_threshold=1.5, bloom_threshold_softness=0.5):
        super().__init__(1 if bloom else 0, bloom_threshold, bloom_threshold_softness, 0)


GBUFFER_TEXEL = np.dtype([("base_color", np.float32, 4), ("normal_roughness", np.float32, 4), ("fresnel", np.float32, 4), ("material_0", np.float32, 4)])


class BptError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str):
        super().__init__(f"{where}: {STATUS_NAMES.get(status, status)} — {detail}")
        self.status = status


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


# Entry points every implementation of the boundary exports: name → (argtypes). restype is int
# (bpt_status) unless listed in _SPECIAL.
_VP, _U32, _U64, _PU32, _PU64 = C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
COMMON_API = {
    "create": [C.POINTER(Config), C.POINTER(_VP)],
    "destroy": [_VP],
    "resize": [_VP, _U32, _U32],
    "scene_upload_geometry": [_VP, C.POINTER(GeometryStreams), _VP, _VP, _U32, _VP, _U32],
    "scene_upload_instances": [_VP, _VP, _U32],
    "scene_upload_materials": [_VP, _VP, _U32, C.POINTER(TextureDesc), _U32],
    "scene_upload_lights": [_VP, _VP, _U32, _VP, _U32, _VP, _U32, C.POINTER(LtcLuts)],
    "scene_upload_sky": [_VP, _VP, _U32, _VP, _VP],
    "scene_upload_light_textures": [_VP, C.POINTER(LightTextureDesc), _U32],
    "denoise_reblur": [_VP, C.POINTER(Camera), _U64, C.POINTER(ReblurSettings), C.POINTER(ReblurInputs), _VP],
    "reblur_reset": [_VP],
    "debug_read_reblur": [_VP, _U32, _VP, _U64],
    "debug_read_light_texture": [_VP, _U32, _VP, _U64, _PU64],
    "scene_update_sky_params": [_VP, _VP, _VP],
    "build_accel": [_VP, _U32],
    "update_tlas": [_VP],
    "debug_read_bvh": [_VP, _U32, _PU32, _VP, _VP, _VP, _U32, C.POINTER(C.c_int32)],
    "debug_read_wide": [_VP, _VP, _VP, _U32],
    "clear_accum": [_VP],
    "render": [_VP, C.POINTER(Camera), _U32, _U32, C.POINTER(Settings)],
    "resolve": [_VP, _U32, _VP],
    "get_counters": [_VP, C.POINTER(Counters)],
    "reset_counters": [_VP],
    "trace_rays": [_VP, _VP, _U64, _U32, _VP],
    "trace_shadow_rays": [_VP, _VP, _U64, _U32, _VP],
    "debug_capture": [_VP, _U32],
    "debug_read_queue": [_VP, _U32, _U32, _VP, _VP, _VP, _U64, _PU64],
    "render_primary": [_VP, C.POINTER(Camera), _U32, C.POINTER(Settings), _VP, _VP],
    "trace_ao": [_VP, C.POINTER(Camera), _U32, C.POINTER(AoSettings), _VP, _VP, _VP],
    "upscale_half_res": [_VP, C.POINTER(Camera), _U32, _VP, _VP, _VP, _VP],
    "precompute_sky_ibl": [_VP, C.POINTER(SkyIblDesc)],
    "debug_read_sky_ibl": [_VP, _VP, _VP, _VP],
    "trace_reflection": [_VP, C.POINTER(Camera), _U32, C.POINTER(ReflectionSettings), _VP, _VP, _VP, _VP],
    "trace_probes": [_VP, C.POINTER(ProbeVolume), _VP, _U32, _U32, _VP],
    "trace_probes_range": [_VP, C.POINTER(ProbeVolume), _VP, _U32, _U32, _U32, _U32, _VP],
    "blend_probes": [_VP, C.POINTER(ProbeVolume), _VP, _U32, _VP, C.POINTER(ProbeBlend), _VP, _VP],
    "set_ddgi_volume": [_VP, C.POINTER(ProbeVolume), C.POINTER(ProbeBlend), _VP, _VP],
    "ddgi_lighting": [_VP, _U64, _VP, _VP, _VP, _VP],
}
# Exported by the CUDA library only.
BPT_ONLY_API = {
    "set_stream": [_VP, _VP],
    "set_wave_budget": [_VP, _U64],
    "sync": [_VP],
    "resolve_device": [_VP, _U32, _VP],
    "resolve_device_rgba16f": [_VP, _U32, _VP],
    "accumulate_ahead_rgba16f": [_VP, _U32, _VP],
    "accum_device_ptr": [_VP, C.POINTER(_VP)],
    "upload_accum": [_VP, _VP],
    "post_process": [_VP, C.POINTER(PostSettings), _U32, _VP],
    "post_process_device": [_VP, C.POINTER(PostSettings), _U32, _VP],
    "render_ahead": [_VP, C.POINTER(Camera), _U32, _U32, C.POINTER(Settings), _PU32],
    "accumulate_ahead": [_VP, _U32],
    "pending_ahead": [_VP, _PU32, _PU32],
    "comm_unique_id": [_VP],
    "comm_init": [_VP, _VP, C.c_int, C.c_int],
    "comm_attach": [_VP, _VP],
    "comm_destroy": [_VP],
    "reduce": [_VP, C.c_int],
    "profile_enable": [_VP, _U32],
    "profile_read": [_VP, C.POINTER(KernelTimes)],
}
BPT_EXPORTS = sorted(["bpt_" + n for n in list(COMMON_API) + list(BPT_ONLY_API)] + ["bpt_last_error", "bpt_version"])


class Library:
    """A loaded implementation of the boundary: (shared object, symbol prefix)."""

    def __init__(self, path: str, prefix: str, extra_api: dict | None = None):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no fallback implementation.")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        api = dict(COMMON_API)
        if extra_api:
            api.update(extra_api)
        for name, argtypes in api.items():
            fn = getattr(self.lib, prefix + name)
            fn.argtypes, fn.restype = argtypes, C.c_int
        le = getattr(self.lib, prefix + "last_error")
        le.argtypes, le.restype = [_VP], C.c_char_p

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def comm_unique_id(self) -> bytes:
        """128-byte NCCL unique id (bpt_comm_unique_id): rank 0 creates it and hands it to every rank."""
        buf = (C.c_uint8 * 128)()
        st = self.fn("comm_unique_id")(C.cast(buf, C.c_void_p))
        if st != 0:
            raise BptError(st, self.prefix + "comm_unique_id", "NCCL could not be loaded" if st == 6 else "")
        return bytes(buf)


class Context:
    """One bpt_context (or obpt_context): thin, explicit wrapper over the C entry points."""

    def __init__(self, library: Library, width: int, height: int, device: int = 0):
        self.L = library
        self.width, self.height = int(width), int(height)
        self._h = _VP()
        cfg = Config(device=device, width=width, height=height, max_lights_per_vertex=0)
        st = self.L.fn("create")(C.byref(cfg), C.byref(self._h))
        if st != 0:
            raise BptError(st, self.L.prefix + "create", "context creation failed (no CUDA device?)" if st == 5 else "")
        self._keep = []

    # -- plumbing -----------------------------------------------------------------------------
    def _call(self, name, *args):
        st = self.L.fn(name)(self._h, *args)
        if st != 0:
            detail = self.L.fn("last_error")(self._h)
            raise BptError(st, self.L.prefix + name, detail.decode() if detail else "")

    def close(self):
        if self._h:
            self.L.fn("destroy")(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ----------------------------------------------------------------------------------
    def upload_scene(self, scene, accel_mode: int = ACCEL_TWO_LEVEL):
        """`scene` is a scenes.SceneData. Uploads everything and builds the acceleration structure."""
        g = GeometryStreams()
        for name, arr in (("positions", scene.positions), ("normals", scene.normals), ("tangents", scene.tangents),
                          ("colors", getattr(scene, "colors", None)), ("texcoords", scene.texcoords), ("texcoords2", None)):
            setattr(g, name, _ptr(arr))
            field = {"positions": "num_position_floats", "normals": "num_normal_floats", "tangents": "num_tangent_floats",
                     "colors": "num_color_floats", "texcoords": "num_texcoord_floats", "texcoords2": "num_texcoord2_floats"}[name]
            setattr(g, field, 0 if arr is None else arr.size)
        g.indices, g.num_indices = _ptr(scene.indices), scene.indices.size
        self._call("scene_upload_geometry", C.byref(g), _ptr(scene.drawables), _ptr(scene.drawable_va),
                   len(scene.drawables), _ptr(scene.blas), len(scene.blas))
        texs = (TextureDesc * max(1, len(scene.textures)))()
        for i, t in enumerate(scene.textures):
            texs[i] = TextureDesc(texels=_ptr(t["texels"]), width=t["width"], height=t["height"], format=t["format"],
                                  address_mode_u=t.get("address_u", ADDRESS_REPEAT), address_mode_v=t.get("address_v", ADDRESS_REPEAT),
                                  filter_linear=t.get("linear", 1))
        self._call("scene_upload_materials", _ptr(scene.materials), len(scene.materials), texs, len(scene.textures))
        self.upload_instances(scene.instances)
        self.upload_lights(scene)
        self.upload_light_textures(getattr(scene, "light_textures", []))
        self.upload_sky(scene)
        self._call("build_accel", accel_mode)

    def upload_instances(self, instances):
        self._call("scene_upload_instances", _ptr(instances), len(instances))

    def upload_lights(self, scene):
        luts = LtcLuts()
        if scene.ltc_luts is not None:
            luts.matrix_lut0, luts.matrix_lut1, luts.matrix_lut2, luts.norm_lut = (_ptr(a) for a in scene.ltc_luts)
        self._call("scene_upload_lights", _ptr(scene.dir_lights), len(scene.dir_lights), _ptr(scene.point_lights),
                   len(scene.point_lights), _ptr(scene.rect_lights), len(scene.rect_lights), C.byref(luts))

    def denoise_reblur(self, camera: Camera, frame_count: int, noised, hit_positions, depth, normal_roughness, velocity=None, history_validation=None,
                       settings: "ReblurSettings | None" = None) -> np.ndarray:
        """ReblurPass::render on one frame (bpt_denoise_reblur). noised / hit_positions: (h, w, 4) float32 at the context's extent or half of it;
        depth (H, W), normal_roughness (H, W, 4), velocity (H, W, 2) or None, history_validation (h, w) uint8 or None. Returns (h, w, 4)."""
        noised = np.ascontiguousarray(noised, f32); hit_positions = np.ascontiguousarray(hit_positions, f32)
        depth = np.ascontiguousarray(depth, f32); normal_roughness = np.ascontiguousarray(normal_roughness, f32)
        velocity = None if velocity is None else np.ascontiguousarray(velocity, f32)
        history_validation = None if history_validation is None else np.ascontiguousarray(history_validation, np.uint8)
        h, w = noised.shape[:2]
        assert depth.shape == (self.height, self.width) and normal_roughness.shape == (self.height, self.width, 4) and hit_positions.shape == noised.shape
        ins = ReblurInputs(w, h, _ptr(noised), _ptr(hit_positions), _ptr(depth), _ptr(normal_roughness), _ptr(velocity), _ptr(history_validation))
        st = settings or ReblurSettings()
        out = np.zeros((h, w, 4), f32)
        self._call("denoise_reblur", C.byref(camera), frame_count, C.byref(st), C.byref(ins), _ptr(out))
        return out

    def reblur_reset(self):
        self._call("reblur_reset")

    def read_reblur(self, which: int, w: int, h: int) -> np.ndarray:
        """Working textures of the last denoise_reblur (bpt_debug_read_reblur): 0 lighting_dist_0 + mips, 1 lighting_dist_1, 2 accumulation, 3 linear depth + mips."""
        chain = sum((w >> l) * (h >> l) for l in range(4))
        n = {0: chain * 4, 1: w * h * 4, 2: w * h, 3: chain}[which]
        out = np.zeros(n, f32)
        self._call("debug_read_reblur", which, _ptr(out), n)
        return out

    def upload_light_textures(self, textures):
        """`textures`: list of dicts {texels (H, W, 4) uint8 or float32, format, levels, address_u, address_v, linear, mip_linear};
        entry k is what bpt_rect_light_data.texture_index == k samples."""
        descs = (LightTextureDesc * max(1, len(textures)))()
        keep = []
        for i, t in enumerate(textures):
            tex = np.ascontiguousarray(t["texels"]); keep.append(tex)
            descs[i] = LightTextureDesc(texels=_ptr(tex), width=tex.shape[1], height=tex.shape[0], format=t["format"], levels=t.get("levels", 1),
                                        address_mode_u=t.get("address_u", ADDRESS_CLAMP), address_mode_v=t.get("address_v", ADDRESS_CLAMP),
                                        filter_linear=t.get("linear", 1), mip_linear=t.get("mip_linear", 0))
        self._call("scene_upload_light_textures", descs, len(textures))

    def read_light_texture(self, index: int):
        """The generated mip chain of light texture `index`: list of (h, w, 4) float32 arrays, level 0 first."""
        n = C.c_uint64(0)
        self._call("debug_read_light_texture", index, None, 0, C.byref(n))
        flat = np.zeros((n.value, 4), f32)
        self._call("debug_read_light_texture", index, _ptr(flat), n.value, C.byref(n))
        return flat

    def upload_sky(self, scene):
        xf = np.ascontiguousarray(scene.sky_transform, dtype=f32)
        col = np.ascontiguousarray(scene.sky_color, dtype=f32)
        size = 0 if scene.sky_faces is None else scene.sky_faces.shape[1]
        self._call("scene_upload_sky", _ptr(scene.sky_faces), size, _ptr(xf), _ptr(col))

    def update_sky_params(self, transform, color):
        self._call("scene_update_sky_params", _ptr(np.ascontiguousarray(transform, dtype=f32)), _ptr(np.ascontiguousarray(color, dtype=f32)))

    def build_accel(self, mode=ACCEL_TWO_LEVEL):
        self._call("build_accel", mode)

    def update_tlas(self):
        self._call("update_tlas")

    def read_bvh(self, which: int):
        n, root = C.c_uint32(), C.c_int32()
        self._call("debug_read_bvh", which, C.byref(n), None, None, None, 0, C.byref(root))
        morton = np.zeros(n.value, dtype=u64)
        prims = np.zeros(n.value, dtype=u32)
        nodes = np.zeros(max(n.value - 1, 0), dtype=BVH_NODE)
        self._call("debug_read_bvh", which, C.byref(n), _ptr(morton), _ptr(prims), _ptr(nodes) if len(nodes) else None,
                   n.value, C.byref(root))
        return {"n": n.value, "root": root.value, "morton": morton, "prims": prims, "nodes": nodes}

    # -- render ---------------------------------------------------------------------------------
    def read_wide(self):
        """Merged mode: the 4-wide quantised nodes ((n-1, 16) uint32, raw bits of the 64-B records) and the exact leaf boxes ((n, 8) float32)."""
        n = self.read_bvh(0)["n"]
        wide = np.zeros((max(n - 1, 0), 16), dtype=np.uint32); leafbox = np.zeros((n, 8), dtype=f32)
        self._call("debug_read_wide", _ptr(wide), _ptr(leafbox), n)
        return wide, leafbox

    def clear_accum(self):
        self._call("clear_accum")

    def render(self, camera: Camera, frame_index_first: int, num_samples: int, settings: Settings):
        self._call("render", C.byref(camera), frame_index_first, num_samples, C.byref(settings))

    def render_ahead(self, camera: Camera, frame_index_first: int, max_samples: int, settings: Settings) -> int:
        """Traces up to max_samples consecutive samples in one wave and keeps them; returns how many (bpt_render_ahead)."""
        n = C.c_uint32(0)
        self._call("render_ahead", C.byref(camera), frame_index_first, max_samples, C.byref(settings), C.byref(n))
        return n.value

    def accumulate_ahead(self, count: int = 1):
        self._call("accumulate_ahead", count)

    def accumulate_ahead_rgba16f(self, total_samples: int, device_ptr: int):
        """accumulate_ahead(1) + resolve_device_rgba16f(total_samples, device_ptr) in one launch."""
        self._call("accumulate_ahead_rgba16f", total_samples, _VP(device_ptr))

    def pending_ahead(self):
        """(samples traced ahead and not yet accumulated, frame_index of the next one) — bpt_pending_ahead."""
        pending, nxt = C.c_uint32(0), C.c_uint32(0)
        self._call("pending_ahead", C.byref(pending), C.byref(nxt))
        return pending.value, nxt.value

    def resolve(self, total_samples: int) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=f32)
        self._call("resolve", total_samples, _ptr(out))
        return out

    def counters(self) -> Counters:
        c = Counters()
        self._call("get_counters", C.byref(c))
        return c

    def reset_counters(self):
        self._call("reset_counters")

    def trace_rays(self, rays: np.ndarray, frame_index: int = 0) -> np.ndarray:
        assert rays.dtype == RAY
        hits = np.zeros(len(rays), dtype=HIT)
        self._call("trace_rays", _ptr(rays), len(rays), frame_index, _ptr(hits))
        return hits

    def trace_shadow_rays(self, rays: np.ndarray, frame_index: int = 0) -> np.ndarray:
        assert rays.dtype == RAY
        vis = np.zeros(len(rays), dtype=np.uint8)
        self._call("trace_shadow_rays", _ptr(rays), len(rays), frame_index, _ptr(vis))
        return vis

    def render_primary(self, camera: Camera, frame_index: int, settings: Settings):
        """OutputData{depth, gbuffer} of the pass: (H, W) float32 reverse-Z depth and (H, W) GBUFFER_TEXEL records."""
        depth = np.zeros((self.height, self.width), dtype=f32)
        gbuffer = np.zeros((self.height, self.width), dtype=GBUFFER_TEXEL)
        self._call("render_primary", C.byref(camera), frame_index, C.byref(settings), _ptr(depth), _ptr(gbuffer))
        return depth, gbuffer

    def trace_ao(self, camera: Camera, frame_index: int, depth: np.ndarray, normal_roughness: np.ndarray, range_: float = 0.5, strength: float = 0.5,
                 half_resolution: bool = True) -> np.ndarray:
        """Ray-traced ambient occlusion: (ah, aw, 2) float32 = (ao, valid)."""
        ao = AoSettings(range_, strength, 1 if half_resolution else 0)
        ah, aw = (self.height // 2, self.width // 2) if half_resolution else (self.height, self.width)
        out = np.zeros((ah, aw, 2), dtype=f32)
        d = np.ascontiguousarray(depth, dtype=f32); nr = np.ascontiguousarray(normal_roughness, dtype=f32)
        assert d.shape == (self.height, self.width) and nr.shape == (self.height, self.width, 4)
        self._call("trace_ao", C.byref(camera), frame_index, C.byref(ao), _ptr(d), _ptr(nr), _ptr(out))
        return out

    def upscale_half_res(self, camera: Camera, frame_index: int, depth: np.ndarray, normal_roughness: np.ndarray, half: np.ndarray) -> np.ndarray:
        """simple_upscale_cs ("RTR Upscale Hit / Color"): ((H+1)/2, (W+1)/2, 4) -> (H, W, 4)."""
        out = np.zeros((self.height, self.width, 4), dtype=f32)
        self._call("upscale_half_res", C.byref(camera), frame_index, _ptr(np.ascontiguousarray(depth, dtype=f32)),
                   _ptr(np.ascontiguousarray(normal_roughness, dtype=f32)), _ptr(np.ascontiguousarray(half, dtype=f32)), _ptr(out))
        return out

    def precompute_sky_ibl(self, desc: SkyIblDesc | None = None):
        """SkyboxPrecomputePass::render on the uploaded sky; returns (diffuse (6, d, d, 4), specular [level] (6, s_l, s_l, 4), brdf lut (n, n, 2))."""
        desc = desc or SkyIblDesc()
        self._call("precompute_sky_ibl", C.byref(desc))
        diffuse = np.zeros((6, desc.diffuse_size, desc.diffuse_size, 4), f32)
        spec = np.zeros((desc.specular_texels(), 4), f32)
        brdf = np.zeros((desc.brdf_lut_size, desc.brdf_lut_size, 2), f32)
        self._call("debug_read_sky_ibl", _ptr(diffuse), _ptr(spec), _ptr(brdf))
        levels, off = [], 0
        for l in range(desc.specular_levels):
            s_ = desc.specular_size >> l
            levels.append(spec[off:off + 6 * s_ * s_].reshape(6, s_, s_, 4)); off += 6 * s_ * s_
        return diffuse, levels, brdf

    def trace_reflection(self, camera: Camera, frame_index: int, depth: np.ndarray, gbuffer: np.ndarray, settings: ReflectionSettings | None = None):
        """ReflectionPass::render_raytraced (reflection.cpp:317-450) from the camera's depth + G-buffer (render_primary's outputs).
        Returns (reflection colour (rh, rw, 4), hit positions (rh, rw, 4))."""
        settings = settings or ReflectionSettings()
        rh, rw = ((self.height + 1) // 2, (self.width + 1) // 2) if settings.half_resolution else (self.height, self.width)
        refl = np.zeros((rh, rw, 4), dtype=f32); hit = np.zeros((rh, rw, 4), dtype=f32)
        self._call("trace_reflection", C.byref(camera), frame_index, C.byref(settings), _ptr(np.ascontiguousarray(depth, dtype=f32)),
                   _ptr(np.ascontiguousarray(gbuffer, dtype=GBUFFER_TEXEL)), _ptr(refl), _ptr(hit))
        return refl, hit

    def trace_probes(self, volume: ProbeVolume, sample_table: np.ndarray, frame_index: int, num_bounces: int) -> np.ndarray:
        """(num_probes * rays_per_probe, 4) float32: radiance rgb + first hit distance (or -1)."""
        n = volume.probe_counts[0] * volume.probe_counts[1] * volume.probe_counts[2] * volume.rays_per_probe
        out = np.zeros((n, 4), dtype=f32)
        table = np.ascontiguousarray(sample_table, dtype=f32)
        assert table.shape == (8192, 2)
        self._call("trace_probes", C.byref(volume), _ptr(table), frame_index, num_bounces, _ptr(out))
        return out

    def trace_probes_range(self, volume: ProbeVolume, sample_table: np.ndarray, frame_index: int, num_bounces: int, first_probe: int, num_probes: int) -> np.ndarray:
        """The rays of probes [first_probe, first_probe + num_probes): (num_probes * rays_per_probe, 4) float32, bit-identical to the
        same rows of trace_probes (the sharding unit of the DDGI update, sharding.probe_range)."""
        out = np.zeros((num_probes * volume.rays_per_probe, 4), dtype=f32)
        table = np.ascontiguousarray(sample_table, dtype=f32)
        assert table.shape == (8192, 2)
        self._call("trace_probes_range", C.byref(volume), _ptr(table), frame_index, num_bounces, first_probe, num_probes, _ptr(out))
        return out

    def trace_probes_range_into(self, volume: ProbeVolume, sample_table: np.ndarray, frame_index: int, num_bounces: int, first_probe: int, num_probes: int,
                                device_ptr: int):
        """trace_probes_range into DEVICE memory (num_probes * rays_per_probe float4 at `device_ptr`): no host round trip."""
        table = np.ascontiguousarray(sample_table, dtype=f32)
        self._call("trace_probes_range", C.byref(volume), _ptr(table), frame_index, num_bounces, first_probe, num_probes, _VP(device_ptr))

    def blend_probes_from_device(self, volume: ProbeVolume, sample_table, frame_index, rays_device_ptr: int, irradiance=None, visibility=None,
                                 irradiance_size=6, visibility_size=14, alpha=0.97):
        """blend_probes with the per-ray results read from DEVICE memory (e.g. the output of an NCCL all-gather)."""
        nx, ny, nz = volume.probe_counts
        bl = ProbeBlend(irradiance_size, visibility_size, alpha, 0 if irradiance is None else 1)
        irr = np.zeros((nz * (irradiance_size + 2), nx * ny * (irradiance_size + 2), 4), f32) if irradiance is None else np.ascontiguousarray(irradiance, f32).copy()
        vis = np.zeros((nz * (visibility_size + 2), nx * ny * (visibility_size + 2), 2), f32) if visibility is None else np.ascontiguousarray(visibility, f32).copy()
        self._call("blend_probes", C.byref(volume), _ptr(np.ascontiguousarray(sample_table, f32)), frame_index, _VP(rays_device_ptr), C.byref(bl), _ptr(irr), _ptr(vis))
        return irr, vis

    def set_ddgi_volume(self, volume: ProbeVolume | None, irradiance=None, visibility=None, irradiance_size=6, visibility_size=14):
        """Binds the atlases of the previous DDGI update (None unbinds): feedback for trace_probes, input of ddgi_lighting."""
        if volume is None:
            self._call("set_ddgi_volume", None, None, None, None)
            return
        bl = ProbeBlend(irradiance_size, visibility_size, 0.0, 0)
        irr = np.ascontiguousarray(irradiance, dtype=f32); vis = np.ascontiguousarray(visibility, dtype=f32)
        self._call("set_ddgi_volume", C.byref(volume), C.byref(bl), _ptr(irr), _ptr(vis))

    def ddgi_lighting(self, position, normal, view) -> np.ndarray:
        """calc_ddgi_volume_lighting for n points: (n, 4) float32 = (irradiance rgb, 1) or 0 outside the volume."""
        p = np.ascontiguousarray(position, dtype=f32); n = np.ascontiguousarray(normal, dtype=f32); v = np.ascontiguousarray(view, dtype=f32)
        assert p.shape == n.shape == v.shape and p.shape[1] == 3
        out = np.zeros((len(p), 4), dtype=f32)
        self._call("ddgi_lighting", len(p), _ptr(p), _ptr(n), _ptr(v), _ptr(out))
        return out

    def blend_probes(self, volume: ProbeVolume, sample_table, frame_index, rays, irradiance=None, visibility=None,
                     irradiance_size=6, visibility_size=14, alpha=0.97):
        """Returns (irradiance atlas (H, W, 4), visibility atlas (H, W, 2)); pass the previous atlases to blend temporally."""
        nx, ny, nz = volume.probe_counts[0], volume.probe_counts[1], volume.probe_counts[2]
        hist = irradiance is not None and visibility is not None
        irr = np.ascontiguousarray(irradiance, f32).copy() if hist else np.zeros((nz * (irradiance_size + 2), nx * ny * (irradiance_size + 2), 4), f32)
        vis = np.ascontiguousarray(visibility, f32).copy() if hist else np.zeros((nz * (visibility_size + 2), nx * ny * (visibility_size + 2), 2), f32)
        bl = ProbeBlend(irradiance_size, visibility_size, alpha, 1 if hist else 0)
        self._call("blend_probes", C.byref(volume), _ptr(np.ascontiguousarray(sample_table, f32)), frame_index,
                   _ptr(np.ascontiguousarray(rays, f32)), C.byref(bl), _ptr(irr), _ptr(vis))
        return irr, vis

    def debug_capture(self, enable: bool):
        self._call("debug_capture", 1 if enable else 0)

    def read_queue(self, bounce: int, kind: int):
        n = C.c_uint64()
        self._call("debug_read_queue", bounce, kind, None, None, None, 0, C.byref(n))
        pixels = np.zeros(n.value, dtype=u32)
        lights = np.zeros(n.value, dtype=u32) if kind == 1 else None
        hits = np.zeros(n.value, dtype=HIT) if kind == 0 else None
        if n.value:
            self._call("debug_read_queue", bounce, kind, _ptr(pixels), _ptr(lights), _ptr(hits), n.value, C.byref(n))
        return {"pixels": pixels, "lights": lights, "hits": hits}

    # -- CUDA-library-only -----------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle: int):
        self._call("set_stream", _VP(cuda_stream_handle))

    def sync(self):
        self._call("sync")

    def accum_device_ptr(self) -> int:
        p = _VP()
        self._call("accum_device_ptr", C.byref(p))
        return p.value

    def resolve_device(self, total_samples: int, device_ptr: int):
        self._call("resolve_device", total_samples, _VP(device_ptr))

    def set_wave_budget(self, max_paths_in_flight: int):
        """Caps the wavefront footprint: paths in flight per wave (bpt_set_wave_budget; 0 = default 2^26)."""
        self._call("set_wave_budget", max_paths_in_flight)

    def resolve_device_rgba16f(self, total_samples: int, device_ptr: int):
        """The resolved image as rgba16_sfloat (the reference's OutputData.color format), 8 bytes per pixel, device memory."""
        self._call("resolve_device_rgba16f", total_samples, _VP(device_ptr))

    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        """Joins the NCCL communicator identified by `unique_id` (128 bytes from Library.comm_unique_id() on rank 0)."""
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._call("comm_init", C.cast(buf, C.c_void_p), rank, world_size)

    def reduce(self, root: int = 0):
        """In-place NCCL sum of every rank's FP32 accumulation buffer into `root` (stream-ordered)."""
        self._call("reduce", root)

    def profile_enable(self, enable: bool):
        self._call("profile_enable", 1 if enable else 0)

    def profile_read(self) -> KernelTimes:
        t = KernelTimes()
        self._call("profile_read", C.byref(t))
        return t

    def upload_accum(self, sums: np.ndarray):
        self._call("upload_accum", _ptr(np.ascontiguousarray(sums, dtype=f32)))

    def post_process(self, settings: PostSettings, total_samples: int) -> np.ndarray:
        """PostProcessPass::render on the accumulated image (bloom + output pass): H x W x 4 float32."""
        out = np.zeros((self.height, self.width, 4), f32)
        self._call("post_process", C.byref(settings), total_samples, _ptr(out))
        return out

    def post_process_device(self, settings: PostSettings, total_samples: int, device_ptr: int):
        self._call("post_process_device", C.byref(settings), total_samples, _VP(device_ptr))
