"""Test-only: host build of the CUDA device functions (tests/hostcheck/hostcheck.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from bisemutum_engine_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
OUT = os.path.join(HERE, "hostcheck", "_build_hostcheck.so")
CSRC = os.path.join(os.path.dirname(HERE), "bisemutum-engine_b200", "csrc")


class HcBvh(C.Structure):
    _fields_ = [("n", C.c_uint32), ("root", C.c_int32), ("nodes", C.c_void_p), ("prims", C.c_void_p)]


class HcLightTex(C.Structure):
    _fields_ = [("chain", C.c_void_p)] + [(n, C.c_uint32) for n in ("w", "h", "levels", "addr_u", "addr_v", "linear", "mip_linear")]


class HcScene(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("normals", C.c_void_p), ("tangents", C.c_void_p), ("texcoords", C.c_void_p),
                ("indices", C.c_void_p),
                ("drawables", C.c_void_p), ("drawable_va", C.c_void_p), ("num_drawables", C.c_uint32),
                ("blas_desc", C.c_void_p), ("num_blas", C.c_uint32),
                ("instances", C.c_void_p), ("num_instances", C.c_uint32),
                ("materials", C.c_void_p), ("num_materials", C.c_uint32),
                ("dir", C.c_void_p), ("num_dir", C.c_uint32),
                ("point", C.c_void_p), ("num_point", C.c_uint32),
                ("rect", C.c_void_p), ("num_rect", C.c_uint32),
                ("ltc_m0", C.c_void_p), ("ltc_m1", C.c_void_p), ("ltc_m2", C.c_void_p), ("ltc_norm", C.c_void_p),
                ("sky_faces", C.c_void_p), ("sky_size", C.c_uint32), ("sky_transform", C.c_float * 9), ("sky_color", C.c_float * 3),
                ("accel_mode", C.c_uint32),
                ("blas_bvh", C.POINTER(HcBvh)), ("tlas", HcBvh),
                ("textures", C.POINTER(capi.TextureDesc)), ("num_textures", C.c_uint32),
                ("light_textures", C.POINTER(HcLightTex)), ("num_light_textures", C.c_uint32), ("colors", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
        if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mavx2",
                            "-fvisibility=hidden", "-I/usr/local/cuda/include", "-o", OUT, SRC], check=True)
        _lib = C.CDLL(OUT)
        _lib.hc_rng_tea.argtypes, _lib.hc_rng_tea.restype = [C.c_uint32, C.c_uint32], C.c_uint32
        _lib.hc_sincos_2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib.hc_atan2.argtypes, _lib.hc_atan2.restype = [C.c_float, C.c_float], C.c_float
        _lib.hc_acos.argtypes, _lib.hc_acos.restype = [C.c_float], C.c_float
        _lib.hc_ggx_vndf_sample.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.hc_surface_eval_lit.argtypes = [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_void_p]
        _lib.hc_q_half.argtypes, _lib.hc_q_half.restype = [C.c_float], C.c_float
        _lib.hc_surface_through_gbuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.hc_render.argtypes = [C.POINTER(HcScene), C.POINTER(capi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.POINTER(capi.Settings), C.c_void_p]
        _lib.hc_render_primary.argtypes = [C.POINTER(HcScene), C.POINTER(capi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(capi.Settings), C.c_void_p, C.c_void_p]
        _lib.hc_trace_ao.argtypes = [C.POINTER(HcScene), C.POINTER(capi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(capi.AoSettings), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.hc_trace_reflection.argtypes = [C.POINTER(HcScene), C.POINTER(capi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(capi.ReflectionSettings),
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.hc_trace_probes.argtypes = [C.POINTER(HcScene), C.POINTER(capi.ProbeVolume), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.hc_blend_probes.argtypes = [C.POINTER(capi.ProbeVolume), C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(capi.ProbeBlend), C.c_void_p, C.c_void_p]
        _lib.hc_set_ddgi.argtypes = [C.POINTER(capi.ProbeVolume), C.POINTER(capi.ProbeBlend), C.c_void_p, C.c_void_p]
        _lib.hc_ddgi_lighting.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.hc_trace.argtypes = [C.POINTER(HcScene), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.hc_trace_wide.argtypes = [C.POINTER(HcScene), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.hc_read_wide.argtypes = [C.POINTER(HcScene), C.c_void_p, C.c_void_p]
        _lib.hc_precompute_sky_ibl.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(capi.SkyIblDesc), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.hc_upscale_half_res.argtypes = [C.POINTER(capi.Camera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.hc_reblur.argtypes = [C.POINTER(capi.Camera), C.c_uint64, C.POINTER(capi.ReblurSettings), C.POINTER(capi.ReblurInputs), C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.hc_reblur_read.argtypes = [C.c_uint32, C.c_void_p]
        _lib.hc_post_process.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(capi.PostSettings), C.c_void_p]
    return _lib


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


class HostScene:
    """hc_scene built from a SceneData and the BVH arrays read back from an oracle context."""

    def __init__(self, scene, oracle_ctx, accel_mode):
        self.keep = []
        h = HcScene()
        h.positions, h.normals, h.tangents, h.texcoords = _p(scene.positions), _p(scene.normals), _p(scene.tangents), _p(scene.texcoords)
        h.indices = _p(scene.indices)
        h.drawables, h.drawable_va, h.num_drawables = _p(scene.drawables), _p(scene.drawable_va), len(scene.drawables)
        h.blas_desc, h.num_blas = _p(scene.blas), len(scene.blas)
        h.instances, h.num_instances = _p(scene.instances), len(scene.instances)
        h.materials, h.num_materials = _p(scene.materials), len(scene.materials)
        h.dir, h.num_dir = _p(scene.dir_lights), len(scene.dir_lights)
        h.point, h.num_point = _p(scene.point_lights), len(scene.point_lights)
        h.rect, h.num_rect = _p(scene.rect_lights), len(scene.rect_lights)
        if scene.ltc_luts is not None:
            h.ltc_m0, h.ltc_m1, h.ltc_m2, h.ltc_norm = (_p(a) for a in scene.ltc_luts)
        if scene.sky_faces is not None:
            h.sky_faces, h.sky_size = _p(scene.sky_faces), scene.sky_faces.shape[1]
        h.sky_transform[:] = list(np.asarray(scene.sky_transform, np.float32))
        h.sky_color[:] = list(np.asarray(scene.sky_color, np.float32))
        h.accel_mode = accel_mode
        nb = len(scene.blas) if accel_mode == capi.ACCEL_TWO_LEVEL else 1
        arr = (HcBvh * nb)()
        for b in range(nb):
            bv = oracle_ctx.read_bvh(b)
            self.keep.append(bv)
            arr[b] = HcBvh(bv["n"], bv["root"], _p(bv["nodes"]), _p(bv["prims"]))
        h.blas_bvh = arr
        if accel_mode == capi.ACCEL_TWO_LEVEL:
            tv = oracle_ctx.read_bvh(capi.BVH_TLAS)
            self.keep.append(tv)
            h.tlas = HcBvh(tv["n"], tv["root"], _p(tv["nodes"]), _p(tv["prims"]))
        self.keep.append(arr)
        texs = (capi.TextureDesc * max(1, len(scene.textures)))()
        for i, t in enumerate(scene.textures):
            texs[i] = capi.TextureDesc(texels=_p(t["texels"]), width=t["width"], height=t["height"], format=t["format"],
                                       address_mode_u=t.get("address_u", 0), address_mode_v=t.get("address_v", 0), filter_linear=t.get("linear", 1))
        h.textures, h.num_textures = texs, len(scene.textures)
        self.keep.append(texs)
        lts = getattr(scene, "light_textures", [])
        ltex = (HcLightTex * max(1, len(lts)))()
        for i, t in enumerate(lts):
            chain = oracle_ctx.read_light_texture(i)                       # the oracle's generated chain (the generator is checked on the GPU)
            self.keep.append(chain)
            hh, ww = t["texels"].shape[:2]
            levels, n = 0, 0
            while n < len(chain):
                n += max(ww >> levels, 1) * max(hh >> levels, 1); levels += 1
            ltex[i] = HcLightTex(_p(chain), ww, hh, levels, t.get("address_u", capi.ADDRESS_CLAMP), t.get("address_v", capi.ADDRESS_CLAMP),
                                 t.get("linear", 1), t.get("mip_linear", 0))
        h.light_textures, h.num_light_textures = ltex, len(lts)
        h.colors = _p(getattr(scene, "colors", None))
        self.keep.append(ltex)
        self.scene = scene
        self.h = h

    def render(self, camera, width, height, frame_first, nsamples, settings):
        accum = np.zeros((height, width, 4), np.float32)
        lib().hc_render(C.byref(self.h), C.byref(camera), width, height, frame_first, nsamples, C.byref(settings), accum.ctypes.data_as(C.c_void_p))
        return accum

    def render_primary(self, camera, width, height, frame_index, settings):
        depth = np.zeros((height, width), np.float32); g = np.zeros((height, width), capi.GBUFFER_TEXEL)
        lib().hc_render_primary(C.byref(self.h), C.byref(camera), width, height, frame_index, C.byref(settings), depth.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p))
        return depth, g

    def trace_ao(self, camera, width, height, frame_index, depth, normal_roughness, range_=0.5, strength=0.5, half_resolution=True):
        ao = capi.AoSettings(range_, strength, 1 if half_resolution else 0)
        ah, aw = (height // 2, width // 2) if half_resolution else (height, width)
        out = np.zeros((ah, aw, 2), np.float32)
        d = np.ascontiguousarray(depth, np.float32); nr = np.ascontiguousarray(normal_roughness, np.float32)
        lib().hc_trace_ao(C.byref(self.h), C.byref(camera), width, height, frame_index, C.byref(ao), d.ctypes.data_as(C.c_void_p), nr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def trace_reflection(self, camera, width, height, frame_index, depth, gbuffer, settings):
        rh, rw = ((height + 1) // 2, (width + 1) // 2) if settings.half_resolution else (height, width)
        refl = np.zeros((rh, rw, 4), np.float32); hit = np.zeros((rh, rw, 4), np.float32)
        d = np.ascontiguousarray(depth, np.float32); g = np.ascontiguousarray(gbuffer, capi.GBUFFER_TEXEL)
        lib().hc_trace_reflection(C.byref(self.h), C.byref(camera), width, height, frame_index, C.byref(settings), d.ctypes.data_as(C.c_void_p),
                                  g.ctypes.data_as(C.c_void_p), refl.ctypes.data_as(C.c_void_p), hit.ctypes.data_as(C.c_void_p))
        return refl, hit

    def trace(self, rays, frame_index=0):
        hits = np.zeros(len(rays), capi.HIT)
        vis = np.zeros(len(rays), np.uint8)
        lib().hc_trace(C.byref(self.h), rays.ctypes.data_as(C.c_void_p), len(rays), frame_index, hits.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p))
        return hits, vis

    def trace_wide(self, rays, frame_index=0):
        """Closest hits and visibility through the 4-wide quantised tree (merged mode)."""
        hits = np.zeros(len(rays), capi.HIT)
        vis = np.zeros(len(rays), np.uint8)
        assert lib().hc_trace_wide(C.byref(self.h), rays.ctypes.data_as(C.c_void_p), len(rays), frame_index, hits.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p)) == 0
        return hits, vis

    def read_wide(self):
        n = self.keep[0]["n"]
        wide = np.zeros((max(n - 1, 0), 16), np.uint32); leafbox = np.zeros((n, 8), np.float32)
        assert lib().hc_read_wide(C.byref(self.h), wide.ctypes.data_as(C.c_void_p), leafbox.ctypes.data_as(C.c_void_p)) == 0
        return wide, leafbox

    def trace_probes(self, volume, table, frame_index, num_bounces):
        n = volume.probe_counts[0] * volume.probe_counts[1] * volume.probe_counts[2] * volume.rays_per_probe
        out = np.zeros((n, 4), np.float32)
        lib().hc_trace_probes(C.byref(self.h), C.byref(volume), table.ctypes.data_as(C.c_void_p), frame_index, num_bounces, out.ctypes.data_as(C.c_void_p))
        return out


def set_ddgi(volume, irr=None, vis=None, irradiance_size=6, visibility_size=14):
    """Binds (or, with volume None, unbinds) the DDGI volume the host build's probe paths and ddgi_lighting read."""
    if volume is None:
        lib().hc_set_ddgi(None, None, None, None)
        return
    bl = capi.ProbeBlend(irradiance_size, visibility_size, 0.0, 0)
    irr = np.ascontiguousarray(irr, np.float32); vis = np.ascontiguousarray(vis, np.float32)
    lib().hc_set_ddgi(C.byref(volume), C.byref(bl), irr.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p))


def ddgi_lighting(position, normal, view):
    p = np.ascontiguousarray(position, np.float32); n = np.ascontiguousarray(normal, np.float32); v = np.ascontiguousarray(view, np.float32)
    out = np.zeros((len(p), 4), np.float32)
    lib().hc_ddgi_lighting(len(p), p.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def blend_probes(volume, table, frame_index, rays, irr, vis, irradiance_size=6, visibility_size=14, alpha=0.97, history_valid=0):
    """Host build of bpt_ddgi.cuh; irr / vis are modified in place (pass copies)."""
    bl = capi.ProbeBlend(irradiance_size, visibility_size, alpha, history_valid)
    lib().hc_blend_probes(C.byref(volume), np.ascontiguousarray(table, np.float32).ctypes.data_as(C.c_void_p), frame_index,
                          np.ascontiguousarray(rays, np.float32).ctypes.data_as(C.c_void_p), C.byref(bl),
                          irr.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p))
    return irr, vis


def post_process(sums, total_samples, settings):
    """bpt_post.cuh's per-pixel functions on the host, driven like post.cu's kernels (sums: H x W x 4 float32)."""
    sums = np.ascontiguousarray(sums, np.float32)
    h, w = sums.shape[:2]
    out = np.zeros_like(sums)
    assert lib().hc_post_process(sums.ctypes.data_as(C.c_void_p), w, h, total_samples, C.byref(settings), out.ctypes.data_as(C.c_void_p)) == 0
    return out


def precompute_sky_ibl(scene, desc):
    """bpt_ibl.cuh's per-texel functions on the host; also binds the result for HostScene.trace_reflection (desc None: unbind)."""
    if desc is None:
        lib().hc_precompute_sky_ibl(None, 0, None, None, None, None)
        return None
    faces = np.ascontiguousarray(scene.sky_faces, np.float32) if scene.sky_faces is not None else None
    size = 0 if faces is None else faces.shape[1]
    diffuse = np.zeros((6, desc.diffuse_size, desc.diffuse_size, 4), np.float32)
    spec = np.zeros((desc.specular_texels(), 4), np.float32)
    brdf = np.zeros((desc.brdf_lut_size, desc.brdf_lut_size, 2), np.float32)
    lib().hc_precompute_sky_ibl(faces.ctypes.data_as(C.c_void_p) if faces is not None else None, size, C.byref(desc),
                                diffuse.ctypes.data_as(C.c_void_p), spec.ctypes.data_as(C.c_void_p), brdf.ctypes.data_as(C.c_void_p))
    levels, off = [], 0
    for l in range(desc.specular_levels):
        s_ = desc.specular_size >> l
        levels.append(spec[off:off + 6 * s_ * s_].reshape(6, s_, s_, 4)); off += 6 * s_ * s_
    return diffuse, levels, brdf


def upscale_half_res(camera, width, height, frame_index, depth, normal_roughness, half):
    d = np.ascontiguousarray(depth, np.float32); nr = np.ascontiguousarray(normal_roughness, np.float32); h = np.ascontiguousarray(half, np.float32)
    out = np.zeros((height, width, 4), np.float32)
    lib().hc_upscale_half_res(C.byref(camera), width, height, frame_index, d.ctypes.data_as(C.c_void_p), nr.ctypes.data_as(C.c_void_p),
                              h.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


class HostReblur:
    """bpt_reblur.cuh's per-pixel passes on the host (hc_reblur), with the same call shape as capi.Context.denoise_reblur."""

    def __init__(self, width, height):
        self.width, self.height = width, height
        lib().hc_reblur_reset()

    def denoise_reblur(self, camera, frame_count, noised, hit_positions, depth, normal_roughness, velocity=None, history_validation=None, settings=None):
        f32 = np.float32
        noised = np.ascontiguousarray(noised, f32); hit_positions = np.ascontiguousarray(hit_positions, f32)
        depth = np.ascontiguousarray(depth, f32); normal_roughness = np.ascontiguousarray(normal_roughness, f32)
        velocity = None if velocity is None else np.ascontiguousarray(velocity, f32)
        history_validation = None if history_validation is None else np.ascontiguousarray(history_validation, np.uint8)
        h, w = noised.shape[:2]
        ins = capi.ReblurInputs(w, h, _p(noised), _p(hit_positions), _p(depth), _p(normal_roughness), _p(velocity), _p(history_validation))
        st = settings or capi.ReblurSettings()
        out = np.zeros((h, w, 4), f32)
        lib().hc_reblur(C.byref(camera), frame_count, C.byref(st), C.byref(ins), self.width, self.height, out.ctypes.data_as(C.c_void_p))
        self._wh = (w, h)
        return out

    def reblur_reset(self):
        lib().hc_reblur_reset()

    def read_reblur(self, which, w, h):
        chain = sum((w >> l) * (h >> l) for l in range(4))
        n = {0: chain * 4, 1: w * h * 4, 2: w * h, 3: chain}[which]
        out = np.zeros(n, np.float32)
        lib().hc_reblur_read(which, out.ctypes.data_as(C.c_void_p))
        return out
