"""Seeded procedural scenes of the shapes BASELINE.json names (synthetic input only — no compute).

Every generator returns a `SceneData` whose arrays are laid out exactly as the C ABI wants them
(flat float/uint streams + DrawableSbtData / BLAS / instance records, include/bpt/bpt.h), i.e. the
form the reference's GpuSceneSystem hands to its ray-tracing shaders
(bisemutum/src/graphics/gpu_scene_data.hpp:10-24, drawable_stb_data.hpp:7-17).

Importer semantics that the generators follow so the data looks like an imported glTF
(bisemutum/src/scene_basic/menu_actions/import_model.cpp:27-430): POSITION/NORMAL/TEXCOORD_0 +
per-vertex tangents with handedness in w, one drawable per (node, primitive), materials from the
fixed metallic-roughness template (:208-230).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi

f32, u32 = np.float32, np.uint32


@dataclass
class SceneData:
    name: str
    positions: np.ndarray
    normals: np.ndarray
    tangents: np.ndarray
    texcoords: np.ndarray
    indices: np.ndarray
    drawables: np.ndarray
    drawable_va: np.ndarray
    blas: np.ndarray
    instances: np.ndarray
    materials: np.ndarray
    textures: list = field(default_factory=list)
    colors: np.ndarray | None = None      # vertex colours, 3 floats per vertex (drawable color_offset), optional
    dir_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, capi.DIR_LIGHT))
    point_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, capi.POINT_LIGHT))
    rect_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, capi.RECT_LIGHT))
    ltc_luts: tuple | None = None
    light_textures: list = field(default_factory=list)   # rect-light textures: dicts for capi.Context.upload_light_textures
    sky_faces: np.ndarray | None = None
    sky_transform: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=f32).reshape(9))
    sky_color: np.ndarray = field(default_factory=lambda: np.ones(3, f32))
    camera: dict = field(default_factory=dict)
    bounds: tuple | None = None  # (lo, hi) world bounds, for light/probe placement

    @property
    def num_triangles(self) -> int:
        """Instanced triangle count (what a ray can hit)."""
        return int(sum(int(self.blas["num_triangles"][int(b)]) for b in self.instances["blas"]))


# ---------------------------------------------------------------------------------------------
# mesh helpers
# ---------------------------------------------------------------------------------------------
def _normalize(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    n[n == 0] = 1.0
    return v / n


def grid_patch(nu, nv, pos_fn):
    """(nu x nv) quad grid; pos_fn(u, v) -> (..., 3) with u, v in [0, 1]. Normals/tangents from
    central differences of pos_fn (tangent = dP/du, handedness w from the uv orientation)."""
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")           # (nv+1, nu+1)
    P = pos_fn(uu, vv).astype(np.float64)
    e = 1e-4
    dPu = (pos_fn(uu + e, vv) - pos_fn(uu - e, vv)) / (2 * e)
    dPv = (pos_fn(uu, vv + e) - pos_fn(uu, vv - e)) / (2 * e)
    N = _normalize(np.cross(dPu, dPv))
    T = _normalize(dPu - N * np.sum(N * dPu, -1, keepdims=True))
    w = np.sign(np.sum(np.cross(N, T) * dPv, -1, keepdims=True))
    w[w == 0] = 1.0
    tang = np.concatenate([T, w], -1)
    uv = np.stack([uu, vv], -1)
    i0 = (np.arange(nv)[:, None] * (nu + 1) + np.arange(nu)[None, :]).reshape(-1)
    tri = np.stack([np.stack([i0, i0 + 1, i0 + nu + 2], -1), np.stack([i0, i0 + nu + 2, i0 + nu + 1], -1)], 1).reshape(-1, 3)
    return (P.reshape(-1, 3).astype(f32), N.reshape(-1, 3).astype(f32), tang.reshape(-1, 4).astype(f32),
            uv.reshape(-1, 2).astype(f32), tri.astype(u32))


def merge_meshes(meshes):
    P, N, T, UV, I = [], [], [], [], []
    base = 0
    for (p, n, t, uv, i) in meshes:
        P.append(p); N.append(n); T.append(t); UV.append(uv); I.append(i + base)
        base += len(p)
    return (np.concatenate(P), np.concatenate(N), np.concatenate(T), np.concatenate(UV), np.concatenate(I).astype(u32))


def box_mesh(lo, hi, n=4):
    """Axis-aligned box, outward normals, each face an n x n grid."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    faces = []
    for axis in range(3):
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        for side in (0, 1):
            def fn(u, v, axis=axis, a1=a1, a2=a2, side=side):
                p = np.zeros(u.shape + (3,))
                p[..., axis] = hi[axis] if side else lo[axis]
                uu = u if side else 1.0 - u     # keep outward orientation
                p[..., a1] = lo[a1] + (hi[a1] - lo[a1]) * uu
                p[..., a2] = lo[a2] + (hi[a2] - lo[a2]) * v
                return p
            faces.append(grid_patch(n, n, fn))
    return merge_meshes(faces)


def cylinder_mesh(radius, height, nseg, nring, bulge=0.0):
    def fn(u, v):
        ang = 2 * np.pi * u
        r = radius * (1.0 + bulge * np.sin(np.pi * v) ** 2 + 0.03 * np.cos(8 * ang) * 1.0)
        return np.stack([r * np.cos(ang), height * v, -r * np.sin(ang)], -1)
    return grid_patch(nseg, nring, fn)


def sphere_mesh(radius, nseg, nring, bumps=0.0, seed=0):
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 4)

    def fn(u, v):
        th = np.pi * (0.02 + 0.96 * v)
        ang = 2 * np.pi * u
        r = radius * (1.0 + bumps * (np.sin(5 * ang + ph[0]) * np.sin(4 * th + ph[1]) + 0.5 * np.sin(9 * ang + ph[2]) * np.sin(7 * th + ph[3])))
        return np.stack([r * np.sin(th) * np.cos(ang), -r * np.cos(th), -r * np.sin(th) * np.sin(ang)], -1)
    return grid_patch(nseg, nring, fn)


class SceneBuilder:
    """Collects meshes (→ flat streams + one BLAS each), materials and drawables (→ instances)."""

    def __init__(self, name):
        self.name = name
        self.P, self.N, self.T, self.UV, self.I = [], [], [], [], []
        self.nverts = 0
        self.nidx = 0
        self.meshes = []      # (vertex_base, index_offset, num_tris)
        self.materials = []
        self.drawables = []   # (mesh, material, transform 3x4, opaque)
        self.textures = []

    def add_mesh(self, mesh):
        p, n, t, uv, idx = mesh
        self.P.append(p); self.N.append(n); self.T.append(t); self.UV.append(uv); self.I.append(idx.reshape(-1))
        self.meshes.append((self.nverts, self.nidx, len(idx)))
        self.nverts += len(p)
        self.nidx += idx.size
        return len(self.meshes) - 1

    def add_material(self, base_color=(0.5, 0.5, 0.5, 1.0), roughness=0.5, metallic=0.0, two_sided=False,
                     blend=capi.BLEND_OPAQUE, kind=capi.MATERIAL_KIND_GLTF_PBR, model=capi.SURFACE_MODEL_LIT,
                     normal_map_scale=1.0, occlusion_strength=1.0, base_color_tex=-1, metallic_roughness_tex=-1,
                     normal_map_tex=-1, occlusion_tex=-1):
        m = np.zeros((), capi.MATERIAL)
        bc = list(base_color) + [1.0] * (4 - len(base_color))
        m["base_color"] = bc
        m["roughness"], m["metallic"] = roughness, metallic
        m["normal_map_scale"], m["occlusion_strength"] = normal_map_scale, occlusion_strength
        m["flags"] = capi.material_flags(kind, blend, model, two_sided)
        m["base_color_tex"], m["metallic_roughness_tex"] = base_color_tex, metallic_roughness_tex
        m["normal_map_tex"], m["occlusion_tex"] = normal_map_tex, occlusion_tex
        self.materials.append(m)
        return len(self.materials) - 1

    def add_texture(self, texels, fmt=capi.TEXTURE_RGBA8_UNORM, address=capi.ADDRESS_REPEAT, linear=1):
        texels = np.ascontiguousarray(texels)
        self.textures.append({"texels": texels, "width": texels.shape[1], "height": texels.shape[0], "format": fmt,
                              "address_u": address, "address_v": address, "linear": linear})
        return len(self.textures) - 1

    def add_drawable(self, mesh, material, transform=None):
        xf = np.eye(4, dtype=np.float64)[:3] if transform is None else np.asarray(transform, np.float64)[:3]
        self.drawables.append((mesh, material, xf.astype(f32)))
        return len(self.drawables) - 1

    def finish(self, **kw) -> SceneData:
        nd = len(self.drawables)
        drawables = np.zeros(nd, capi.DRAWABLE_SBT)
        instances = np.zeros(nd, capi.INSTANCE_DESC)
        blas = np.zeros(len(self.meshes), capi.BLAS_DESC)
        for b, (vb, io, nt) in enumerate(self.meshes):
            blas[b] = (vb * 3, io, nt, 0)
        mats = np.array(self.materials, dtype=capi.MATERIAL)
        for i, (mesh, mat, xf) in enumerate(self.drawables):
            vb, io, nt = self.meshes[mesh]
            drawables[i] = (i, vb * 3, vb * 3, vb * 4, 0, vb * 2, 0, io, mat * capi.MATERIAL.itemsize)
            blend = (int(mats[mat]["flags"]) >> 16) & 0xff
            flag = capi.INSTANCE_FORCE_OPAQUE if blend == capi.BLEND_OPAQUE else capi.INSTANCE_FORCE_NON_OPAQUE
            instances[i]["transform"] = xf
            instances[i]["instance_id_and_mask"] = i | (0xff << 24)
            instances[i]["sbt_offset_and_flags"] = i | (flag << 24)
            instances[i]["blas"] = mesh
        va = np.full(nd, capi.VA_POSITION | capi.VA_NORMAL | capi.VA_TANGENT | capi.VA_TEXCOORD, u32)
        return SceneData(
            name=self.name,
            positions=np.ascontiguousarray(np.concatenate(self.P).reshape(-1), f32),
            normals=np.ascontiguousarray(np.concatenate(self.N).reshape(-1), f32),
            tangents=np.ascontiguousarray(np.concatenate(self.T).reshape(-1), f32),
            texcoords=np.ascontiguousarray(np.concatenate(self.UV).reshape(-1), f32),
            indices=np.ascontiguousarray(np.concatenate(self.I), u32),
            drawables=drawables, drawable_va=va, blas=blas, instances=instances, materials=mats,
            textures=self.textures, **kw)


def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def rotate_y(a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def scale(sx, sy=None, sz=None):
    sy = sx if sy is None else sy
    sz = sx if sz is None else sz
    return np.diag([sx, sy, sz, 1.0])


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    return m


def dir_light(direction_to_light, color=(1.0, 0.9, 0.8), strength=4.0):
    """DirLightData as LightsContext::collect_all_lights packs it (lights.cpp:52-63) with
    cast_shadow irrelevant (shadow rays replace shadow maps)."""
    l = np.zeros(1, capi.DIR_LIGHT)
    d = np.asarray(direction_to_light, np.float64)
    l["emission"] = (np.asarray(color, f32) * f32(strength)).astype(f32)
    l["direction"] = (d / np.linalg.norm(d)).astype(f32)
    l["sm_index"] = -1
    return l


def procedural_sky(size=256, sun_dir=(0.3, 1.0, 0.2), seed=0):
    """Gradient + sun-lobe cubemap, 6 x size x size x 4 float32, Vulkan face order."""
    s = (np.arange(size) + 0.5) / size * 2.0 - 1.0
    ux, uy = np.meshgrid(s, s, indexing="xy")       # uv.x along columns, uv.y along rows
    one = np.ones_like(ux)
    dirs = [np.stack([one, -uy, -ux], -1), np.stack([-one, -uy, ux], -1), np.stack([ux, one, uy], -1),
            np.stack([ux, -one, -uy], -1), np.stack([ux, -uy, one], -1), np.stack([-ux, -uy, -one], -1)]
    sd = np.asarray(sun_dir, np.float64)
    sd /= np.linalg.norm(sd)
    faces = np.zeros((6, size, size, 4), f32)
    for f, d in enumerate(dirs):
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        t = np.clip(d[..., 1] * 0.5 + 0.5, 0, 1)[..., None]
        horizon, zenith = np.array([0.85, 0.9, 1.0]), np.array([0.25, 0.45, 0.9])
        col = horizon * (1 - t) + zenith * t
        col = np.where(d[..., 1:2] < 0, np.array([0.18, 0.16, 0.14]) * (1 + d[..., 1:2] * 0.5), col)
        lobe = np.clip(np.sum(d * sd, -1), 0, 1)[..., None]
        col = col + np.array([1.0, 0.9, 0.7]) * (lobe ** 64) * 4.0
        faces[f, ..., :3] = col
        faces[f, ..., 3] = 1.0
    return faces


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: Cornell box, 512x512, 16 spp, max depth 5, one directional light
# ---------------------------------------------------------------------------------------------
def cornell_box(tess=16) -> SceneData:
    b = SceneBuilder("cornell_box")
    white = b.add_material((0.73, 0.73, 0.73), roughness=0.9)
    red = b.add_material((0.65, 0.05, 0.05), roughness=0.9)
    green = b.add_material((0.12, 0.45, 0.15), roughness=0.9)
    metal = b.add_material((0.9, 0.85, 0.6), roughness=0.25, metallic=1.0)
    glossy = b.add_material((0.2, 0.3, 0.8), roughness=0.35)

    def quad(o, du, dv):
        o, du, dv = (np.asarray(a, np.float64) for a in (o, du, dv))
        return grid_patch(tess, tess, lambda u, v: o + u[..., None] * du + v[..., None] * dv)
    # room x,z in [-1,1], y in [0,2], open towards +z; normals point inwards
    b.add_drawable(b.add_mesh(quad((-1, 0, 1), (2, 0, 0), (0, 0, -2))), white)     # floor   (+y)
    b.add_drawable(b.add_mesh(quad((-1, 2, -1), (2, 0, 0), (0, 0, 2))), white)     # ceiling (-y)
    b.add_drawable(b.add_mesh(quad((-1, 0, -1), (2, 0, 0), (0, 2, 0))), white)     # back    (+z)
    b.add_drawable(b.add_mesh(quad((-1, 0, 1), (0, 0, -2), (0, 2, 0))), red)       # left    (+x)
    b.add_drawable(b.add_mesh(quad((1, 0, -1), (0, 0, 2), (0, 2, 0))), green)      # right   (-x)
    box = b.add_mesh(box_mesh((-0.5, 0.0, -0.5), (0.5, 1.0, 0.5), n=max(2, tess // 4)))
    b.add_drawable(box, metal, translate(-0.35, 0, -0.3) @ rotate_y(0.3) @ scale(0.6, 1.2, 0.6))
    b.add_drawable(box, white, translate(0.4, 0, 0.3) @ rotate_y(-0.3) @ scale(0.6, 0.6, 0.6))
    ball = b.add_mesh(sphere_mesh(0.25, 4 * tess // 2, 2 * tess // 2))
    b.add_drawable(ball, glossy, translate(0.4, 0.6 + 0.25, 0.3))
    cam = dict(position=(0.0, 1.0, 4.6), front_dir=(0.0, 0.0, -1.0), up_dir=(0.0, 1.0, 0.0), yfov=30.0, near_z=0.001, far_z=1e5)
    return b.finish(dir_lights=dir_light((0.25, 0.55, 1.0), (1.0, 0.9, 0.8), 4.0), camera=cam,
                    bounds=(np.array([-1, 0, -1.0]), np.array([1, 2, 1.0])))


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[1]: procedural Sponza-scale atrium, ~262k triangles, 25 PBR materials,
# directional light + skybox (README gallery set-up). Seed fixed by SURVEY §8d.
# ---------------------------------------------------------------------------------------------
ATRIUM_SEED = 0x5B0A2A1D
ATRIUM_TRIANGLES = 262144


def atrium(seed: int = ATRIUM_SEED, target_triangles: int = ATRIUM_TRIANGLES, sky_size: int = 256) -> SceneData:
    rng = np.random.default_rng(seed)
    b = SceneBuilder("atrium")
    mats = []
    for _ in range(25):   # base U[0.05,0.9]^3, roughness U[0.1,1], metallic in {0,1} p=0.2
        mats.append(b.add_material(tuple(rng.uniform(0.05, 0.9, 3)), roughness=float(rng.uniform(0.1, 1.0)),
                                   metallic=1.0 if rng.uniform() < 0.2 else 0.0))
    nxt = iter(range(10 ** 9))

    def mat():
        return mats[next(nxt) % 25]
    LX, LY, LZ = 18.0, 12.0, 7.0     # half length, height, half width of the hall

    def quad(o, du, dv, nu, nv, bump=0.0, freq=6.0):
        o, du, dv = (np.asarray(a, np.float64) for a in (o, du, dv))
        n = np.cross(du, dv)
        n = n / np.linalg.norm(n)
        ph = rng.uniform(0, 6.28, 2)

        def fn(u, v):
            p = o + u[..., None] * du + v[..., None] * dv
            if bump:
                p = p + n * (bump * np.sin(freq * 2 * np.pi * u + ph[0]) * np.sin(freq * 2 * np.pi * v + ph[1]))[..., None]
            return p
        return grid_patch(nu, nv, fn)
    # columns: two rows of 8, shared mesh (instancing), 64 x 32 quads each
    col = b.add_mesh(cylinder_mesh(0.45, 5.0, 48, 24, bulge=0.12))
    col_up = b.add_mesh(cylinder_mesh(0.32, 4.0, 32, 16, bulge=0.08))
    xs = np.linspace(-LX + 2.5, LX - 2.5, 8)
    for x in xs:
        for z in (-3.6, 3.6):
            b.add_drawable(col, mat(), translate(x, 0.0, z))
            b.add_drawable(col_up, mat(), translate(x, 5.6, z))
    # arches between neighbouring columns (half torus), one mesh instanced 14 times
    def arch_fn(u, v):
        a = np.pi * u
        R, r = (xs[1] - xs[0]) * 0.5, 0.28
        ang = 2 * np.pi * v
        cx = -R * np.cos(a)
        cy = R * np.sin(a) * 0.6
        rr = r * (1 + 0.15 * np.cos(6 * a))
        return np.stack([cx + rr * np.cos(ang) * (-np.cos(a)), cy + rr * np.cos(ang) * np.sin(a) * 0.6, rr * np.sin(ang)], -1)
    arch = b.add_mesh(grid_patch(32, 12, arch_fn))
    for i in range(7):
        for z in (-3.6, 3.6):
            b.add_drawable(arch, mat(), translate((xs[i] + xs[i + 1]) * 0.5, 5.0, z))
    # gallery slabs on both sides (boxes) and balustrade spheres
    slab = b.add_mesh(box_mesh((-LX, 5.25, -0.5), (LX, 5.6, 0.5), n=16))
    b.add_drawable(slab, mat(), translate(0, 0, -5.2) @ scale(1, 1, 3.4))
    b.add_drawable(slab, mat(), translate(0, 0, 5.2) @ scale(1, 1, 3.4))
    orb = b.add_mesh(sphere_mesh(0.35, 32, 16, bumps=0.05, seed=seed & 0xffff))
    for x in xs:
        for z in (-3.6, 3.6):
            b.add_drawable(orb, mat(), translate(x, 9.95, z) @ random_rotation(rng))
    # drapes: wavy two-sided cloth hanging between the upper columns
    drape_two_sided = [b.add_material(tuple(rng.uniform(0.1, 0.9, 3)), roughness=float(rng.uniform(0.5, 1.0)), two_sided=True) for _ in range(3)]
    for i in range(6):
        x0 = xs[i] + 0.6
        w = (xs[1] - xs[0]) - 1.2
        z = -3.6 if i % 2 == 0 else 3.6
        ph = rng.uniform(0, 6.28, 3)

        def drape_fn(u, v, x0=x0, w=w, z=z, ph=ph):
            sag = 0.35 * np.sin(np.pi * u)
            fold = 0.18 * np.sin(10 * np.pi * u + ph[0]) * (0.3 + v) + 0.05 * np.sin(23 * u + 9 * v + ph[1])
            return np.stack([x0 + w * u, 9.4 - sag * (1 - v) - 3.4 * v, z + fold], -1)
        b.add_drawable(b.add_mesh(grid_patch(64, 48, drape_fn)), drape_two_sided[i % 3])
    # vases / planters on the floor: displaced spheres, per-instance rotation + scale
    vase = b.add_mesh(sphere_mesh(0.6, 64, 32, bumps=0.12, seed=(seed >> 8) & 0xffff))
    for i in range(10):
        x = rng.uniform(-LX + 2, LX - 2)
        z = rng.choice([-1.0, 1.0]) * rng.uniform(0.5, 2.6)
        s = rng.uniform(0.6, 1.3)
        b.add_drawable(vase, mat(), translate(x, 0.6 * s, z) @ rotate_y(rng.uniform(0, 6.28)) @ scale(s))
    # shell: walls, end walls, ceiling ring with a central opening (the sky + sun come through it)
    b.add_drawable(b.add_mesh(quad((-LX, 0, -LZ), (0, LY, 0), (2 * LX, 0, 0), 32, 96, bump=0.03)), mat())   # -z wall (+z normal)
    b.add_drawable(b.add_mesh(quad((-LX, 0, LZ), (2 * LX, 0, 0), (0, LY, 0), 96, 32, bump=0.03)), mat())    # +z wall (-z normal)
    b.add_drawable(b.add_mesh(quad((-LX, 0, -LZ), (0, 0, 2 * LZ), (0, LY, 0), 32, 32, bump=0.02)), mat())    # -x end  (+x normal)
    b.add_drawable(b.add_mesh(quad((LX, 0, -LZ), (0, LY, 0), (0, 0, 2 * LZ), 32, 32, bump=0.02)), mat())     # +x end  (-x normal)
    ox, oz = 11.0, 2.6    # half extents of the roof opening
    b.add_drawable(b.add_mesh(quad((-LX, LY, -LZ), (2 * LX, 0, 0), (0, 0, LZ - oz), 64, 12)), mat())         # ceiling strips (-y normal)
    b.add_drawable(b.add_mesh(quad((-LX, LY, oz), (2 * LX, 0, 0), (0, 0, LZ - oz), 64, 12)), mat())
    b.add_drawable(b.add_mesh(quad((-LX, LY, -oz), (LX - ox, 0, 0), (0, 0, 2 * oz), 16, 12)), mat())
    b.add_drawable(b.add_mesh(quad((ox, LY, -oz), (LX - ox, 0, 0), (0, 0, 2 * oz), 16, 12)), mat())
    # floor last: its tessellation absorbs the remainder so the scene has exactly `target_triangles`
    used = sum(b.meshes[m][2] for (m, _, _) in b.drawables)
    remaining = target_triangles - used
    if remaining < 2 * 64:
        raise ValueError("target_triangles too small for the atrium layout")
    nu = 256
    nv = remaining // (2 * nu)
    b.add_drawable(b.add_mesh(quad((-LX, 0, LZ), (2 * LX, 0, 0), (0, 0, -2 * LZ), nu, nv, bump=0.015, freq=20.0)), mat())
    rest = remaining - 2 * nu * nv        # < 512 triangles: a plinth strip under the -x end wall
    if rest >= 2:
        b.add_drawable(b.add_mesh(quad((-LX + 0.01, 0.0, -LZ), (0, 0, 2 * LZ), (0, 0.4, 0), rest // 2, 1)), mat())
    sun = (0.3, 1.0, 0.2)
    cam = dict(position=(-LX + 1.5, 3.2, 0.6), front_dir=(1.0, 0.12, -0.03), up_dir=(0.0, 1.0, 0.0), yfov=30.0, near_z=0.001, far_z=1e5)
    return b.finish(dir_lights=dir_light(sun, (1.0, 0.9, 0.8), 4.0), sky_faces=procedural_sky(sky_size, sun) if sky_size else None,
                    camera=cam, bounds=(np.array([-LX, 0, -LZ]), np.array([LX, LY, LZ])))


def small_test_scene(seed=7, with_translucent=True) -> SceneData:
    """A few hundred triangles exercising: instancing with rotation + non-uniform scale, two-sided,
    alpha-test and stochastic-opacity (any-hit) drawables, a metallic and a rough material."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder("small_test")
    ground = b.add_material((0.6, 0.6, 0.55), roughness=0.8)
    shiny = b.add_material((0.9, 0.6, 0.3), roughness=0.2, metallic=1.0)
    rough = b.add_material((0.2, 0.5, 0.8), roughness=0.7)
    two = b.add_material((0.7, 0.2, 0.6), roughness=0.6, two_sided=True)
    cut = b.add_material((0.9, 0.9, 0.2, 0.005), roughness=0.5, blend=capi.BLEND_ALPHA_TEST, two_sided=True)
    glass = b.add_material((0.3, 0.9, 0.4, 0.45), roughness=0.4, blend=capi.BLEND_TRANSLUCENT, two_sided=True)
    b.add_drawable(b.add_mesh(grid_patch(12, 12, lambda u, v: np.stack([8 * u - 4, 0.15 * np.sin(5 * u) * np.cos(4 * v), 4 - 8 * v], -1))), ground)
    ball = b.add_mesh(sphere_mesh(0.5, 16, 8, bumps=0.1, seed=3))
    box = b.add_mesh(box_mesh((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), n=2))
    for i in range(6):
        m = [shiny, rough, two][i % 3]
        xf = translate(rng.uniform(-2.5, 2.5), rng.uniform(0.6, 1.6), rng.uniform(-2.5, 2.5)) @ random_rotation(rng) @ scale(*rng.uniform(0.5, 1.4, 3))
        b.add_drawable(ball if i % 2 == 0 else box, m, xf)
    sheet = b.add_mesh(grid_patch(4, 4, lambda u, v: np.stack([3 * u - 1.5, 0.3 + 2.2 * v, 0 * u + 1.2], -1)))
    if with_translucent:
        b.add_drawable(sheet, glass)
        b.add_drawable(sheet, cut, translate(0, 0, 0.6))
    cam = dict(position=(0.3, 2.2, 6.5), front_dir=(-0.03, -0.22, -1.0), up_dir=(0.0, 1.0, 0.0), yfov=40.0, near_z=0.001, far_z=1e5)
    return b.finish(dir_lights=dir_light((0.4, 1.0, 0.6), (1.0, 0.95, 0.9), 3.0), sky_faces=procedural_sky(16, (0.4, 1.0, 0.6)),
                    camera=cam, bounds=(np.array([-4, 0, -4.0]), np.array([4, 3, 4.0])))


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: mixed lights — point / spot / LTC rect lights packed exactly as
# LightsContext::collect_all_lights does (bisemutum/src/renderer/context/lights.cpp:125-229).
# ---------------------------------------------------------------------------------------------
def _rotation_rows(rng):
    return random_rotation(rng)[:3, :3]


def add_mixed_lights(scene: SceneData, n_point: int, n_rect: int, ltc_luts, seed: int = 3, strength: float = 6.0,
                     light_range: float = 30.0, keep_dir_lights: bool = False) -> SceneData:
    """n_point point lights (every second one a spot, inner 30 / outer 60 degrees) at uniform positions
    inside the scene bounds + n_rect one-sided, untextured 1x1 m rect lights."""
    rng = np.random.default_rng(seed)
    lo, hi = (np.asarray(b, np.float64) for b in scene.bounds)
    pad = 0.08 * (hi - lo)
    pl = np.zeros(n_point, capi.POINT_LIGHT)
    for i in range(n_point):
        R = _rotation_rows(rng)
        color = rng.uniform(0.3, 1.0, 3)
        pl[i]["emission"] = (color * strength).astype(f32)
        pl[i]["position"] = rng.uniform(lo + pad, hi - pad).astype(f32)
        pl[i]["direction"] = (R @ np.array([0.0, 1.0, 0.0])).astype(f32)          # lights point along local +Y
        if i % 2 == 1:
            pl[i]["cos_outer"] = np.cos(np.radians(f32(60.0)))
            pl[i]["cos_inner"] = np.cos(np.radians(f32(30.0)))
        pl[i]["range_sqr_inv"] = f32(1.0) / (f32(light_range) * f32(light_range))
        pl[i]["sm_index"] = -1
    rl = np.zeros(n_rect, capi.RECT_LIGHT)
    for i in range(n_rect):
        R = _rotation_rows(rng)
        c = rng.uniform(lo + pad, hi - pad)
        w = h = 0.5
        rl[i]["emission"] = (rng.uniform(0.3, 1.0, 3) * strength).astype(f32)
        rl[i]["texture_index"] = -1
        rl[i]["center_position"] = c.astype(f32)
        rl[i]["two_sided"] = 0
        rl[i]["position0"] = (c + R @ np.array([w, h, 0.0])).astype(f32)
        rl[i]["position1"] = (c + R @ np.array([-w, h, 0.0])).astype(f32)
        rl[i]["position2"] = (c + R @ np.array([-w, -h, 0.0])).astype(f32)
        rl[i]["position3"] = (c + R @ np.array([w, -h, 0.0])).astype(f32)
        rl[i]["normal"] = (R @ np.array([0.0, 0.0, 1.0])).astype(f32)
        rl[i]["inv_width_sqr"] = 1.0
        rl[i]["inv_height_sqr"] = 1.0
    scene.point_lights, scene.rect_lights = pl, rl
    if not keep_dir_lights:
        scene.dir_lights = np.zeros(0, capi.DIR_LIGHT)
    scene.ltc_luts = tuple(np.ascontiguousarray(a, f32) for a in ltc_luts) if n_rect else None
    scene.name += f"+{n_point}point+{n_rect}rect"
    return scene


def light_texture(width: int, height: int, fmt: int = None, levels: int = 16, seed: int = 5, mip_linear: int = 1, linear: int = 1) -> dict:
    """A procedural rect-light texture (colour bars over a soft gradient, a dark frame): enough structure that a wrong (u, v), mip level or
    filter shows up. 8-bit formats as uint8 (H, W, 4), RGBA32F as float32."""
    fmt = capi.TEXTURE_RGBA8_SRGB if fmt is None else fmt
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width]
    u, v = (x + 0.5) / width, (y + 0.5) / height
    bars = rng.uniform(0.1, 1.0, (7, 3))[(u * 7).astype(int) % 7]
    img = bars * (0.35 + 0.65 * v[..., None]) * (0.6 + 0.4 * np.sin(9.0 * u + 4.0 * v)[..., None] ** 2)
    frame = (np.minimum(np.minimum(u, 1 - u), np.minimum(v, 1 - v)) < 0.06)[..., None]
    img = np.where(frame, 0.03, img)
    rgba = np.concatenate([img, np.ones((height, width, 1))], -1)
    texels = rgba.astype(f32) * 2.0 if fmt == capi.TEXTURE_RGBA32_FLOAT else np.clip(np.rint(rgba * 255.0), 0, 255).astype(np.uint8)
    return {"texels": np.ascontiguousarray(texels), "format": fmt, "levels": levels, "address_u": capi.ADDRESS_CLAMP, "address_v": capi.ADDRESS_CLAMP,
            "linear": linear, "mip_linear": mip_linear}


def texture_rect_lights(scene: SceneData, textures: list, light_width: float = 1.0, light_height: float = 1.0) -> SceneData:
    """Gives rect light i the texture i % len(textures), with inv_texel_size as LightsContext::collect_all_lights computes it
    (lights.cpp:233-237: max(texture width / light width, texture height / light height))."""
    scene.light_textures = list(textures)
    for i in range(len(scene.rect_lights)):
        k = i % len(textures)
        t = textures[k]["texels"]
        scene.rect_lights[i]["texture_index"] = k
        scene.rect_lights[i]["inv_texel_size"] = max(f32(t.shape[1]) / f32(light_width), f32(t.shape[0]) / f32(light_height))
    scene.name += "+lighttex"
    return scene


def load_ltc_luts(npz_path: str):
    z = np.load(npz_path)
    return tuple(np.ascontiguousarray(z[k], f32) for k in ("matrix_lut0", "matrix_lut1", "matrix_lut2", "norm_lut"))


def mixed_lights(ltc_luts, seed: int = ATRIUM_SEED) -> SceneData:
    """configs[2]: the atrium with 64 point/spot lights + 16 LTC rect lights."""
    return add_mixed_lights(atrium(seed), 64, 16, ltc_luts, seed=seed & 0xffff)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: one large mesh x many instances (two-level BVH). Full size: a displaced
# sphere of 2048 x 512 quads = 2 097 152 triangles, 512 instances on a jittered 8x8x8 grid with
# random rotations. Smaller parameters give the parity-test version.
# ---------------------------------------------------------------------------------------------
INSTANCED_SEED = 0x1257A9CE


def instanced(mesh_seg: int = 2048, mesh_ring: int = 512, grid: int = 8, seed: int = INSTANCED_SEED, sky_size: int = 64) -> SceneData:
    rng = np.random.default_rng(seed)
    b = SceneBuilder(f"instanced_{mesh_seg * mesh_ring * 2}x{grid ** 3}")
    mats = [b.add_material(tuple(rng.uniform(0.1, 0.9, 3)), roughness=float(rng.uniform(0.15, 1.0)),
                           metallic=1.0 if rng.uniform() < 0.25 else 0.0) for _ in range(16)]
    mesh = b.add_mesh(sphere_mesh(1.0, mesh_seg, mesh_ring, bumps=0.18, seed=seed & 0xffff))
    spacing = 3.4
    half = (grid - 1) * spacing * 0.5
    k = 0
    for ix in range(grid):
        for iy in range(grid):
            for iz in range(grid):
                c = np.array([ix, iy, iz]) * spacing - half + rng.uniform(-0.6, 0.6, 3)
                s = rng.uniform(0.7, 1.25)
                b.add_drawable(mesh, mats[k % len(mats)], translate(*c) @ random_rotation(rng) @ scale(s))
                k += 1
    ext = half + 2.5
    ground = b.add_mesh(grid_patch(8, 8, lambda u, v: np.stack([(2 * u - 1) * ext * 2, 0 * u - ext, (1 - 2 * v) * ext * 2], -1)))
    b.add_drawable(ground, b.add_material((0.55, 0.55, 0.5), roughness=0.85))
    sun = (0.35, 1.0, 0.45)
    cam = dict(position=(0.2 * ext, 0.35 * ext, 3.1 * ext), front_dir=(-0.05, -0.1, -1.0), up_dir=(0.0, 1.0, 0.0), yfov=30.0, near_z=0.001, far_z=1e5)
    return b.finish(dir_lights=dir_light(sun, (1.0, 0.95, 0.9), 3.5), sky_faces=procedural_sky(sky_size, sun) if sky_size else None,
                    camera=cam, bounds=(np.full(3, -ext), np.full(3, ext)))


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: DDGI-style probe grid over a scene's bounds.
# ---------------------------------------------------------------------------------------------
def ddgi_sample_randoms() -> np.ndarray:
    """The 8192-entry R2 sequence of DdgiContext::init_sample_randoms (src/renderer/context/ddgi.cpp:125-145),
    float32 accumulation exactly as the reference does it."""
    phi2 = 1.0 / 1.3247179572447
    delta = np.array([phi2, phi2 * phi2], dtype=f32)
    out = np.zeros((8192, 2), f32)
    out[0] = (0.5, 0.5)
    for i in range(1, 8192):
        v = out[i - 1] + delta
        v = np.where(v >= f32(1.0), v - f32(1.0), v).astype(f32)
        out[i] = v
    return out


def probe_volume(scene: SceneData, counts=(32, 32, 16), rays_per_probe: int = 256, ray_length: float = 1024.0, margin: float = 0.04):
    """Axis-aligned probe volume covering the scene bounds (slightly inset); counts follow the axes (x, y, z)."""
    lo, hi = (np.asarray(b, np.float64) for b in scene.bounds)
    pad = margin * (hi - lo)
    v = capi.ProbeVolume()
    v.base_position[:] = list((lo + pad).astype(f32))
    v.frame_x[:] = [1, 0, 0]; v.frame_y[:] = [0, 1, 0]; v.frame_z[:] = [0, 0, 1]
    v.extent[:] = list((hi - lo - 2 * pad).astype(f32))
    v.ray_length = ray_length
    v.probe_counts[:] = list(counts)
    v.rays_per_probe = rays_per_probe
    return v


# ---------------------------------------------------------------------------------------------
# The reference's own example project (examples/scene_basic): its three .biasset meshes, its five
# materials and its scene.toml placement. The asset bytes travel as tests/golden/scene_basic.npz
# (made by tests/golden/make_scene_basic_fixture.py); transforms follow Transform (TRS, rotation =
# glm::eulerAngleZXY(z, x, y) in degrees → radians, src/math/transform.cpp:98-123).
# ---------------------------------------------------------------------------------------------
def _euler_zxy(deg):
    x, y, z = np.radians(np.asarray(deg, np.float64))
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    return Rz @ Rx @ Ry


def _trs(translation, rotation_deg, scaling):
    m = np.eye(4)
    m[:3, :3] = _euler_zxy(rotation_deg) @ np.diag(scaling)
    m[:3, 3] = translation
    return m


def scene_basic(npz_path: str) -> SceneData:
    z = np.load(npz_path)
    b = SceneBuilder("scene_basic")

    def mesh(name):
        idx = z[f"{name}.indices"].reshape(-1, 3).astype(u32)
        return b.add_mesh((z[f"{name}.positions"], z[f"{name}.normals"], z[f"{name}.tangents"], z[f"{name}.texcoords"], idx))
    plane, cube, sphere = mesh("plane"), mesh("cube"), mesh("sphere")
    t_earth = b.add_texture(z["earth_diffuse_srgb"], fmt=capi.TEXTURE_RGBA8_SRGB)
    t_normal = b.add_texture(z["earth_normal"], fmt=capi.TEXTURE_RGBA8_UNORM)
    t_cage = b.add_texture(z["cage"], fmt=capi.TEXTURE_RGBA8_UNORM)
    K = capi
    checker = b.add_material((0.8, 0.8, 0.8, 1.0), roughness=0.1, kind=K.MATERIAL_KIND_CHECKERBOARD)     # base_color.a = roughness_0 = 1.0
    b.materials[checker]["emission"] = (0.1, 0.1, 0.1)                                                   # base_color_1; roughness_1 = 0.1
    textured = b.add_material(roughness=0.2, kind=K.MATERIAL_KIND_TEXTURED, base_color_tex=t_earth, normal_map_tex=t_normal)
    white = b.add_material((1.0, 1.0, 1.0), kind=K.MATERIAL_KIND_CONSTANT_COLOR)
    transparent = b.add_material((1.0, 0.0, 0.5, 0.5), kind=K.MATERIAL_KIND_TRANSPARENT, blend=K.BLEND_TRANSLUCENT, two_sided=True)
    cage = b.add_material(kind=K.MATERIAL_KIND_CAGE, blend=K.BLEND_ALPHA_TEST, two_sided=True, base_color_tex=t_cage)
    # scene.toml objects, in file order (drawable index = order of StaticMeshRenderSystem's view)
    b.add_drawable(sphere, textured, _trs((-1.0, 0.0, 1.0), (0, 0, 0), (1, 1, 1)))
    b.add_drawable(cube, white, _trs((1.0, 0.0, -1.0), (0.0, 30.000001907348633, 0.0), (1, 1, 1)))
    b.add_drawable(cube, transparent, _trs((2.5, 0.01, 2.5), (0.0, 30.000001907348633, 0.0), (1, 1, 1)))
    b.add_drawable(cube, cage, _trs((1.0, 2.5, -1.0), (0.0, 30.000001907348633, 0.0), (1, 1, 1)))
    b.add_drawable(plane, checker, _trs((0.0, -1.0, 0.0), (0, 0, 0), (5.0, 1.0, 5.0)))
    # Dir Light: rotation (10, 0, -30) deg; lights point along local +Y (lights.cpp:61); colour (1, .9, .8) x strength 4
    light_dir = _euler_zxy((10.000005722045898, 0.0, -30.000001907348633)) @ np.array([0.0, 1.0, 0.0])
    # Camera: looks down local -Z, up +Y (camera_system.cpp:17-29); yfov 30, near 0.01, far 1e4, 1298 x 635 target
    R = _euler_zxy((-9.3324127197265625, 63.346549987792969, 17.90446662902832))
    cam = dict(position=(10.5, 4.5, 6.0), front_dir=tuple(R @ np.array([0.0, 0.0, -1.0])), up_dir=tuple(R @ np.array([0.0, 1.0, 0.0])),
               yfov=30.0, near_z=0.010000000707805157, far_z=10000.0)
    return b.finish(dir_lights=dir_light(light_dir, (1.0, 0.89999997615814209, 0.80000001192092896), 4.0), camera=cam,
                    sky_faces=procedural_sky(32, light_dir), bounds=(np.array([-5.0, -1.0, -5.0]), np.array([5.0, 3.5, 5.0])))
