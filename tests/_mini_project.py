"""Test helper: writes a small bisemutum project directory (project.toml, asset_metadata.toml, scene.toml, two materials, two .biasset
meshes, one .biasset texture) in the reference's on-disk formats, so that the C++ project loader can be exercised
without the reference checkout (e.g. on the GPU box)."""
import os
import struct
import zlib

import numpy as np


def _header(type_name: str, version: int) -> bytes:
    return struct.pack("<I", 0x0B1A55E7) + struct.pack("<Q", len(type_name)) + type_name.encode() + struct.pack("<I", version)


def _vec(a: np.ndarray, items: int) -> bytes:
    return struct.pack("<Q", items) + np.ascontiguousarray(a).tobytes()


def _compressed(raw: bytes) -> bytes:
    c = zlib.compress(raw)
    return struct.pack("<QQ", len(raw), len(c)) + c


def write_mesh(path, positions, normals, tangents, texcoords, indices, version=2):
    nv = len(positions)
    sub = struct.pack("<IIIB3x", 0, 0, 0xFFFFFFFF, 0)                      # one submesh: whole mesh ("to the end")
    body = (_vec(np.float32(positions), nv) + _vec(np.float32(normals), nv) + _vec(np.float32(tangents), nv) + _vec(np.zeros((0, 3), np.float32), 0) +
            _vec(np.float32(texcoords), nv) + _vec(np.zeros((0, 2), np.float32), 0) + _vec(np.uint32(indices), len(indices)) + struct.pack("<Q", 1) + sub)
    with open(path, "wb") as f:
        f.write(_header("StaticMesh", version) + (_compressed(body) if version >= 2 else body))


def write_texture(path, texels_rgba8, fmt=37, storage="v2"):
    """storage: "v2" (zlib part), "v1_raw" (storage type 0) or "v1_png" (storage type 1: one PNG per layer, texture.cpp:110-131)."""
    h, w, _ = texels_rgba8.shape
    sampler = bytes([1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]) + struct.pack("<ffff", 0.0, 0.0, 0.0, 1000.0)      # linear / linear / repeat
    desc = struct.pack("<IIII", w, h, 1, 1) + bytes([fmt, 1, 1, 0])
    raw = struct.pack("<Q", texels_rgba8.size) + np.ascontiguousarray(texels_rgba8, np.uint8).tobytes()
    if storage == "v1_raw":
        body, version = struct.pack("<I", 0) + raw, 1
    elif storage == "v1_png":
        import _gltf_writer
        png = _gltf_writer.png_bytes(texels_rgba8)
        body, version = struct.pack("<I", 1) + struct.pack("<Q", len(png)) + png, 1
    else:
        body, version = _compressed(raw), 2
    with open(path, "wb") as f:
        f.write(_header("Texture", version) + sampler + desc + body)


def quad(size=4.0):
    p = np.float32([[-size, 0, -size], [size, 0, -size], [size, 0, size], [-size, 0, size]])
    n = np.float32([[0, 1, 0]] * 4); t = np.float32([[1, 0, 0, 1]] * 4); uv = np.float32([[0, 0], [1, 0], [1, 1], [0, 1]])
    return p, n, t, uv, np.uint32([0, 2, 1, 0, 3, 2])


def box(h=0.5):
    P, N, T, UV, I = [], [], [], [], []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            n = np.zeros(3); n[axis] = sgn
            u = np.zeros(3); u[(axis + 1) % 3] = 1.0
            v = np.cross(n, u)
            base = len(P)
            for a, b in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                P.append(n * h + u * a * h + v * b * h); N.append(n); T.append([*u, 1.0]); UV.append([(a + 1) / 2, (b + 1) / 2])
            I += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.float32(P), np.float32(N), np.float32(T), np.float32(UV), np.uint32(I)


def write(directory: str) -> dict:
    os.makedirs(os.path.join(directory, "meshes"), exist_ok=True); os.makedirs(os.path.join(directory, "materials"), exist_ok=True)
    os.makedirs(os.path.join(directory, "textures"), exist_ok=True)
    q, b = quad(), box()
    write_mesh(os.path.join(directory, "meshes", "plane.static_mesh.biasset"), *q, version=2)
    write_mesh(os.path.join(directory, "meshes", "cube.static_mesh.biasset"), *b, version=1)
    rng = np.random.default_rng(4)
    tex = rng.integers(0, 256, (16, 32, 4), dtype=np.uint8); tex[..., 3] = np.where(rng.uniform(size=(16, 32)) < 0.4, 0, 255)
    write_texture(os.path.join(directory, "textures", "cage.texture.biasset"), tex, 37)
    open(os.path.join(directory, "project.toml"), "w").write('name = "mini"\nasset_metadata_file = "/project/asset_metadata.toml"\nscene_file = "/project/scene.toml"\nrenderer = "BasicRenderer"\n')
    assets = [("meshes/plane.static_mesh.biasset", "StaticMesh"), ("meshes/cube.static_mesh.biasset", "StaticMesh"), ("materials/checker.toml", "Material"),
              ("materials/cage.toml", "Material"), ("textures/cage.texture.biasset", "Texture"), ("materials/white.toml", "Material")]
    open(os.path.join(directory, "asset_metadata.toml"), "w").write("".join(f"[[assets]]\nid = {i}\npath = '/project/{p}'\ntype = '{t}'\n\n" for i, (p, t) in enumerate(assets)))
    open(os.path.join(directory, "materials", "checker.toml"), "w").write("""blend_mode = 'opaque'
material_function = '''int grid = int(floor(vertex.position_world.x))
    ^ int(floor(vertex.position_world.z));
surface.base_color = (grid & 1) == 1 ? PARAM_base_color_0 : PARAM_base_color_1;
surface.roughness = (grid & 1) == 1 ? PARAM_roughness_0 : PARAM_roughness_1;
'''
surface_model = 'lit'

[[params]]
name = 'base_color_0'
value = [ 0.8, 0.7, 0.6 ]

[[params]]
name = 'base_color_1'
value = [ 0.1, 0.2, 0.3 ]

[[params]]
name = 'roughness_0'
value = 0.9

[[params]]
name = 'roughness_1'
value = 0.25
""")
    open(os.path.join(directory, "materials", "cage.toml"), "w").write("""blend_mode = 'alpha_test'
material_function = '''float4 value = PARAM_cage_tex.Sample(PARAM_cage_tex_sampler, vertex.texcoord);
surface.base_color = value.xyz;
surface.f0_color = value.xyz;
surface.opacity = value.w < 0.5 ? 0.0 : 1.0;
surface.two_sided = true;
'''
surface_model = 'lit'

[[params]]
name = 'cage_tex'

    [params.value]
    asset_id = 4
""")
    open(os.path.join(directory, "materials", "white.toml"), "w").write("blend_mode = 'opaque'\nmaterial_function = '''surface.base_color = PARAM_base_color;\n'''\nsurface_model = 'lit'\n\n[[params]]\nname = 'base_color'\nvalue = [ 0.9, 0.9, 0.9 ]\n")

    def obj(name, comps):
        return f"[[objects]]\nname = '{name}'\n\n" + "".join(comps) + "\n"

    def xf(t, r=(0, 0, 0), s=(1, 1, 1)):
        return f"    [[objects.components]]\n    type = 'Transform'\n\n        [objects.components.value]\n        rotation = [ {r[0]}, {r[1]}, {r[2]} ]\n        scaling = [ {s[0]}, {s[1]}, {s[2]} ]\n        translation = [ {t[0]}, {t[1]}, {t[2]} ]\n\n"

    def mesh(mesh_id, mat_id):
        return (f"    [[objects.components]]\n    type = 'StaticMeshComponent'\n\n        [objects.components.value.static_mesh]\n        asset_id = {mesh_id}\n\n"
                f"    [[objects.components]]\n    type = 'MeshRendererComponent'\n\n        [objects.components.value]\n        submesh_start_index = 0\n\n"
                f"            [[objects.components.value.materials]]\n            asset_id = {mat_id}\n\n")
    scene = (obj("Camera", [xf((3.0, 2.5, 5.0), (-20.0, 30.0, 0.0)), "    [[objects.components]]\n    type = 'CameraComponent'\n\n        [objects.components.value]\n        far_z = 1000.0\n        near_z = 0.01\n        projection_type = 'perspective'\n        render_target_size = [ 96, 64 ]\n        yfov = 40.0\n\n"]) +
             obj("Sun", [xf((0, 0, 0), (25.0, 0.0, -35.0)), "    [[objects.components]]\n    type = 'DirectionalLightComponent'\n\n        [objects.components.value]\n        cast_shadow = true\n        color = [ 1.0, 0.9, 0.8 ]\n        strength = 3.0\n\n"]) +
             obj("Lamp", [xf((1.5, 2.0, 1.0)), "    [[objects.components]]\n    type = 'PointLightComponent'\n\n        [objects.components.value]\n        color = [ 0.4, 0.6, 1.0 ]\n        range = 12.0\n        spot = false\n        strength = 5.0\n\n"]) +
             obj("Floor", [xf((0, 0, 0)), mesh(0, 2)]) +
             obj("Crate", [xf((0.5, 0.5, 0.0), (0.0, 25.0, 0.0)), mesh(1, 5)]) +
             obj("Cage", [xf((-1.2, 0.75, 0.8), (0.0, -15.0, 0.0), (1.5, 1.5, 1.5)), mesh(1, 3)]) +
             obj("Settings", [xf((0, 0, 0)), "    [[objects.components]]\n    type = 'BasicRendererOverrideVolume'\n\n        [objects.components.value]\n        priority = 0.0\n\n            [objects.components.value.settings]\n            pipeline_mode = 'path_tracing'\n\n                [objects.components.value.settings.path_tracing]\n                accumulate = true\n                denoise = false\n                max_bounces = 4\n                ray_length = 50.0\n\n"]))
    open(os.path.join(directory, "scene.toml"), "w").write(scene)
    return {"quad": q, "box": b, "texture": tex}
