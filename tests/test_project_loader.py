"""Headless ingestion: the C++ host library loads the reference's OWN example project
(/root/reference/examples/scene_basic: project.toml, asset_metadata.toml, scene.toml, materials/*.toml, *.biasset meshes and
textures) and produces the arrays the C ABI takes. Checked against (a) the independent Python readers of
tests/golden/make_scene_basic_fixture.py, (b) the hand-assembled scenes.scene_basic() that the golden / parity tests use, and
(c) the oracle: the loaded scene renders. /root/reference exists only in the build container: skipped elsewhere."""
import os
import sys

import numpy as np
import pytest

from bisemutum_engine_b200 import capi, engine, scenes

REF = "/root/reference/examples/scene_basic"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not present on this machine")


@pytest.fixture(scope="module")
def project():
    p = engine.Project(REF)
    yield p
    p.close()


@needs_ref
def test_biasset_and_toml_match_the_python_readers(project):
    sys.path.insert(0, GOLDEN)
    import make_scene_basic_fixture as fx
    info = project.info
    assert (info.num_drawables, info.num_blas, info.num_materials, info.num_textures, info.num_dir_lights) == (5, 3, 5, 3, 1)
    assert (info.target_width, info.target_height) == (1298, 635) and info.max_bounces == 3 and info.ray_length == 100.0 and info.accumulate == 1
    assert info.ambient_occlusion.range == 0.5 and info.ambient_occlusion.strength == 0.5 and info.ambient_occlusion.half_resolution == 1
    # meshes in order of first use by scene.toml: sphere (asset 2), cube (1), plane (0)
    meshes = [fx.read_mesh(os.path.join(REF, "meshes", f"{n}.static_mesh.biasset")) for n in ("sphere", "cube", "plane")]
    for key, comps in (("positions", 3), ("normals", 3), ("tangents", 4), ("texcoords", 2)):
        np.testing.assert_array_equal(project.array(key), np.concatenate([m[key].reshape(-1) for m in meshes]))
    np.testing.assert_array_equal(project.array("indices"), np.concatenate([m["indices"] for m in meshes]))
    blas = project.array("blas")
    vb = np.cumsum([0] + [len(m["positions"]) for m in meshes])[:3]; ib = np.cumsum([0] + [len(m["indices"]) for m in meshes])[:3]
    np.testing.assert_array_equal(blas["position_offset"], vb * 3); np.testing.assert_array_equal(blas["index_offset"], ib)
    np.testing.assert_array_equal(blas["num_triangles"], [len(m["indices"]) // 3 for m in meshes])
    # textures in order of first use: earth_diffuse (asset 6, sRGB), earth_normal (7), cage (10)
    for k, (name, fmt) in enumerate((("earth_diffuse", 43), ("earth_normal", 37), ("cage", 37))):
        want, wf = fx.read_texture(os.path.join(REF, "textures", f"{name}.texture.biasset"))
        got, gf = project.texture(k)
        assert gf == wf == fmt
        np.testing.assert_array_equal(got, want)


@needs_ref
def test_scene_matches_the_hand_assembled_fixture(project):
    """Same drawables / transforms / materials / light / camera as scenes.scene_basic(), which was written by reading scene.toml by hand
    (meshes appear in a different order there, so offsets are compared through the data they point to)."""
    ref = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz"))
    inst, rinst = project.array("instances"), ref.instances
    assert len(inst) == len(rinst) == 5
    np.testing.assert_allclose(inst["transform"], rinst["transform"], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(inst["instance_id_and_mask"], rinst["instance_id_and_mask"])
    np.testing.assert_array_equal(inst["sbt_offset_and_flags"], rinst["sbt_offset_and_flags"])          # opaque / non-opaque flags
    dr, rdr = project.array("drawables"), ref.drawables
    pos, rpos = project.array("positions"), ref.positions
    idx, ridx = project.array("indices"), ref.indices
    blas, rblas = project.array("blas"), ref.blas
    mats, rmats = project.array("materials"), ref.materials
    for i in range(5):
        b, rb = blas[int(inst["blas"][i])], rblas[int(rinst["blas"][i])]
        assert b["num_triangles"] == rb["num_triangles"]
        tri = pos[b["position_offset"]:].reshape(-1, 3)[idx[b["index_offset"]: b["index_offset"] + 3 * b["num_triangles"]]]
        rtri = rpos[rb["position_offset"]:].reshape(-1, 3)[ridx[rb["index_offset"]: rb["index_offset"] + 3 * rb["num_triangles"]]]
        np.testing.assert_array_equal(tri, rtri)
        m, rm = mats[dr["material_offset"][i] // capi.MATERIAL.itemsize], rmats[rdr["material_offset"][i] // capi.MATERIAL.itemsize]
        assert m["flags"] == rm["flags"]
        for f in ("base_color", "emission", "roughness"):
            np.testing.assert_array_equal(m[f], rm[f], err_msg=f"drawable {i} {f}")
        assert (m["base_color_tex"] >= 0) == (rm["base_color_tex"] >= 0) and (m["normal_map_tex"] >= 0) == (rm["normal_map_tex"] >= 0)
    dl, rdl = project.array("dir_lights"), ref.dir_lights
    np.testing.assert_allclose(dl["direction"], rdl["direction"], atol=1e-7); np.testing.assert_array_equal(dl["emission"], rdl["emission"])
    cam, rcam = project.camera(), ref.camera
    for k in ("position", "front_dir", "up_dir"):
        np.testing.assert_allclose(cam[k], rcam[k], atol=1e-6)
    assert cam["yfov"] == rcam["yfov"] and abs(cam["near_z"] - rcam["near_z"]) < 1e-9 and cam["far_z"] == rcam["far_z"]


@needs_ref
def test_loaded_project_renders_like_the_fixture_scene(project, oracle):
    """The loaded project through the oracle vs the hand-assembled fixture scene: identical geometry / lights / camera, so every pixel
    that does not see the earth-textured sphere (whose textures the fixture box-filters 4x4) is bit-equal; the sphere differs slightly."""
    ref = scenes.scene_basic(os.path.join(GOLDEN, "scene_basic.npz"))
    sd = project.scene_data()
    W, H = 96, 48
    imgs = []
    for scene in (sd, ref):
        scene.sky_faces = None
        ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
        ctx.render(oracle.camera_matrices(ref.camera, W, H), 0, 1, capi.Settings(max_bounces=2))   # primary hits + direct light only
        imgs.append(ctx.resolve(1)[..., :3])
    a, b = imgs
    assert np.isfinite(a).all() and a.mean() > 0.02
    same = (a == b).all(axis=2)
    assert same.mean() > 0.85 and np.abs(a - b).mean() < 0.01


def test_mini_project_round_trip(tmp_path, oracle):
    """A project written by tests/_mini_project.py (v1 and v2 meshes, a texture, three material snippets, dir + point light, camera,
    renderer override) loads back to exactly what was written, and renders through the oracle."""
    import _mini_project
    wrote = _mini_project.write(str(tmp_path))
    p = engine.Project(str(tmp_path))
    info = p.info
    assert (info.num_drawables, info.num_blas, info.num_materials, info.num_textures, info.num_dir_lights, info.num_point_lights) == (3, 2, 3, 1, 1, 1)
    assert (info.target_width, info.target_height, info.max_bounces, info.ray_length) == (96, 64, 4, 50.0)
    q, b = wrote["quad"], wrote["box"]
    np.testing.assert_array_equal(p.array("positions"), np.concatenate([q[0].reshape(-1), b[0].reshape(-1)]))
    np.testing.assert_array_equal(p.array("tangents"), np.concatenate([q[2].reshape(-1), b[2].reshape(-1)]))
    np.testing.assert_array_equal(p.array("indices"), np.concatenate([q[4], b[4]]))
    np.testing.assert_array_equal(p.array("blas")["num_triangles"], [2, 12])
    tex, fmt = p.texture(0)
    assert fmt == 37
    np.testing.assert_array_equal(tex, wrote["texture"])
    mats = p.array("materials")
    kinds = [(int(f) >> 8) & 0xff for f in mats["flags"]]
    assert kinds == [capi.MATERIAL_KIND_CHECKERBOARD, capi.MATERIAL_KIND_CONSTANT_COLOR, capi.MATERIAL_KIND_CAGE]
    assert (int(mats["flags"][2]) >> 16) & 0xff == capi.BLEND_ALPHA_TEST and int(mats["flags"][2]) & 1 == 1 and mats["base_color_tex"][2] == 0
    np.testing.assert_allclose(mats["base_color"][0], [0.8, 0.7, 0.6, 0.9]); np.testing.assert_allclose(mats["emission"][0], [0.1, 0.2, 0.3])
    inst = p.array("instances")
    assert [int(x) >> 24 for x in inst["sbt_offset_and_flags"]] == [capi.INSTANCE_FORCE_OPAQUE, capi.INSTANCE_FORCE_OPAQUE, capi.INSTANCE_FORCE_NON_OPAQUE]
    np.testing.assert_allclose(inst["transform"][2][:, 3], [-1.2, 0.75, 0.8])
    np.testing.assert_allclose(np.linalg.norm(inst["transform"][2][:, :3], axis=0), [1.5, 1.5, 1.5], rtol=1e-6)
    sd = p.scene_data()
    ctx = oracle.OracleContext(96, 64); ctx.upload_scene(sd, capi.ACCEL_TWO_LEVEL)
    ctx.render(oracle.camera_matrices(sd.camera, 96, 64), 0, 2, capi.Settings(max_bounces=info.max_bounces, ray_length=info.ray_length))
    img = ctx.resolve(2)
    assert np.isfinite(img).all() and img[..., :3].mean() > 0.02 and (img[..., :3].max(axis=2) > 0).mean() > 0.3
    p.close()


@pytest.mark.parametrize("storage", ["v1_raw", "v1_png"])
def test_texture_storage_variants(tmp_path, storage):
    """Texture v1 with raw texels and with one PNG per layer (TextureAsset::load, texture.cpp:105-132) load to the same texels as v2."""
    import _mini_project
    wrote = _mini_project.write(str(tmp_path))
    _mini_project.write_texture(str(tmp_path / "textures" / "cage.texture.biasset"), wrote["texture"], storage=storage)
    p = engine.Project(str(tmp_path))
    tex, fmt = p.texture(0)
    assert fmt == 37
    np.testing.assert_array_equal(tex, wrote["texture"])
    p.close()
    if storage == "v1_png":                                                # a PNG whose channel count does not match the format fails loudly
        import struct
        import _gltf_writer
        path = str(tmp_path / "textures" / "cage.texture.biasset")
        png = _gltf_writer.png_bytes(wrote["texture"][..., :3])
        head = open(path, "rb").read()
        cut = head.index(b"\x89PNG") - 8
        open(path, "wb").write(head[:cut] + struct.pack("<Q", len(png)) + png)
        with pytest.raises(RuntimeError, match="does not match the texture description"):
            engine.Project(str(tmp_path))


def test_mesh_with_out_of_range_indices_is_rejected(tmp_path):
    import _mini_project
    _mini_project.write(str(tmp_path))
    q = _mini_project.quad()
    bad = q[4].copy(); bad[-1] = len(q[0])                                  # one past the last vertex
    _mini_project.write_mesh(str(tmp_path / "meshes" / "plane.static_mesh.biasset"), q[0], q[1], q[2], q[3], bad)
    with pytest.raises(RuntimeError, match="beyond the vertex count"):
        engine.Project(str(tmp_path))


def test_toml_subset_and_error_paths(tmp_path):
    with pytest.raises(RuntimeError):
        engine.Project(str(tmp_path))                                              # no project.toml
    (tmp_path / "project.toml").write_text('name = "x"\nscene_file = "/project/s.toml"\nasset_metadata_file = "/project/a.toml"\n')
    (tmp_path / "a.toml").write_text("[[assets]]\nid = 0\npath = '/project/m.toml'\ntype = 'Material'\n")
    (tmp_path / "s.toml").write_text("[[objects]]\nname = 'only a light'\n  [[objects.components]]\n  type = 'Transform'\n")
    with pytest.raises(RuntimeError, match="no renderable object"):
        engine.Project(str(tmp_path))
    (tmp_path / "s.toml").write_text("[[objects]\nname = 'broken'\n")
    with pytest.raises(RuntimeError, match="toml line 1"):
        engine.Project(str(tmp_path))


def test_truncated_texture_payload_is_rejected(tmp_path):
    """width / height / format come from the asset header; a payload that holds fewer texels than they describe (truncated
    or hostile asset) must fail at load, before upload_project hands the short buffer to bpt_scene_upload_materials."""
    import _mini_project
    wrote = _mini_project.write(str(tmp_path))
    tex = wrote["texture"]
    for storage in ("v2", "v1_raw"):
        _mini_project.write_texture(str(tmp_path / "textures" / "cage.texture.biasset"), tex[: tex.shape[0] // 2], storage=storage)   # half the rows ...
        path = str(tmp_path / "textures" / "cage.texture.biasset")
        blob = bytearray(open(path, "rb").read())
        import struct
        at = blob.index(struct.pack("<IIII", tex.shape[1], tex.shape[0] // 2, 1, 1))
        blob[at + 4: at + 8] = struct.pack("<I", tex.shape[0])                                                                             # ... under the full height
        open(path, "wb").write(bytes(blob))
        with pytest.raises(RuntimeError, match="texel payload"):
            engine.Project(str(tmp_path))
    _mini_project.write_texture(str(tmp_path / "textures" / "cage.texture.biasset"), tex[:0], storage="v2")                              # height 0
    with pytest.raises(RuntimeError, match="extent"):
        engine.Project(str(tmp_path))


def test_toml_keeps_64_bit_asset_ids_and_bounds_nesting(tmp_path):
    """the engine's built-in assets have ids 2^62 + k (bisemutum/assets/asset_metadata.toml); as doubles they all collide.
    Two assets whose ids differ only below 2^53's resolution must stay distinct; nesting is bounded so a hostile file cannot overflow the stack."""
    import _mini_project
    _mini_project.write(str(tmp_path))
    meta = (tmp_path / "asset_metadata.toml").read_text()
    scene = (tmp_path / "scene.toml").read_text()
    import re
    ids = sorted({int(x) for x in re.findall(r"^id = (\d+)", meta, flags=re.M)})
    assert len(ids) >= 4
    big = {i: (1 << 62) + k for k, i in enumerate(ids)}                       # consecutive integers above 2^62: identical as doubles
    assert len({float(v) for v in big.values()}) == 1

    def remap(text, key):
        return re.sub(rf"({key} = )(\d+)", lambda m: m.group(1) + str(big.get(int(m.group(2)), int(m.group(2)))), text)
    (tmp_path / "asset_metadata.toml").write_text(remap(meta, "id"))
    (tmp_path / "scene.toml").write_text(remap(scene, "asset_id"))
    for f in (tmp_path / "materials").glob("*.toml"):
        f.write_text(remap(f.read_text(), "asset_id"))
    p = engine.Project(str(tmp_path))
    assert (p.info.num_drawables, p.info.num_blas, p.info.num_materials, p.info.num_textures) == (3, 2, 3, 1)
    kinds = [(int(f) >> 8) & 0xff for f in p.array("materials")["flags"]]
    assert kinds == [capi.MATERIAL_KIND_CHECKERBOARD, capi.MATERIAL_KIND_CONSTANT_COLOR, capi.MATERIAL_KIND_CAGE]   # each drawable found ITS material
    p.close()
    (tmp_path / "scene.toml").write_text("x = " + "[" * 5000 + "]" * 5000 + "\n")
    with pytest.raises(RuntimeError, match="nested deeper"):
        engine.Project(str(tmp_path))
