"""glTF import (BASELINE configs[0] "Cornell-box glTF"): host/gltf.cpp restates the reference's editor action
`menu_action_import_model_gltf` (import_model.cpp:27-430) + `StaticMesh::calculate_tspace` (static_mesh.cpp:93-152) headless. The files are
written by tests/_gltf_writer.py (numpy / json / zlib only), so the importer reads bytes it did not produce; expectations come from the
arrays that went into the files and from float64 numpy restatements of the node transforms and of the tangent definition."""
import numpy as np
import pytest

import _gltf_writer as gw
from bisemutum_engine_b200 import capi, engine, scenes


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-30)


def _reference_tangents(pos, nrm, uv, idx):
    """Tangent definition MikkTSpace implements, for meshes whose index sharing equals its welding (no duplicated (p, n, uv) triples) and
    whose fans are orientation-consistent: angle-weighted sum over the non-degenerate incident faces of the face's unit dP/du direction
    projected into the vertex' tangent plane. float64. Returns (tangent xyz, sign, vertex has a non-degenerate face)."""
    pos, nrm, uv = (np.asarray(a, np.float64) for a in (pos, nrm, uv))
    tri = np.asarray(idx, np.int64).reshape(-1, 3)
    p, t = pos[tri], uv[tri]
    degenerate = (np.all(p[:, 0] == p[:, 1], 1) | np.all(p[:, 0] == p[:, 2], 1) | np.all(p[:, 1] == p[:, 2], 1))
    d1, d2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
    t21, t31 = t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]
    area = t21[:, 0] * t31[:, 1] - t21[:, 1] * t31[:, 0]
    os_ = _unit(t31[:, 1:2] * d1 - t21[:, 1:2] * d2) * np.where(area > 0, 1.0, -1.0)[:, None]
    usable = ~degenerate & (np.abs(area) > 0)
    acc = np.zeros_like(pos); sign = np.zeros(len(pos)); seen = np.zeros(len(pos), bool)
    for c in range(3):
        v = tri[:, c]
        n = nrm[v]
        proj = lambda x: _unit(x - n * np.sum(n * x, -1, keepdims=True))
        e1, e2 = proj(p[:, (c + 2) % 3] - p[:, c]), proj(p[:, (c + 1) % 3] - p[:, c])
        ang = np.arccos(np.clip(np.sum(e1 * e2, -1), -1, 1))
        np.add.at(acc, v[usable], (ang[:, None] * proj(os_))[usable])
        sign[v[usable]] = np.where(area > 0, 1.0, -1.0)[usable]
        seen[v[usable]] = True
    return _unit(acc), sign, seen


def _mikk(pos, nrm, uv, idx, base_vertex=0):
    import ctypes as C
    h = engine.host_library()
    pos, nrm, uv = (np.ascontiguousarray(a, np.float32) for a in (pos, nrm, uv))
    idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
    out = np.zeros((len(pos.reshape(-1, 3)), 4), np.float32)
    h.bpt_host_mikk_tangents.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_uint32]
    h.bpt_host_mikk_tangents.restype = None
    h.bpt_host_mikk_tangents(pos.ctypes.data, nrm.ctypes.data, uv.ctypes.data, out.ctypes.data, idx.ctypes.data, idx.size, base_vertex)
    return out


def test_tangents_follow_the_mikktspace_definition():
    # a flat grid: the tangent is exactly the direction of increasing u; mirrored u flips the sign and the direction
    p, n, t, uv, idx = scenes.grid_patch(5, 3, lambda u, v: np.stack([2 * u, 0 * u, -3 * v], -1))
    tg = _mikk(p, n, uv, idx)
    np.testing.assert_allclose(tg, np.tile(np.float32([1, 0, 0, 1]), (len(p), 1)), rtol=0, atol=1.2e-7)    # 1 / |v| * v: one ulp
    uv_m = uv.copy(); uv_m[:, 0] = 1 - uv_m[:, 0]
    np.testing.assert_allclose(_mikk(p, n, uv_m, idx), np.tile(np.float32([-1, 0, 0, -1]), (len(p), 1)), rtol=0, atol=1.2e-7)
    # curved, non-uniformly parametrised patch and a sphere with poles (degenerate triangles) and a uv seam: the float64 definition
    for (p, n, t, uv, idx) in (scenes.grid_patch(9, 7, lambda u, v: np.stack([u * u + 0.3 * v, 0.4 * np.sin(3 * u) * np.cos(2 * v), v + 0.2 * u], -1)),
                               scenes.sphere_mesh(0.7, 12, 8)):
        tg = _mikk(p, n, uv, idx)
        ref, sign, seen = _reference_tangents(p, n, uv, idx)
        assert seen.mean() > 0.8
        np.testing.assert_allclose(tg[seen, :3], ref[seen], atol=2e-6)
        np.testing.assert_array_equal(tg[seen, 3], sign[seen])
        np.testing.assert_allclose(np.linalg.norm(tg[seen, :3], axis=1), 1.0, atol=1e-6)
        assert np.abs(np.sum(tg[seen, :3] * n[seen], 1)).max() < 1e-6                       # in the tangent plane of the vertex normal
    # base_vertex offsets both the reads and the writes; vertices that no index touches keep what they held
    p, n, t, uv, idx = scenes.grid_patch(2, 2, lambda u, v: np.stack([u, v, 0 * u], -1))
    pad = 3
    P = np.concatenate([np.zeros((pad, 3), np.float32), p]); N = np.concatenate([np.zeros((pad, 3), np.float32), n]); UV = np.concatenate([np.zeros((pad, 2), np.float32), uv])
    tg = _mikk(P, N, UV, idx, base_vertex=pad)
    assert (tg[:pad] == 0).all() and np.abs(tg[pad:] - np.float32([1, 0, 0, 1])).max() < 1.2e-7
    # zero uv area: no usable tangent anywhere -> MikkTSpace's default (1, 0, 0) with orientation flag clear; fully degenerate positions likewise
    tg = _mikk(p, n, np.zeros_like(uv), idx)
    assert np.isfinite(tg).all() and (tg[:, 3] == -1).all()
    tg = _mikk(np.zeros_like(p), n, uv, idx)
    np.testing.assert_array_equal(tg, np.tile(np.float32([1, 0, 0, -1]), (len(p), 1)))


@pytest.mark.parametrize("container,index_dtype,strided", [("gltf+bin", np.uint32, False), ("gltf+data", np.uint16, True), ("glb", np.uint16, False)])
def test_cornell_box_gltf_round_trip(tmp_path, container, index_dtype, strided):
    """configs[0]: the Cornell box written as glTF and imported: geometry streams, BLAS / drawable / instance records and materials equal
    what went in; tangents equal the generator's analytic dP/du on the flat walls and the float64 definition elsewhere."""
    src = scenes.cornell_box(tess=4)
    path = str(tmp_path / ("cornell." + ("glb" if container == "glb" else "gltf")))
    gw.from_scene(src, index_dtype=index_dtype, strided=strided).write(path, container)
    p = engine.Project.from_gltf(path)
    n_inst = len(src.instances)
    assert (p.info.num_drawables, p.info.num_materials, p.info.num_textures) == (n_inst, len(src.materials), 0)
    inst, blas, dr = p.array("instances"), p.array("blas"), p.array("drawables")
    pos, nrm, tan, uv, idx = (p.array(k) for k in ("positions", "normals", "tangents", "texcoords", "indices"))
    assert len(tan) // 4 == len(pos) // 3 == len(nrm) // 3 == len(uv) // 2
    mats = p.array("materials")
    for f in ("base_color", "emission", "roughness", "metallic", "flags"):
        np.testing.assert_array_equal(mats[f], src.materials[f])
    assert (mats["base_color_tex"] == -1).all() and (mats["normal_map_scale"] == 1).all() and (mats["occlusion_strength"] == 1).all()
    for k in range(n_inst):
        s_in = src.instances[k]
        s_bd = src.blas[int(s_in["blas"])]
        bd = blas[int(inst["blas"][k])]
        assert int(inst["instance_id_and_mask"][k]) == (k | 0xFF000000) and int(inst["sbt_offset_and_flags"][k]) == (k | (capi.INSTANCE_FORCE_OPAQUE << 24))
        assert int(dr["drawable_index"][k]) == k and int(bd["num_triangles"]) == int(s_bd["num_triangles"])
        assert int(dr["material_offset"][k]) == int(src.drawables[k]["material_offset"])
        assert int(dr["position_offset"][k]) == int(bd["position_offset"]) and int(dr["index_offset"][k]) == int(bd["index_offset"])
        assert int(dr["tangent_offset"][k]) * 3 == int(dr["position_offset"][k]) * 4 and int(dr["texcoord_offset"][k]) * 3 == int(dr["position_offset"][k]) * 2
        nt = int(bd["num_triangles"])
        s_idx = src.indices[int(s_bd["index_offset"]): int(s_bd["index_offset"]) + 3 * nt]
        np.testing.assert_array_equal(idx[int(bd["index_offset"]): int(bd["index_offset"]) + 3 * nt], s_idx)
        nv = int(s_idx.max()) + 1
        v0, s0 = int(bd["position_offset"]) // 3, int(s_bd["position_offset"]) // 3
        np.testing.assert_array_equal(pos.reshape(-1, 3)[v0:v0 + nv], src.positions.reshape(-1, 3)[s0:s0 + nv])
        np.testing.assert_array_equal(nrm.reshape(-1, 3)[v0:v0 + nv], src.normals.reshape(-1, 3)[s0:s0 + nv])
        np.testing.assert_array_equal(uv.reshape(-1, 2)[v0:v0 + nv], src.texcoords.reshape(-1, 2)[s0:s0 + nv])
        # node matrix -> Transform::from_matrix -> matrix(): the rotation columns are renormalised in FP32
        np.testing.assert_allclose(inst["transform"][k], s_in["transform"], rtol=0, atol=3e-7)
        got = tan.reshape(-1, 4)[v0:v0 + nv]
        ref, sign, seen = _reference_tangents(pos.reshape(-1, 3)[v0:v0 + nv], nrm.reshape(-1, 3)[v0:v0 + nv], uv.reshape(-1, 2)[v0:v0 + nv], s_idx)
        np.testing.assert_allclose(got[seen, :3], ref[seen], atol=2e-6)
        np.testing.assert_array_equal(got[seen, 3], sign[seen])
        if k < 5:                                                           # the five flat walls: the generator's analytic tangent frame
            np.testing.assert_allclose(got, src.tangents.reshape(-1, 4)[s0:s0 + nv], atol=1e-6)
    p.close()


def test_node_hierarchy_trs_and_names(tmp_path):
    """TRS + quaternion + matrix nodes, nested three deep, one mesh with two primitives and a LINES primitive in between: world transforms
    against float64 composition, drawable order = depth-first node order x primitive order, name de-duplication (import_model.cpp:367-372)."""
    p_, n_, t_, uv_, i_ = scenes.grid_patch(2, 2, lambda u, v: np.stack([u, v, 0 * u], -1))
    b = gw.GltfBuilder()
    m0, m1 = b.material((0.5, 0.25, 0.125, 1.0), roughness=0.5, metallic=0.0), b.material((1, 1, 1, 0.5), double_sided=True, emissive=(1, 2, 3))
    mesh = b.mesh([b.primitive(p_, n_, uv_, i_, m0), b.primitive(p_, None, None, np.arange(4), m0, mode=1), b.primitive(p_ * 2, n_, None, i_[::-1].copy(), m1, index_dtype=np.uint8)])
    q = np.float64([0.1, 0.7, -0.2, 0.6]); q /= np.linalg.norm(q)
    leaf = b.node(mesh=mesh, translation=(0.5, 0, 0), scale=(1, 2, 3), name="dup", root=False)
    mtx = np.eye(4); mtx[:3, :3] = np.float64([[0, 0, 2], [0, 2, 0], [-2, 0, 0]]); mtx[:3, 3] = (1, 2, 3)
    mid = b.node(mesh=mesh, matrix=mtx, translation=(9, 9, 9), children=[leaf], name="dup", root=False)     # matrix wins over TRS
    b.node(rotation=q, translation=(0, 1, 0), children=[mid], name="top")
    path = str(tmp_path / "nodes.gltf")
    b.write(path, "gltf+data")
    p = engine.Project.from_gltf(path)
    x, y, z, w = q
    R = np.float64([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                    [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    top = np.eye(4); top[:3, :3] = R; top[:3, 3] = (0, 1, 0)
    lf = np.eye(4); lf[:3, :3] = np.diag([1.0, 2, 3]); lf[:3, 3] = (0.5, 0, 0)
    inst = p.array("instances")
    assert len(inst) == 4 and p.info.num_blas == 2 and p.info.num_materials == 2
    np.testing.assert_allclose(inst["transform"][0], (top @ mtx)[:3], atol=2e-6)
    np.testing.assert_allclose(inst["transform"][1], (top @ mtx)[:3], atol=2e-6)
    np.testing.assert_allclose(inst["transform"][2], (top @ mtx @ lf)[:3], atol=4e-6)
    assert [int(b_) for b_ in inst["blas"]] == [0, 1, 0, 1]
    dr, mats = p.array("drawables"), p.array("materials")
    assert [int(o) // mats.dtype.itemsize for o in dr["material_offset"]] == [0, 1, 0, 1]
    assert int(mats["flags"][1]) & 1 == 1 and int(mats["flags"][0]) & 1 == 0 and (int(mats["flags"][1]) >> 16) & 0xFF == capi.BLEND_OPAQUE   # alpha mode is a TODO upstream
    np.testing.assert_array_equal(mats["base_color"][1], np.float32([1, 1, 1, 0.5])); np.testing.assert_array_equal(mats["emission"][1], np.float32([1, 2, 3]))
    assert mats["metallic"][1] == 1 and mats["roughness"][1] == 1                            # glTF defaults
    # second primitive: u8 indices, missing NORMAL / TEXCOORD_0 zero-filled, base_vertex 9
    blas, idx, uv = p.array("blas"), p.array("indices"), p.array("texcoords").reshape(-1, 2)
    assert int(blas["position_offset"][1]) == 9 * 3 and int(blas["index_offset"][1]) == i_.size and int(blas["num_triangles"][1]) == i_.size // 3
    np.testing.assert_array_equal(idx[i_.size:], i_[::-1].reshape(-1))
    assert (uv[9:] == 0).all() and np.isfinite(p.array("tangents")).all()
    p.close()


def test_png_textures_and_samplers(tmp_path):
    rng = np.random.default_rng(7)
    rgba = rng.integers(0, 256, (13, 9, 4), dtype=np.uint8)
    rgb = rng.integers(0, 256, (6, 10, 3), dtype=np.uint8)
    grey = rng.integers(0, 256, (5, 4), dtype=np.uint8)
    ga = rng.integers(0, 256, (4, 7, 2), dtype=np.uint8)
    p_, n_, t_, uv_, i_ = scenes.grid_patch(1, 1, lambda u, v: np.stack([u, v, 0 * u], -1))
    for container in ("glb", "gltf+bin"):
        b = gw.GltfBuilder()
        t0 = b.texture(gw.png_bytes(rgba), "view", sampler={"magFilter": 9728, "minFilter": 9984, "wrapS": 33071})
        t1 = b.texture(gw.png_bytes(rgb, filters=(4, 3)), "uri")
        t2 = b.texture(gw.png_bytes(grey, filters=(1,)), "view")
        t3 = b.texture(gw.png_bytes(ga, filters=(2, 0)), "uri", sampler={"wrapT": 10497})
        m = b.material(base_color_tex=t1, normal_tex=t3, normal_scale=0.5)
        b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_, m)]))
        path = str(tmp_path / ("tex." + ("glb" if container == "glb" else "gltf")))
        b.write(path, container)
        p = engine.Project.from_gltf(path)
        assert p.info.num_textures == 4
        np.testing.assert_array_equal(p.texture(t0)[0], rgba)
        np.testing.assert_array_equal(p.texture(t1)[0], np.concatenate([rgb, np.full((6, 10, 1), 255, np.uint8)], 2))
        np.testing.assert_array_equal(p.texture(t2)[0], np.stack([grey, grey, grey, np.full_like(grey, 255)], 2))
        np.testing.assert_array_equal(p.texture(t3)[0], np.stack([ga[..., 0]] * 3 + [ga[..., 1]], 2))
        assert all(p.texture(k)[1] == 37 for k in range(4))                                  # rgba8_unorm (import_model.cpp:90)
        mats = p.array("materials")
        assert (int(mats["base_color_tex"][0]), int(mats["normal_map_tex"][0]), int(mats["metallic_roughness_tex"][0]), int(mats["occlusion_tex"][0])) == (t1, t3, -1, -1)
        assert mats["normal_map_scale"][0] == 0.5
        p.close()


def test_import_fails_loudly(tmp_path):
    p_, n_, t_, uv_, i_ = scenes.grid_patch(1, 1, lambda u, v: np.stack([u, v, 0 * u], -1))

    def attempt(build, match, container="gltf+data"):
        b = gw.GltfBuilder()
        build(b)
        path = str(tmp_path / "bad.gltf")
        b.write(path, container)
        with pytest.raises(RuntimeError, match=match):
            engine.Project.from_gltf(path)
    with pytest.raises(RuntimeError, match="cannot read"):
        engine.Project.from_gltf(str(tmp_path / "missing.gltf"))
    (tmp_path / "broken.gltf").write_text('{"asset": {"version": "2.0"}, "scenes": [}')
    with pytest.raises(RuntimeError, match="json"):
        engine.Project.from_gltf(str(tmp_path / "broken.gltf"))
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, None, b.material())])), "non-indexed")
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_, None)])), "valid material")
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_ + 7, b.material())])), "out of range")
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_[:2], uv_, i_, b.material())])), "NORMAL has 2 elements")
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_, b.material(base_color_tex=b.texture(b"\xff\xd8\xff\xe0 not a png")))])), "not a PNG")
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_, b.material(base_color_tex=3))])), "texture that does not exist")
    attempt(lambda b: b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_, b.material(base_color_tex=b.texture(gw.png_bytes(np.zeros((2, 2, 3), np.uint8)), sampler={"wrapS": 33648})))])), "mirrored")
    attempt(lambda b: b.mesh([b.primitive(p_, n_, uv_, i_, b.material())]), "no renderable primitive")
    b = gw.GltfBuilder(); b.node(mesh=b.mesh([b.primitive(p_, n_, uv_, i_, b.material())]))
    b.write(str(tmp_path / "short.gltf"), "gltf+bin")
    with open(str(tmp_path / "short data.bin"), "r+b") as f:
        f.truncate(16)
    with pytest.raises(RuntimeError, match="shorter than its byteLength"):
        engine.Project.from_gltf(str(tmp_path / "short.gltf"))


def test_index_accessor_with_a_hostile_byte_stride(tmp_path):
    """the bounds check of an accessor steps by bufferView.byteStride, so the index loop must read with that same step.
    byteStride = 1 on a u32 index view placed LAST in the buffer passes the check with count + 3 bytes; reading it as tightly packed
    would run 4 * count bytes past it. The importer must refuse (or read inside the checked range) — never over-read."""
    p_, n_, t_, uv_, i_ = scenes.grid_patch(8, 8, lambda u, v: np.stack([u, v, 0 * u], -1))
    b = gw.GltfBuilder()
    mat = b.material()
    attrs = {"POSITION": b.accessor(np.float32(p_).reshape(-1, 3), "VEC3"), "NORMAL": b.accessor(np.float32(n_).reshape(-1, 3), "VEC3"),
             "TEXCOORD_0": b.accessor(np.float32(uv_).reshape(-1, 2), "VEC2")}
    count = len(i_)
    view = b._view(bytes(count + 3), stride=1)                                # the LAST bytes of the buffer: count + 3 of them
    b.doc["accessors"].append({"bufferView": view, "componentType": 5125, "count": count, "type": "SCALAR"})
    b.node(mesh=b.mesh([{"attributes": attrs, "indices": len(b.doc["accessors"]) - 1, "material": mat}]))
    path = str(tmp_path / "stride.gltf")
    b.write(path, "gltf+data")
    with pytest.raises(RuntimeError, match="byteStride smaller than the index size"):
        engine.Project.from_gltf(path)
    # a legal (if unusual) stride >= the element size is read with that stride
    b2 = gw.GltfBuilder()
    mat = b2.material()
    attrs = {"POSITION": b2.accessor(np.float32(p_).reshape(-1, 3), "VEC3"), "NORMAL": b2.accessor(np.float32(n_).reshape(-1, 3), "VEC3"),
             "TEXCOORD_0": b2.accessor(np.float32(uv_).reshape(-1, 2), "VEC2")}
    idx = b2.accessor(np.uint32(i_).reshape(-1), "SCALAR", interleave_pad=4)  # u32 index, 4 junk bytes, ...
    b2.node(mesh=b2.mesh([{"attributes": attrs, "indices": idx, "material": mat}]))
    path2 = str(tmp_path / "stride8.gltf")
    b2.write(path2, "gltf+data")
    pr = engine.Project.from_gltf(path2)
    np.testing.assert_array_equal(pr.array("indices"), np.uint32(i_).reshape(-1))
    pr.close()


def test_imported_cornell_box_renders_through_the_oracle(tmp_path, oracle):
    """The imported model carries everything the renderer needs: with the source scene's camera and light it renders through the oracle, and the
    image agrees with the hand-assembled scene wherever tangents do not matter (primary visibility + direct light of rough dielectrics: the
    tangent frame only rotates the sampled bounce direction)."""
    src = scenes.cornell_box(tess=4)
    path = str(tmp_path / "cornell.glb")
    gw.from_scene(src, index_dtype=np.uint16).write(path, "glb")
    p = engine.Project.from_gltf(path)
    sd = p.scene_data()
    sd.dir_lights, sd.camera = src.dir_lights, src.camera
    W, H = 64, 64
    imgs = []
    for scene in (sd, src):
        ctx = oracle.OracleContext(W, H); ctx.upload_scene(scene, capi.ACCEL_TWO_LEVEL)
        ctx.render(oracle.camera_matrices(src.camera, W, H), 0, 1, capi.Settings(max_bounces=2))
        imgs.append(ctx.resolve(1)[..., :3]); ctx.close()
    a, b = imgs
    assert np.isfinite(a).all() and a.mean() > 0.02
    assert np.abs(a - b).max() < 1e-3 * max(1.0, float(b.max()))
    p.close()


@pytest.mark.gpu
def test_imported_gltf_uploads_and_renders_on_the_gpu(tmp_path, oracle):
    """configs[0] end to end: Cornell-box glTF -> C++ importer -> bpt_host_project_upload -> CUDA path tracer, 512x512 at max depth 5 with one
    directional light, against the oracle fed the same imported arrays through the Python path (per-sample radiance within 1e-4, ray counts equal)."""
    import bisemutum_engine_b200 as pkg
    lib = pkg.load_library()
    src = scenes.cornell_box(tess=8)
    path = str(tmp_path / "cornell.glb")
    b = gw.from_scene(src)
    # plus a textured, normal-mapped panel (PNG base colour with a nearest / clamp sampler, PNG normal map at scale 0.5): the material
    # textures, their samplers and the MikkTSpace frame all reach the shade kernel
    rng = np.random.default_rng(3)
    base = b.texture(gw.png_bytes(rng.integers(40, 256, (16, 16, 3), dtype=np.uint8)), "view", sampler={"magFilter": 9728, "wrapS": 33071})
    nm = np.concatenate([rng.integers(96, 160, (8, 8, 2), dtype=np.uint8), np.full((8, 8, 1), 255, np.uint8)], 2)
    mat = b.material((0.9, 0.9, 0.9, 1.0), roughness=0.6, metallic=0.0, base_color_tex=base, normal_tex=b.texture(gw.png_bytes(nm), "uri"), normal_scale=0.5)
    pp, pn, pt, puv, pi = scenes.grid_patch(3, 3, lambda u, v: np.stack([-0.9 + 0.8 * u, 1.0 + 0.7 * v, -0.6 + 0.5 * u], -1))
    b.node(mesh=b.mesh([b.primitive(pp, pn, puv * 1.5 - 0.25, pi, mat)]), name="panel")
    b.write(path, "glb")
    p = engine.Project.from_gltf(path)
    assert p.info.num_textures == 2 and p.info.num_drawables == len(src.instances) + 1
    sd = p.scene_data()
    assert (sd.textures[0]["linear"], sd.textures[0]["address_u"], sd.textures[0]["address_v"]) == (0, capi.ADDRESS_CLAMP, capi.ADDRESS_REPEAT)
    sd.dir_lights, sd.camera = src.dir_lights, src.camera
    W = H = 512
    cam = engine.camera_matrices(src.camera, W, H)
    st = capi.Settings(max_bounces=5)
    for mode in (capi.ACCEL_TWO_LEVEL, capi.ACCEL_MERGED):
        gpu = capi.Context(lib, W, H)
        p.upload(gpu, mode)
        gpu.upload_lights(sd)
        ref = oracle.OracleContext(W, H); ref.upload_scene(sd, mode)
        gpu.render(cam, 0, 2, st); ref.render(cam, 0, 2, st)
        a, b = gpu.resolve(2), ref.resolve(2)
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6)
        ca, cb = gpu.counters(), ref.counters()
        assert ca.extend_rays == cb.extend_rays and ca.shadow_rays == cb.shadow_rays and a[..., :3].mean() > 0.02
        gpu.close(); ref.close()
    p.close()


# ---- committed fixture: tests/golden/cornell_box.glb (input) + cornell_box_gltf.npz (import result + oracle image), made by make_gltf_fixture.py
def _fixture():
    import importlib.util
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_gltf_fixture", os.path.join(d, "make_gltf_fixture.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m, np.load(m.NPZ)


def test_committed_glb_imports_to_the_committed_arrays(oracle):
    m, gold = _fixture()
    arrays, sd = m.imported()
    for k, v in arrays.items():
        np.testing.assert_array_equal(v.view(np.uint8) if v.dtype.names else v, gold[k], err_msg=k)
    np.testing.assert_array_equal(m.render(oracle.OracleContext, sd, capi.ACCEL_TWO_LEVEL), gold["image"])
    # merged mode bakes the instance transforms into the vertices: same scene, hits differ in the last bits
    np.testing.assert_allclose(m.render(oracle.OracleContext, sd, capi.ACCEL_MERGED), gold["image"], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_committed_glb_renders_to_the_golden_image_on_the_gpu():
    import bisemutum_engine_b200 as pkg
    m, gold = _fixture()
    _, sd = m.imported()
    lib = pkg.load_library()
    np.testing.assert_array_equal(m.render(lambda w, h: capi.Context(lib, w, h), sd, capi.ACCEL_TWO_LEVEL), gold["image"])   # one light: no summation-order freedom
    np.testing.assert_allclose(m.render(lambda w, h: capi.Context(lib, w, h), sd, capi.ACCEL_MERGED), gold["image"], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_native_host_renders_the_committed_glb(oracle, tmp_path):
    """render_project model.glb (C++ only: glTF import -> C ABI -> CudaPathTracingRenderer -> PostProcessPass -> PFM) with camera and light
    from the command line == the oracle's render + output pass of the imported scene."""
    import os
    import subprocess
    import bisemutum_engine_b200 as pkg
    m, _ = _fixture()
    _, sd = m.imported()
    exe = os.path.join(pkg.PACKAGE_DIR, "host", "render_project")
    assert os.path.exists(exe), "build it first: make -C bisemutum-engine_b200/host (there is no fallback)"
    out = str(tmp_path / "out.pfm")
    cam, light = sd.camera, sd.dir_lights[0]
    emission = np.float32(light["emission"])
    args = [exe, m.GLB, out, "3", "80", "56", "--bounces", "5", "--camera", *(repr(float(v)) for v in (*cam["position"], *cam["front_dir"])),
            "--dir-light", *(repr(float(np.float32(v))) for v in light["direction"]), *(repr(float(v)) for v in emission), "1"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(out, "rb") as f:
        assert f.readline() == b"PF\n"
        w, h = (int(v) for v in f.readline().split())
        f.readline()
        img = np.frombuffer(f.read(), "<f4").reshape(h, w, 3)[::-1]
    assert (w, h) == (80, 56)
    ref = oracle.OracleContext(w, h); ref.upload_scene(sd, capi.ACCEL_TWO_LEVEL)
    ref.render(oracle.camera_matrices(cam, w, h), 0, 3, capi.Settings(max_bounces=5))
    want = oracle.post_process_image(ref.resolve(3), capi.PostSettings(False, 1.5, 0.5))[..., :3]
    np.testing.assert_allclose(img, want, rtol=1e-4, atol=1e-6)
    assert img.mean() > 0.02
    ref.close()


def test_mutated_files_never_crash_the_importer(tmp_path):
    """Robustness: 400 seeded mutations of the committed .glb (byte flips in the JSON chunk, number replacements, truncations) either import or
    raise — the importer runs in this process, so a wild read would take the test run down."""
    import re
    import struct
    m, _ = _fixture()
    blob = open(m.GLB, "rb").read()
    jlen = struct.unpack_from("<I", blob, 12)[0]
    js, rest = blob[20:20 + jlen], blob[20 + jlen:]
    rng = np.random.default_rng(11)
    numbers = [mm.span() for mm in re.finditer(rb"-?\d+(\.\d+)?", js)]
    outcomes = {"ok": 0, "error": 0}
    for k in range(400):
        j = bytearray(js)
        kind = k % 4
        if kind == 0:                                                        # flip a few bytes
            for pos in rng.integers(0, len(j), 3):
                j[pos] = int(rng.integers(32, 127))
        elif kind == 1:                                                      # replace one number by a hostile one
            a, b_ = numbers[int(rng.integers(0, len(numbers)))]
            j[a:b_] = [b"-1", b"4294967295", b"1e300", b"-1e300", b"0", b"99999999", b"2.5", b"18446744073709551615"][int(rng.integers(0, 8))]
        elif kind == 2:                                                      # truncate the JSON
            j = j[:int(rng.integers(1, len(j)))]
        body = bytes(j) + b" " * (-len(j) % 4)
        data = blob[:12] + struct.pack("<II", len(body), 0x4E4F534A) + body + rest
        if kind == 3:                                                        # truncate the whole file (binary chunk included)
            data = data[:int(rng.integers(12, len(data)))]
        path = str(tmp_path / "mut.glb")
        open(path, "wb").write(data)
        try:
            engine.Project.from_gltf(path).close()
            outcomes["ok"] += 1
        except RuntimeError:
            outcomes["error"] += 1
    assert outcomes["error"] > 100 and outcomes["ok"] + outcomes["error"] == 400
