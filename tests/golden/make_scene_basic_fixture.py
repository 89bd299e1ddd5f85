"""Generates tests/golden/scene_basic.npz from the reference's OWN example project
(/root/reference/examples/scene_basic): the three .biasset meshes verbatim, the cage texture verbatim
(1299x1300 RGBA8) and the two 2048x1024 earth textures box-filtered 4x4 to 512x256 (keeps the fixture small;
earth_diffuse is RGBA8 sRGB, format 43, earth_normal RGBA8 unorm, format 37). Formats: SURVEY.md Appendix C
(readers follow bisemutum/src/graphics/mesh.cpp:127-146 and bisemutum/src/scene_basic/texture.cpp:83-137).

    python tests/golden/make_scene_basic_fixture.py        (build container only)
"""
import os
import struct
import zlib

import numpy as np

REF = "/root/reference/examples/scene_basic"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scene_basic.npz")


def header(b):
    off = 0
    magic, = struct.unpack_from("<I", b, off); off += 4
    assert magic == 0x0B1A55E7
    n, = struct.unpack_from("<Q", b, off); off += 8
    type_name = b[off:off + n].decode(); off += n
    version, = struct.unpack_from("<I", b, off); off += 4
    return type_name, version, off


def compressed_part(b, off):
    ulen, clen = struct.unpack_from("<QQ", b, off); off += 16
    raw = zlib.decompress(b[off:off + clen])
    assert len(raw) == ulen
    return raw


def read_mesh(path):
    b = open(path, "rb").read()
    type_name, version, off = header(b)
    assert type_name == "StaticMesh"
    raw = compressed_part(b, off) if version >= 2 else b[off:]
    o = 0
    out = {}
    for name, dt, comps in (("positions", np.float32, 3), ("normals", np.float32, 3), ("tangents", np.float32, 4), ("colors", np.float32, 3),
                            ("texcoords", np.float32, 2), ("texcoords2", np.float32, 2), ("indices", np.uint32, 1), ("submeshes", np.uint8, 16)):
        cnt, = struct.unpack_from("<Q", raw, o); o += 8
        nbytes = cnt * comps * np.dtype(dt).itemsize
        out[name] = np.frombuffer(raw[o:o + nbytes], dt).reshape(cnt, comps).copy() if comps > 1 else np.frombuffer(raw[o:o + nbytes], dt).copy()
        o += nbytes
    return out


def read_texture(path):
    b = open(path, "rb").read()
    type_name, version, off = header(b)
    assert type_name == "Texture" and version >= 2
    off += 28
    w, h, d, levels = struct.unpack_from("<IIII", b, off); off += 16
    fmt, dim, usages, _ = struct.unpack_from("<BBBB", b, off); off += 4
    raw = compressed_part(b, off)
    cnt, = struct.unpack_from("<Q", raw, 0)
    return np.frombuffer(raw[8:8 + cnt], np.uint8).reshape(h, w, 4).copy(), fmt


def box4(img):
    h, w, c = img.shape
    return (img.reshape(h // 4, 4, w // 4, 4, c).astype(np.float64).mean((1, 3)) + 0.5).astype(np.uint8)


def main():
    out = {}
    for name in ("plane", "cube", "sphere"):
        m = read_mesh(os.path.join(REF, "meshes", f"{name}.static_mesh.biasset"))
        print(name, {k: v.shape for k, v in m.items()})
        for k in ("positions", "normals", "tangents", "texcoords", "indices"):
            out[f"{name}.{k}"] = m[k]
    cage, f = read_texture(os.path.join(REF, "textures", "cage.texture.biasset")); assert f == 37
    out["cage"] = cage
    ed, f = read_texture(os.path.join(REF, "textures", "earth_diffuse.texture.biasset")); assert f == 43
    en, f = read_texture(os.path.join(REF, "textures", "earth_normal.texture.biasset")); assert f == 37
    out["earth_diffuse_srgb"] = box4(ed)
    out["earth_normal"] = box4(en)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
