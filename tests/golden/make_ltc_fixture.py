"""Generates tests/golden/ltc_luts.npz from the reference's own LTC look-up tables
(/root/reference/bisemutum/assets/textures/ltc_*.texture.biasset; format: SURVEY.md Appendix C,
reader follows bisemutum/src/scene_basic/texture.cpp:83-137 and src/prelude/byte_stream.cpp:28-97).

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_ltc_fixture.py
The LUTs are input DATA the engine hands to bpt_scene_upload_lights at run time (float32, stored
zlib-compressed by numpy).
"""
import os
import struct
import sys
import zlib

import numpy as np

REF = "/root/reference/bisemutum/assets/textures"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ltc_luts.npz")


def read_biasset_texture(path):
    b = open(path, "rb").read()
    off = 0
    magic, = struct.unpack_from("<I", b, off); off += 4
    assert magic == 0x0B1A55E7, hex(magic)
    n, = struct.unpack_from("<Q", b, off); off += 8
    type_name = b[off:off + n].decode(); off += n
    assert type_name == "Texture", type_name
    version, = struct.unpack_from("<I", b, off); off += 4
    off += 28                                               # SamplerDesc (raw)
    w, h, d, levels = struct.unpack_from("<IIII", b, off); off += 16
    fmt, dim, usages, _pad = struct.unpack_from("<BBBB", b, off); off += 4
    if version >= 2:
        ulen, clen = struct.unpack_from("<QQ", b, off); off += 16
        raw = zlib.decompress(b[off:off + clen])
        assert len(raw) == ulen
        cnt, = struct.unpack_from("<Q", raw, 0)
        texels = raw[8:8 + cnt]
    else:
        storage, = struct.unpack_from("<I", b, off); off += 4
        assert storage == 0
        cnt, = struct.unpack_from("<Q", b, off); off += 8
        texels = b[off:off + cnt]
    return dict(width=w, height=h, depth=d, levels=levels, format=fmt, dim=dim, texels=texels)


def main():
    out = {}
    for name, ch in (("matrix_lut0", 4), ("matrix_lut1", 4), ("matrix_lut2", 4), ("norm_lut", 2)):
        t = read_biasset_texture(os.path.join(REF, f"ltc_{name}.texture.biasset"))
        assert (t["width"], t["height"], t["depth"]) == (8, 8, 64), t
        a = np.frombuffer(t["texels"], np.float32)
        assert a.size == 8 * 8 * 64 * ch, (a.size, ch, t["format"])
        out[name] = a.reshape(64, 8, 8, ch).copy()
        print(name, "format", t["format"], "range", a.min(), a.max())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
