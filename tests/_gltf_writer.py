"""Test helper: writes glTF 2.0 files (.gltf + .bin, .gltf with data: URIs, .glb) and PNG images with nothing but numpy / zlib / json,
so the C++ importer (bisemutum-engine_b200/host/gltf.cpp) is exercised on files it did not produce itself. Independent of the importer:
the byte layout here follows the Khronos glTF 2.0 / PNG specifications."""
import base64
import json
import struct
import zlib

import numpy as np


def png_bytes(img: np.ndarray, filters=(0, 1, 2, 3, 4)) -> bytes:
    """(H, W) / (H, W, 2|3|4) uint8 -> PNG; row y uses filter type filters[y % len(filters)] so that every PNG predictor is covered."""
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, ch = img.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    rows = img.reshape(h, w * ch).astype(np.int32)
    out = bytearray()
    zero = np.zeros(w * ch, np.int32)
    for y in range(h):
        f = filters[y % len(filters)]
        cur, up = rows[y], rows[y - 1] if y else zero
        a = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        c = np.concatenate([np.zeros(ch, np.int32), up[:-ch]])
        if f == 0:
            pred = zero
        elif f == 1:
            pred = a
        elif f == 2:
            pred = up
        elif f == 3:
            pred = (a + up) >> 1
        else:
            p = a + up - c
            pa, pb, pc = np.abs(p - a), np.abs(p - up), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, up, c))
        out.append(f)
        out += ((cur - pred) & 0xFF).astype(np.uint8).tobytes()

    def chunk(kind: bytes, body: bytes) -> bytes:
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    z = zlib.compress(bytes(out), 6)
    half = len(z) // 2                                                       # two IDAT chunks: the decoder must concatenate them
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) + chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:])
            + chunk(b"IEND", b""))


class GltfBuilder:
    def __init__(self):
        self.doc = {"asset": {"version": "2.0"}, "buffers": [], "bufferViews": [], "accessors": [], "meshes": [], "materials": [], "nodes": [],
                    "scenes": [{"nodes": []}], "scene": 0}
        self.blob = bytearray()

    def _view(self, data: bytes, stride: int = 0) -> int:
        while len(self.blob) % 4:
            self.blob.append(0)
        v = {"buffer": 0, "byteOffset": len(self.blob), "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        self.blob += data
        self.doc["bufferViews"].append(v)
        return len(self.doc["bufferViews"]) - 1

    def accessor(self, arr: np.ndarray, kind: str, interleave_pad: int = 0) -> int:
        arr = np.ascontiguousarray(arr)
        ctype = {np.dtype(np.float32): 5126, np.dtype(np.uint32): 5125, np.dtype(np.uint16): 5123, np.dtype(np.uint8): 5121}[arr.dtype]
        count = arr.shape[0]
        if interleave_pad:                                                   # strided view: element, then `interleave_pad` junk bytes
            elem = arr.reshape(count, -1).view(np.uint8).reshape(count, -1)
            padded = np.concatenate([elem, np.full((count, interleave_pad), 0xAB, np.uint8)], 1)
            view = self._view(padded.tobytes(), stride=padded.shape[1])
        else:
            view = self._view(arr.tobytes())
        acc = {"bufferView": view, "componentType": ctype, "count": count, "type": kind}
        if kind == "VEC3" and arr.dtype == np.float32:
            acc["min"], acc["max"] = arr.reshape(count, 3).min(0).tolist(), arr.reshape(count, 3).max(0).tolist()
        self.doc["accessors"].append(acc)
        return len(self.doc["accessors"]) - 1

    def primitive(self, pos, nrm, uv, idx, material, index_dtype=np.uint32, strided=False, mode=None) -> dict:
        attrs = {"POSITION": self.accessor(np.asarray(pos, np.float32).reshape(-1, 3), "VEC3", 4 if strided else 0)}
        if nrm is not None:
            attrs["NORMAL"] = self.accessor(np.asarray(nrm, np.float32).reshape(-1, 3), "VEC3")
        if uv is not None:
            attrs["TEXCOORD_0"] = self.accessor(np.asarray(uv, np.float32).reshape(-1, 2), "VEC2", 8 if strided else 0)
        p = {"attributes": attrs}
        if idx is not None:
            p["indices"] = self.accessor(np.asarray(idx).reshape(-1).astype(index_dtype), "SCALAR")
        if material is not None:
            p["material"] = material
        if mode is not None:
            p["mode"] = mode
        return p

    def mesh(self, primitives, name=None) -> int:
        m = {"primitives": primitives}
        if name:
            m["name"] = name
        self.doc["meshes"].append(m)
        return len(self.doc["meshes"]) - 1

    def material(self, base_color=(1, 1, 1, 1), roughness=None, metallic=None, emissive=None, double_sided=None, base_color_tex=None,
                 normal_tex=None, normal_scale=None, name=None) -> int:
        pbr = {"baseColorFactor": [float(x) for x in base_color]}
        if roughness is not None:
            pbr["roughnessFactor"] = float(roughness)
        if metallic is not None:
            pbr["metallicFactor"] = float(metallic)
        if base_color_tex is not None:
            pbr["baseColorTexture"] = {"index": base_color_tex}
        m = {"pbrMetallicRoughness": pbr}
        if emissive is not None:
            m["emissiveFactor"] = [float(x) for x in emissive]
        if double_sided is not None:
            m["doubleSided"] = bool(double_sided)
        if normal_tex is not None:
            m["normalTexture"] = {"index": normal_tex}
            if normal_scale is not None:
                m["normalTexture"]["scale"] = float(normal_scale)
        if name:
            m["name"] = name
        self.doc["materials"].append(m)
        return len(self.doc["materials"]) - 1

    def texture(self, png: bytes, embed="view", sampler=None) -> int:
        self.doc.setdefault("images", []); self.doc.setdefault("textures", [])
        if embed == "view":
            self.doc["images"].append({"bufferView": self._view(png), "mimeType": "image/png"})
        else:
            self.doc["images"].append({"uri": "data:image/png;base64," + base64.b64encode(png).decode()})
        t = {"source": len(self.doc["images"]) - 1}
        if sampler is not None:
            self.doc.setdefault("samplers", []).append(sampler)
            t["sampler"] = len(self.doc["samplers"]) - 1
        self.doc["textures"].append(t)
        return len(self.doc["textures"]) - 1

    def node(self, mesh=None, matrix=None, translation=None, rotation=None, scale=None, children=None, name=None, root=True) -> int:
        n = {}
        if mesh is not None:
            n["mesh"] = mesh
        if matrix is not None:
            n["matrix"] = [float(x) for x in np.asarray(matrix, np.float64).reshape(4, 4).T.reshape(-1)]      # column-major
        if translation is not None:
            n["translation"] = [float(x) for x in translation]
        if rotation is not None:
            n["rotation"] = [float(x) for x in rotation]
        if scale is not None:
            n["scale"] = [float(x) for x in scale]
        if children:
            n["children"] = list(children)
        if name:
            n["name"] = name
        self.doc["nodes"].append(n)
        if root:
            self.doc["scenes"][0]["nodes"].append(len(self.doc["nodes"]) - 1)
        return len(self.doc["nodes"]) - 1

    def write(self, path: str, container="gltf+bin"):
        doc = json.loads(json.dumps(self.doc))
        blob = bytes(self.blob)
        if container == "glb":
            doc["buffers"] = [{"byteLength": len(blob)}]
            js = json.dumps(doc).encode()
            js += b" " * (-len(js) % 4)
            bin_ = blob + b"\0" * (-len(blob) % 4)
            body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(bin_), 0x004E4942) + bin_
            with open(path, "wb") as f:
                f.write(b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body)
            return
        if container == "gltf+bin":
            name = path.rsplit("/", 1)[-1].rsplit(".", 1)[0] + " data.bin"                                  # a space: the URI is percent-encoded
            with open(path.rsplit("/", 1)[0] + "/" + name, "wb") as f:
                f.write(blob)
            doc["buffers"] = [{"byteLength": len(blob), "uri": name.replace(" ", "%20")}]
        else:                                                                                                # "gltf+data"
            doc["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]
        with open(path, "w") as f:
            json.dump(doc, f, indent=1)


def from_scene(scene, index_dtype=np.uint32, strided=False) -> GltfBuilder:
    """A scenes.SceneData (single-submesh meshes, glTF-template materials) as a glTF document: one glTF mesh per BLAS, one node per instance
    carrying the instance's 3x4 transform as a `matrix`, one material per bpt_material."""
    b = GltfBuilder()
    for m in scene.materials:
        b.material(base_color=m["base_color"], roughness=m["roughness"], metallic=m["metallic"], emissive=m["emission"], double_sided=bool(int(m["flags"]) & 1))
    mesh_of_blas = {}
    for inst in scene.instances:
        k = int(inst["blas"])
        dr = scene.drawables[int(inst["instance_id_and_mask"]) & 0xFFFFFF]
        mat = int(dr["material_offset"]) // scene.materials.dtype.itemsize
        if k not in mesh_of_blas:
            bd = scene.blas[k]
            idx = scene.indices[int(bd["index_offset"]): int(bd["index_offset"]) + 3 * int(bd["num_triangles"])]
            nv = int(idx.max()) + 1
            v0 = int(bd["position_offset"]) // 3
            mesh_of_blas[k] = (b.primitive(scene.positions.reshape(-1, 3)[v0:v0 + nv], scene.normals.reshape(-1, 3)[v0:v0 + nv],
                                           scene.texcoords.reshape(-1, 2)[v0:v0 + nv], idx, None, index_dtype=index_dtype, strided=strided), {})
        prim, per_mat = mesh_of_blas[k]
        if mat not in per_mat:                                               # the same geometry with another material = another glTF mesh
            per_mat[mat] = b.mesh([dict(prim, material=mat)])
        xf = np.eye(4); xf[:3, :] = np.asarray(inst["transform"], np.float64)
        b.node(mesh=per_mat[mat], matrix=xf)
    return b
