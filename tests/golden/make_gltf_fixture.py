"""Writes tests/golden/cornell_box.glb (BASELINE configs[0]'s scene as a glTF 2.0 binary, through tests/_gltf_writer.py) and
tests/golden/cornell_box_gltf.npz: what the C++ importer (host/gltf.cpp) made of it — geometry streams incl. the MikkTSpace tangents,
BLAS / drawable / instance / material records — and the oracle's 2-spp image of the imported scene at depth 5.
    python tests/golden/make_gltf_fixture.py
The .glb is a committed INPUT fixture; the .npz pins the importer and the render of its output against regressions."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
GLB, NPZ = os.path.join(HERE, "cornell_box.glb"), os.path.join(HERE, "cornell_box_gltf.npz")
W = H = 48
ARRAYS = ("positions", "normals", "tangents", "texcoords", "indices", "blas", "drawables", "instances", "materials")


def source_scene():
    from bisemutum_engine_b200 import scenes
    return scenes.cornell_box(tess=2)


def imported(path=GLB):
    """(arrays of the import, SceneData with the source scene's camera and light)."""
    from bisemutum_engine_b200 import engine
    p = engine.Project.from_gltf(path)
    arrays = {k: p.array(k) for k in ARRAYS}
    sd = p.scene_data()
    src = source_scene()
    sd.dir_lights, sd.camera = src.dir_lights, src.camera
    p.close()
    return arrays, sd


def render(context_cls, sd, mode):
    from bisemutum_engine_b200 import capi, engine
    ctx = context_cls(W, H); ctx.upload_scene(sd, mode)
    ctx.render(engine.camera_matrices(sd.camera, W, H), 0, 2, capi.Settings(max_bounces=5))
    img = ctx.resolve(2); ctx.close()
    return img


if __name__ == "__main__":
    import _gltf_writer as gw
    from bisemutum_engine_b200 import capi
    from oracle import oracle_py
    oracle_py.build()
    gw.from_scene(source_scene(), index_dtype=np.uint16).write(GLB, "glb")
    arrays, sd = imported()
    out = {k: v.view(np.uint8) if v.dtype.names else v for k, v in arrays.items()}          # records as raw bytes
    out["image"] = render(oracle_py.OracleContext, sd, capi.ACCEL_TWO_LEVEL)
    np.savez_compressed(NPZ, **out)
    print(f"{GLB}: {os.path.getsize(GLB)} B, {NPZ}: {os.path.getsize(NPZ)} B, mean radiance {out['image'][..., :3].mean():.4f}")
