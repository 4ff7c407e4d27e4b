#!/usr/bin/env python
"""Convert the reference's OBJ inputs into the mesh fixtures the tests and bench.py use.

Run HERE (where /root/reference is mounted); the GPU box only sees the committed outputs:

    python tools/make_mesh_fixtures.py            # writes assets/bunny.npz, assets/monkey.npz

Inputs: /root/reference/assets/bunny.obj (Stanford 3D Scanning Repository, modified) and
monkey.obj (Blender's Suzanne) -- see /root/reference/assets/attributions.md.  These are the
input datasets BASELINE.json's configs name, not reference source code.

Loader rules (ours, applied identically to every implementation because all of them receive the
resulting VertexPNT[] / u32[] arrays -- SURVEY.md §8c "OBJ loading is ours on both sides"):
  * one output vertex per distinct (v, vn) index pair, in first-seen order (what assimp's
    JoinIdenticalVertices does for `f a//n` files, reference src/mesh.cpp:13-16);
  * faces fan-triangulated in file order, so triangle i is the i-th triangle of the file;
  * textureCoord = (0, 0) (neither file has vt records).
"""
import os
import sys

import numpy as np

REF_ASSETS = "/root/reference/assets"
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def load_obj(path):
    positions, normals = [], []
    verts, index_of, indices = [], {}, []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                positions.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("vn "):
                normals.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                corners = []
                for tok in line.split()[1:]:
                    parts = tok.split("/")
                    vi = int(parts[0]) - 1
                    ni = int(parts[2]) - 1 if len(parts) > 2 and parts[2] else -1
                    key = (vi, ni)
                    if key not in index_of:
                        index_of[key] = len(verts)
                        verts.append(key)
                    corners.append(index_of[key])
                for k in range(1, len(corners) - 1):
                    indices.extend([corners[0], corners[k], corners[k + 1]])
    pos = np.asarray(positions, dtype=np.float32)
    nrm = np.asarray(normals, dtype=np.float32)
    vertices = np.zeros((len(verts), 8), dtype=np.float32)
    for i, (vi, ni) in enumerate(verts):
        vertices[i, 0:3] = pos[vi]
        if ni >= 0:
            vertices[i, 3:6] = nrm[ni]
    return vertices, np.asarray(indices, dtype=np.uint32)


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for name in ("bunny", "monkey"):
        src = os.path.join(REF_ASSETS, name + ".obj")
        if not os.path.exists(src):
            print("missing", src, "- run this where the reference is mounted", file=sys.stderr)
            return 1
        vertices, indices = load_obj(src)
        out = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(out, vertices=vertices, indices=indices)
        print(f"{name}: {len(vertices)} vertices, {len(indices) // 3} triangles -> {out} "
              f"({os.path.getsize(out)} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
