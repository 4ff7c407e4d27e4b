// spb_assets.cpp -- host-side asset input for the sp_ path: Wavefront OBJ.
//
// The reference loads meshes through assimp (src/mesh.cpp:5-62: aiProcess_Triangulate |
// aiProcess_JoinIdenticalVertices | ..., first mesh of the file, positions + normals into
// VertexPNT, three indices per face).  assimp is a third-party library that is neither vendored
// in the reference checkout nor installed here, so its output order cannot be reproduced; what
// the path needs from it is a VertexPNT[] / u32[] pair, and this loader defines that pair the way
// tools/make_mesh_fixtures.py (which made assets/*.npz, the meshes every parity test uses) does:
//   * one output vertex per distinct (v, vt, vn) index triple, in first-seen order -- what
//     JoinIdenticalVertices yields for files whose corners share indices;
//   * polygons fan-triangulated in file order, so triangle i is the i-th triangle of the file;
//   * numbers parsed as double and rounded once to float (Python's float() -> numpy float32);
//   * missing vn -> normal (0,0,0); missing vt -> textureCoord (0,0); negative (relative) indices
//     and the four corner syntaxes a, a/t, a//n, a/t/n are accepted; objects/groups are merged.
// No CUDA in this file.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <tuple>
#include <vector>

#include "../../include/sp_b200.h"

namespace {

// parses "a", "a/t", "a//n" or "a/t/n" starting at p; returns the end of the token or null
const char *parse_corner(const char *p, long counts[3], long out[3])
{
    for (int k = 0; k < 3; ++k) out[k] = -1; // -1 = absent
    for (int k = 0; k < 3; ++k)
    {
        if (*p != '/' || k == 0)
        {
            char *end = nullptr;
            long v = strtol(p, &end, 10);
            if (end != p)
            {
                if (v > 0) out[k] = v - 1;
                else if (v < 0) out[k] = counts[k] + v; // relative to the records read so far
                else return nullptr;                     // index 0 is not valid OBJ
                if (out[k] < 0 || out[k] >= counts[k]) return nullptr;
                p = end;
            }
            else if (k == 0) return nullptr;
        }
        if (*p != '/') break;
        ++p; // the separator before field k + 1
    }
    return p;
}

} // namespace

extern "C" int sp_b200_LoadObj(const char *path, sp_b200_MeshData *out)
{
    if (!out) return 0;
    memset(out, 0, sizeof(*out));
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return 0;
    std::vector<float> pos, nrm, tex;
    std::vector<VertexPNT> vertices;
    std::vector<u32> indices;
    std::map<std::tuple<long, long, long>, u32> indexOf;
    std::vector<char> line(1 << 16);
    bool ok = true;
    while (ok && fgets(line.data(), (int)line.size(), f))
    {
        const char *p = line.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t' || ((p[1] == 'n' || p[1] == 't') && (p[2] == ' ' || p[2] == '\t'))))
        {
            std::vector<float> &dst = p[1] == 'n' ? nrm : (p[1] == 't' ? tex : pos);
            const int want = p[1] == 't' ? 2 : 3;
            p += p[1] == ' ' || p[1] == '\t' ? 1 : 2;
            for (int k = 0; k < want; ++k)
            {
                char *end = nullptr;
                double v = strtod(p, &end);
                if (end == p) { if (want == 2 && k == 1) v = 0.0; else { ok = false; break; } }
                dst.push_back((float)v);
                p = end;
            }
        }
        else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t'))
        {
            ++p;
            long counts[3] = {(long)(pos.size() / 3), (long)(tex.size() / 2), (long)(nrm.size() / 3)};
            std::vector<u32> corners;
            for (;;)
            {
                while (*p == ' ' || *p == '\t') ++p;
                if (*p == '\0' || *p == '\n' || *p == '\r' || *p == '#') break;
                long idx[3];
                p = parse_corner(p, counts, idx);
                if (!p) { ok = false; break; }
                auto key = std::make_tuple(idx[0], idx[1], idx[2]);
                auto it = indexOf.find(key);
                if (it == indexOf.end())
                {
                    VertexPNT v;
                    memset(&v, 0, sizeof(v));
                    v.position.x = pos[(size_t)idx[0] * 3 + 0];
                    v.position.y = pos[(size_t)idx[0] * 3 + 1];
                    v.position.z = pos[(size_t)idx[0] * 3 + 2];
                    if (idx[2] >= 0)
                    {
                        v.normal.x = nrm[(size_t)idx[2] * 3 + 0];
                        v.normal.y = nrm[(size_t)idx[2] * 3 + 1];
                        v.normal.z = nrm[(size_t)idx[2] * 3 + 2];
                    }
                    if (idx[1] >= 0)
                    {
                        v.textureCoord.x = tex[(size_t)idx[1] * 2 + 0];
                        v.textureCoord.y = tex[(size_t)idx[1] * 2 + 1];
                    }
                    it = indexOf.emplace(key, (u32)vertices.size()).first;
                    vertices.push_back(v);
                }
                corners.push_back(it->second);
            }
            if (ok && corners.size() < 3) ok = false;
            for (size_t k = 1; ok && k + 1 < corners.size(); ++k)
            {
                indices.push_back(corners[0]);
                indices.push_back(corners[k]);
                indices.push_back(corners[k + 1]);
            }
        }
    }
    fclose(f);
    if (!ok || indices.empty()) return 0;
    // caller frees with sp_b200_FreeMeshData (malloc/free, the convention of LoadExrImage,
    // src/asset_loader/asset_loader.h:11-22)
    out->vertices = (VertexPNT *)malloc(vertices.size() * sizeof(VertexPNT));
    out->indices = (u32 *)malloc(indices.size() * sizeof(u32));
    if (!out->vertices || !out->indices)
    {
        free(out->vertices);
        free(out->indices);
        memset(out, 0, sizeof(*out));
        return 0;
    }
    memcpy(out->vertices, vertices.data(), vertices.size() * sizeof(VertexPNT));
    memcpy(out->indices, indices.data(), indices.size() * sizeof(u32));
    out->vertexCount = (u32)vertices.size();
    out->indexCount = (u32)indices.size();
    return 1;
}

extern "C" void sp_b200_FreeMeshData(sp_b200_MeshData *mesh)
{
    if (!mesh) return;
    free(mesh->vertices);
    free(mesh->indices);
    memset(mesh, 0, sizeof(*mesh));
}
