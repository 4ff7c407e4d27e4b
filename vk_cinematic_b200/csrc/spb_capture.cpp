// spb_capture.cpp -- see spb_capture.h
#include "spb_capture.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#define SPB_CAPTURE_ASSERT(c) do { if (!(c)) { fprintf(stderr, "spb_capture: assertion failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); abort(); } } while (0)

namespace spb {

static inline v4f mk4(float x, float y, float z, float w)
{
    v4f r;
    r.x = x; r.y = y; r.z = z; r.w = w;
    return r;
}

std::shared_ptr<MeshAccel> build_mesh_accel(const VertexPNT *vertices, uint32_t vertexCount,
                                            const uint32_t *indices, uint32_t indexCount,
                                            TreeBuilderFn treeBuilder)
{
    auto accel = std::make_shared<MeshAccel>();
    accel->vertices.assign(vertices, vertices + vertexCount);
    accel->indices.assign(indices, indices + indexCount);
    uint32_t triangleCount = indexCount / 3;
    accel->triangleCount = triangleCount;

    // per-triangle AABB = component-wise min/max of the three positions (sp_scene.cpp:35-50)
    std::vector<float> mn((size_t)triangleCount * 3), mx((size_t)triangleCount * 3);
    for (uint32_t i = 0; i < triangleCount; ++i)
    {
        const vec3 &a = vertices[indices[i * 3 + 0]].position;
        const vec3 &b = vertices[indices[i * 3 + 1]].position;
        const vec3 &c = vertices[indices[i * 3 + 2]].position;
        const float pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z}, pc[3] = {c.x, c.y, c.z};
        for (int k = 0; k < 3; ++k)
        {
            // Min(v0, Min(v1, v2)) / Max(v0, Max(v1, v2)) with the reference's ternaries
            float lo = pb[k] < pc[k] ? pb[k] : pc[k];
            lo = pa[k] < lo ? pa[k] : lo;
            float hi = pb[k] > pc[k] ? pb[k] : pc[k];
            hi = pa[k] > hi ? pa[k] : hi;
            mn[(size_t)i * 3 + k] = lo;
            mx[(size_t)i * 3 + k] = hi;
        }
    }
    accel->bvh = treeBuilder ? treeBuilder(mn.data(), mx.data(), triangleCount)
                             : build_bvh4(mn.data(), mx.data(), triangleCount);
    return accel;
}

void compute_object_transform(const MeshAccel *mesh, const VertexPNT *vertices,
                              uint32_t vertexCount, vec3 position, quat orientation, vec3 scale,
                              spbh::M4 *model, spbh::M4 *invModel, float *aabbMin, float *aabbMax)
{
    (void)mesh;
    // ComputeAabb over every vertex of the mesh (sp_scene.cpp:57-75)
    spbh::V3 lo = spbh::v3(vertices[0].position.x, vertices[0].position.y, vertices[0].position.z);
    spbh::V3 hi = lo;
    for (uint32_t i = 1; i < vertexCount; ++i)
    {
        const vec3 &p = vertices[i].position;
        lo = spbh::v3(spbh::fminr(lo.x, p.x), spbh::fminr(lo.y, p.y), spbh::fminr(lo.z, p.z));
        hi = spbh::v3(spbh::fmaxr(hi.x, p.x), spbh::fmaxr(hi.y, p.y), spbh::fmaxr(hi.z, p.z));
    }
    spbh::V3 pos = spbh::v3(position.x, position.y, position.z);
    spbh::V4 q = {orientation.x, orientation.y, orientation.z, orientation.w};
    spbh::V3 sc = spbh::v3(scale.x, scale.y, scale.z);
    spbh::V3 tmin, tmax;
    spbh::transform_aabb(lo, hi, pos, q, sc, &tmin, &tmax);
    aabbMin[0] = tmin.x; aabbMin[1] = tmin.y; aabbMin[2] = tmin.z;
    aabbMax[0] = tmax.x; aabbMax[1] = tmax.y; aabbMax[2] = tmax.z;
    *model = spbh::model_matrix(pos, q, sc);
    *invModel = spbh::inverse_model_matrix(pos, q, sc);
}

static void append_matrix(std::vector<v4f> &dst, const spbh::M4 &m)
{
    for (int c = 0; c < 4; ++c) dst.push_back(mk4(m.c[c].x, m.c[c].y, m.c[c].z, m.c[c].w));
}

static uint32_t append_tree(FlatScene &fs, const Bvh4 &bvh, uint32_t leafBase)
{
    // returns the global index of the tree's root node, or EMPTY
    if (bvh.nodes.empty()) return SPB_REF_EMPTY;
    uint32_t nodeBase = (uint32_t)(fs.nodes.size() / 8);
    for (const Node4 &n : bvh.nodes)
    {
        for (int a = 0; a < 3; ++a) fs.nodes.push_back(mk4(n.bmin[a][0], n.bmin[a][1], n.bmin[a][2], n.bmin[a][3]));
        for (int a = 0; a < 3; ++a) fs.nodes.push_back(mk4(n.bmax[a][0], n.bmax[a][1], n.bmax[a][2], n.bmax[a][3]));
        uint32_t refs[4];
        for (int k = 0; k < 4; ++k)
        {
            uint32_t r = n.ref[k];
            if (r == SPB_REF_EMPTY) refs[k] = r;
            else if (r & SPB_REF_LEAF) refs[k] = SPB_REF_LEAF | ((r & ~SPB_REF_LEAF) + leafBase);
            else refs[k] = r + nodeBase;
        }
        fs.nodes.push_back(mk4(u2f(refs[0]), u2f(refs[1]), u2f(refs[2]), u2f(refs[3])));
        fs.nodes.push_back(mk4(u2f(n.meta[0]), u2f(n.meta[1]), u2f(n.meta[2]), u2f(n.meta[3])));
    }
    if (bvh.maxDepth > fs.maxDepth) fs.maxDepth = bvh.maxDepth;
    return nodeBase;
}

FlatScene flatten_scene(const std::vector<ObjectInstance> &objects)
{
    FlatScene fs;
    struct Placed { uint32_t root, shadeBase, triBase, triCount; float extent; };
    std::map<const MeshAccel *, Placed> placed;
    uint32_t meshStackNeed = 0;

    for (const ObjectInstance &ob : objects)
    {
        const MeshAccel *mesh = ob.mesh.get();
        if (!mesh || placed.count(mesh)) continue;
        Placed p;
        uint32_t triBase = (uint32_t)(fs.tris.size() / 3);
        p.shadeBase = (uint32_t)(fs.shade.size() / 4);
        p.triBase = triBase;
        p.triCount = (uint32_t)mesh->bvh.slotPrim.size();
        // triangles in leaf-slot order (breadth-first, so neighbours in the tree are neighbours
        // in memory); w of the first vertex carries the triangle's index within its mesh
        for (uint32_t slot = 0; slot < mesh->bvh.slotPrim.size(); ++slot)
        {
            uint32_t tri = mesh->bvh.slotPrim[slot];
            const vec3 &a = mesh->vertices[mesh->indices[tri * 3 + 0]].position;
            const vec3 &b = mesh->vertices[mesh->indices[tri * 3 + 1]].position;
            const vec3 &c = mesh->vertices[mesh->indices[tri * 3 + 2]].position;
            fs.tris.push_back(mk4(a.x, a.y, a.z, u2f(tri)));
            fs.tris.push_back(mk4(b.x, b.y, b.z, 0.0f));
            fs.tris.push_back(mk4(c.x, c.y, c.z, 0.0f));
        }
        // shading attributes by triangle index
        for (uint32_t tri = 0; tri < mesh->triangleCount; ++tri)
        {
            const VertexPNT &v0 = mesh->vertices[mesh->indices[tri * 3 + 0]];
            const VertexPNT &v1 = mesh->vertices[mesh->indices[tri * 3 + 1]];
            const VertexPNT &v2 = mesh->vertices[mesh->indices[tri * 3 + 2]];
            fs.shade.push_back(mk4(v0.normal.x, v0.normal.y, v0.normal.z, v0.textureCoord.x));
            fs.shade.push_back(mk4(v1.normal.x, v1.normal.y, v1.normal.z, v0.textureCoord.y));
            fs.shade.push_back(mk4(v2.normal.x, v2.normal.y, v2.normal.z, v1.textureCoord.x));
            fs.shade.push_back(mk4(v1.textureCoord.y, v2.textureCoord.x, v2.textureCoord.y, 0.0f));
        }
        p.root = append_tree(fs, mesh->bvh, triBase);
        p.extent = bvh4_extent(mesh->bvh);
        SPB_CAPTURE_ASSERT(mesh->bvh.stackNeed <= SPB_MESH_STACK_LIMIT);
        if (mesh->bvh.stackNeed > meshStackNeed) meshStackNeed = mesh->bvh.stackNeed;
        fs.triangleCount += mesh->triangleCount;
        placed[mesh] = p;
    }

    uint32_t count = (uint32_t)objects.size();
    fs.objectCount = count;
    fs.objTris.assign((size_t)count * 2 + 1, 0u);
    for (uint32_t i = 0; i < count; ++i)
    {
        uint32_t n = 0;
        if (objects[i].mesh)
        {
            const Placed &p = placed[objects[i].mesh.get()];
            fs.objTris[i] = p.triBase;
            n = p.triCount;
        }
        SPB_CAPTURE_ASSERT(fs.instancedTriangles + n < 0xFFFFFFFFull);
        fs.objTris[count + i] = (uint32_t)fs.instancedTriangles;
        fs.instancedTriangles += n;
    }
    fs.objTris[(size_t)count * 2] = (uint32_t)fs.instancedTriangles;
    std::vector<float> mn((size_t)count * 3), mx((size_t)count * 3);
    for (uint32_t i = 0; i < count; ++i)
    {
        const ObjectInstance &ob = objects[i];
        append_matrix(fs.objInv, ob.invModel);
        append_matrix(fs.objModel, ob.model);
        v4u info;
        info.x = SPB_REF_EMPTY;
        info.y = 0;
        if (ob.mesh)
        {
            const Placed &p = placed[ob.mesh.get()];
            info.x = p.root;
            info.y = p.shadeBase;
        }
        info.z = ob.smooth;
        info.w = ob.material;
        fs.objInfo.push_back(info);
        for (int k = 0; k < 3; ++k)
        {
            mn[(size_t)i * 3 + k] = ob.aabbMin[k];
            mx[(size_t)i * 3 + k] = ob.aabbMax[k];
            float lo = std::fabs(ob.aabbMin[k]), hi = std::fabs(ob.aabbMax[k]);
            if (lo > fs.tlasExtent) fs.tlasExtent = lo; // (NaN compares false; inf: every ray takes the exact walk)
            if (hi > fs.tlasExtent) fs.tlasExtent = hi;
        }
        fs.objBox.push_back(mk4(ob.aabbMin[0], ob.aabbMin[1], ob.aabbMin[2], ob.mesh ? placed[ob.mesh.get()].extent : 0.0f));
        fs.objBox.push_back(mk4(ob.aabbMax[0], ob.aabbMax[1], ob.aabbMax[2], 0.0f));
    }
    if (count > 0)
    {
        // (a TLAS deeper than its share of the stack -- thousands of objects in a row -- comes back
        // rebuilt with median splits)
        Bvh4 tlas = build_bvh4(mn.data(), mx.data(), count, SPB_TLAS_STACK_LIMIT);
        SPB_CAPTURE_ASSERT(tlas.stackNeed <= SPB_TLAS_STACK_LIMIT);
        // leaf slots of the TLAS refer to object indices
        for (Node4 &n : tlas.nodes)
            for (int k = 0; k < 4; ++k)
                if (n.ref[k] != SPB_REF_EMPTY && (n.ref[k] & SPB_REF_LEAF))
                    n.ref[k] = SPB_REF_LEAF | tlas.slotPrim[n.ref[k] & ~SPB_REF_LEAF];
        fs.tlasRoot = append_tree(fs, tlas, 0);
        fs.tlasNodeCount = (uint32_t)tlas.nodes.size();
        fs.stackNeed = tlas.stackNeed + meshStackNeed;
    }
    // never hand the device a null array
    if (fs.nodes.empty()) fs.nodes.resize(8, mk4(0, 0, 0, 0));
    if (fs.tris.empty()) fs.tris.resize(3, mk4(0, 0, 0, 0));
    if (fs.shade.empty()) fs.shade.resize(4, mk4(0, 0, 0, 0));
    if (fs.objInv.empty()) { fs.objInv.resize(4, mk4(0, 0, 0, 0)); fs.objModel.resize(4, mk4(0, 0, 0, 0)); }
    if (fs.objInfo.empty()) { v4u z; z.x = SPB_REF_EMPTY; z.y = z.z = z.w = 0; fs.objInfo.push_back(z); }
    if (fs.objBox.empty()) fs.objBox.resize(2, mk4(0, 0, 0, 0));
    return fs;
}

void convert_materials(const sp_MaterialSystem *ms, const v4f *const *imagePixels, DMaterials *out)
{
    memset(out, 0, sizeof(*out));
    out->count = ms->count < SPB_MAX_MATERIALS ? ms->count : SPB_MAX_MATERIALS;
    out->backgroundId = ms->backgroundMaterialId;
    out->imageCount = ms->imageCount < SPB_MAX_IMAGES ? ms->imageCount : SPB_MAX_IMAGES;
    auto find_image = [&](uint32_t id) -> int32_t {
        for (uint32_t i = 0; i < out->imageCount; ++i)
            if (ms->imageKeys[i] == id) return (int32_t)i;
        return -1;
    };
    for (uint32_t i = 0; i < out->count; ++i)
    {
        const sp_Material &m = ms->materials[i];
        out->keys[i] = ms->keys[i];
        out->albedo[i][0] = m.albedo.x; out->albedo[i][1] = m.albedo.y; out->albedo[i][2] = m.albedo.z;
        out->emission[i][0] = m.emission.x; out->emission[i][1] = m.emission.y; out->emission[i][2] = m.emission.z;
        out->roughness[i] = m.roughness;
        out->albedoImage[i] = find_image(m.albedoTexture);
        out->emissionImage[i] = find_image(m.emissionTexture);
    }
    // see miss_radiance() in spb_core.cuh; an unknown background id shades as albedo 0, roughness 0
    {
        int slot = -1;
        for (uint32_t i = 0; i < out->count; ++i)
            if (out->keys[i] == out->backgroundId) { slot = (int)i; break; }
        out->simpleBackground = 1;
        if (slot >= 0)
        {
            bool ok = out->albedoImage[slot] < 0 && std::isfinite(out->roughness[slot]) && out->roughness[slot] >= 0.0f;
            for (int k = 0; k < 3; ++k) ok = ok && std::isfinite(out->albedo[slot][k]) && out->albedo[slot][k] >= 0.0f;
            out->simpleBackground = ok ? 1u : 0u;
        }
    }
    for (uint32_t i = 0; i < out->imageCount; ++i)
    {
        out->images[i].pixels = imagePixels[i];
        out->images[i].width = ms->images[i].width;
        out->images[i].height = ms->images[i].height;
    }
}

void convert_camera(const sp_Camera *camera, DCamera *out)
{
    out->right = mk3(camera->basis.right.x, camera->basis.right.y, camera->basis.right.z);
    out->up = mk3(camera->basis.up.x, camera->basis.up.y, camera->basis.up.z);
    out->position = mk3(camera->position.x, camera->position.y, camera->position.z);
    out->filmCenter = mk3(camera->filmCenter.x, camera->filmCenter.y, camera->filmCenter.z);
    out->halfPixelWidth = camera->halfPixelWidth;
    out->halfPixelHeight = camera->halfPixelHeight;
    out->halfFilmWidth = camera->halfFilmWidth;
    out->halfFilmHeight = camera->halfFilmHeight;
    out->width = camera->imagePlane ? camera->imagePlane->width : 0;
    out->height = camera->imagePlane ? camera->imagePlane->height : 0;
}

void configure_camera(sp_Camera *camera, ImagePlane *imagePlane, vec3 position, quat rotation,
                      float filmDistance)
{
    camera->imagePlane = imagePlane;
    camera->position = position;
    spbh::V4 q = {rotation.x, rotation.y, rotation.z, rotation.w};
    spbh::V3 r = spbh::rotate_vector(spbh::v3(1, 0, 0), q);
    spbh::V3 u = spbh::rotate_vector(spbh::v3(0, 1, 0), q);
    spbh::V3 f = spbh::rotate_vector(spbh::v3(0, 0, -1), q);
    camera->basis.right = vec3{r.x, r.y, r.z};
    camera->basis.up = vec3{u.x, u.y, u.z};
    camera->basis.forward = vec3{f.x, f.y, f.z};
    // filmCenter = position + forward * filmDistance
    camera->filmCenter = vec3{position.x + f.x * filmDistance, position.y + f.y * filmDistance,
                              position.z + f.z * filmDistance};
    camera->halfPixelWidth = 0.5f / (float)imagePlane->width;
    camera->halfPixelHeight = 0.5f / (float)imagePlane->height;
    float filmWidth = 1.0f, filmHeight = 1.0f;
    if (imagePlane->width > imagePlane->height)
        filmHeight = (float)imagePlane->height / (float)imagePlane->width;
    else if (imagePlane->width < imagePlane->height)
        filmWidth = (float)imagePlane->width / (float)imagePlane->height;
    camera->halfFilmWidth = 0.5f * filmWidth;
    camera->halfFilmHeight = 0.5f * filmHeight;
}

} // namespace spb
