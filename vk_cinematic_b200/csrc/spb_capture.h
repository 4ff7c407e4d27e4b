// spb_capture.h -- host-side snapshot of the caller's sp_ scene into flat, device-ready arrays.
//
// The reference aliases caller memory (sp_CreateMesh, sp_scene.cpp:8-19) and walks pointers at
// render time; a GPU needs the data resident, so the library snapshots at the two build calls
// (sp_BuildMeshMidphase, sp_BuildSceneBroadphase) and uploads once.  No CUDA in this file.
#pragma once

#include <stdint.h>
#include <memory>
#include <vector>

#include "../../include/sp_b200.h"
#include "spb_bvh.h"
#include "spb_core.cuh"
#include "spb_hostmath.h"

namespace spb {

// What sp_BuildMeshMidphase produces; sp_Mesh::midphaseTree.root points at one of these.
struct MeshAccel
{
    uint32_t magic = 0x4D424853u; // "SHBM"
    std::vector<VertexPNT> vertices;
    std::vector<uint32_t> indices;
    uint32_t triangleCount = 0;
    Bvh4 bvh;
    // device residency of this mesh alone (for sp_RayIntersectMesh), managed by spb_api
    void *deviceScene = nullptr;
};

struct ObjectInstance
{
    std::shared_ptr<MeshAccel> mesh; // may be null: mesh whose midphase was never built
    uint32_t material = 0;
    uint32_t smooth = 0;
    spbh::M4 model, invModel;
    float aabbMin[3], aabbMax[3];
};

// Flat arrays in the layout DScene points into.
struct FlatScene
{
    std::vector<v4f> nodes, tris, shade, objInv, objModel;
    std::vector<v4u> objInfo;
    // per object: first triangle slot of its mesh, then objectCount + 1 prefix sums of the
    // objects' triangle counts (DScene::objTris)
    std::vector<uint32_t> objTris;
    // per object: (world AABB min, extent of its mesh tree) (world AABB max, -)   (DScene::objBox)
    std::vector<v4f> objBox;
    float tlasExtent = 0.0f;     // largest |coordinate| of the object boxes
    uint32_t tlasNodeCount = 0;
    uint64_t instancedTriangles = 0;
    uint32_t tlasRoot = SPB_REF_EMPTY;
    uint32_t objectCount = 0;
    uint64_t triangleCount = 0;
    uint32_t maxDepth = 0;
    uint32_t stackNeed = 0; // worst-case traversal stack entries: TLAS + deepest mesh tree
};

// treeBuilder: how the per-triangle boxes become a Bvh4 (default: build_bvh4, the host SAH builder)
typedef Bvh4 (*TreeBuilderFn)(const float *aabbMin, const float *aabbMax, uint32_t count);
std::shared_ptr<MeshAccel> build_mesh_accel(const VertexPNT *vertices, uint32_t vertexCount,
                                            const uint32_t *indices, uint32_t indexCount,
                                            TreeBuilderFn treeBuilder = nullptr);

// sp_AddObjectToScene's arithmetic (sp_scene.cpp:77-117) for one object.
void compute_object_transform(const MeshAccel *mesh, const VertexPNT *vertices,
                              uint32_t vertexCount, vec3 position, quat orientation, vec3 scale,
                              spbh::M4 *model, spbh::M4 *invModel, float *aabbMin, float *aabbMax);

FlatScene flatten_scene(const std::vector<ObjectInstance> &objects);

// sp_MaterialSystem -> DMaterials with texture ids resolved (first match, like sp_FindTexture).
// imagePixels[i] is what DImage::pixels should be for materialSystem->images[i].
void convert_materials(const sp_MaterialSystem *ms, const v4f *const *imagePixels, DMaterials *out);
void convert_camera(const sp_Camera *camera, DCamera *out);

// sp_ConfigureCamera (simd_path_tracer.cpp:1-36)
void configure_camera(sp_Camera *camera, ImagePlane *imagePlane, vec3 position, quat rotation,
                      float filmDistance);

} // namespace spb
