/* oracle/ora_api.h -- harness ABI shared by the two CPU checkers.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load these
 * libraries.  The product (vk_cinematic_b200/, include/sp_b200.h) never links or calls them.
 *
 * Two libraries export exactly this interface:
 *   oracle/_ref/libspref.so   -- the reference's OWN sources (/root/reference/src, unmodified,
 *                                unity-included by oracle/ref_driver.cpp) behind this harness.
 *                                3 bounces only (literal at simd_path_tracer.cpp:195).
 *   oracle/libsporacle.so     -- oracle/sp_oracle.cpp, an independent restatement of the same
 *                                algorithm (any bounce count, any object count).
 *
 * All arrays are caller-owned host memory.  float3 = 3 packed floats, quat = (x,y,z,w).
 */
#ifndef ORA_API_H
#define ORA_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_Scene ora_Scene;

/* sp_Metrics slot order (sp_metrics.h:3-41) */
enum {
    ORA_METRIC_CYCLES = 0, ORA_METRIC_PATHS, ORA_METRIC_RAYS, ORA_METRIC_HITS, ORA_METRIC_MISSES,
    ORA_METRIC_CYC_SCENE, ORA_METRIC_CYC_BROADPHASE, ORA_METRIC_CYC_MESH, ORA_METRIC_CYC_MIDPHASE,
    ORA_METRIC_CYC_TRIANGLE, ORA_METRIC_MIDPHASE_AABB_TESTS, ORA_METRIC_MESH_TESTS, ORA_METRIC_COUNT
};

const char *ora_name(void);     /* "reference" or "port" */
uint32_t ora_max_bounces(void); /* 3 for the verbatim reference, 16 for the port */

ora_Scene *ora_create(void);
void ora_destroy(ora_Scene *s);

/* vertices: VertexPNT[vertexCount] = 8 floats each (mesh.h:11-16).  Data is copied.
 * Builds the midphase tree (sp_CreateMesh + sp_BuildMeshMidphase).  Returns mesh index. */
int ora_add_mesh(ora_Scene *s, const float *vertices, uint32_t vertexCount,
                 const uint32_t *indices, uint32_t indexCount, uint32_t smooth);
/* sp_AddObjectToScene.  Returns object index or -1 when the implementation's cap is hit. */
int ora_add_object(ora_Scene *s, uint32_t mesh, uint32_t material, const float *position3,
                   const float *quat4, const float *scale3);
/* sp_BuildSceneBroadphase */
void ora_build(ora_Scene *s);

int ora_register_material(ora_Scene *s, uint32_t id, const float *albedo3, uint32_t albedoTexture,
                          const float *emission3, uint32_t emissionTexture, float roughness);
/* pixels: RGBA f32, row-major.  The checkers copy the image and append width+1 texels equal to
 * the last one (SampleImageNearest is unclamped, image.h:3-18; see ref_driver.cpp). */
int ora_register_texture(ora_Scene *s, uint32_t id, const float *pixels, uint32_t width,
                         uint32_t height);
void ora_set_background(ora_Scene *s, uint32_t materialId);

/* sp_ConfigureCamera on an image plane of width x height */
void ora_configure_camera(ora_Scene *s, const float *position3, const float *quat4,
                          float filmDistance, uint32_t width, uint32_t height);

/* Seed of the XorShift32 stream of one (pixel, sample): see DESIGN.md "RNG". */
uint32_t ora_seed(uint32_t pixelIndex, uint32_t sample, uint32_t frame);

/* Per-(pixel,sample)-seeded render of the rectangle [x0,x1) x [y0,y1) into rgba
 * (full image, width*height*4 floats; pixels outside the rectangle are untouched).
 * Each sample runs the reference integrator on a 1x1 tile with rng.state = ora_seed(...),
 * and the harness accumulates sum_s radiance_s * (1/spp) in sample order in f32.
 * metrics: 12 u64, accumulated over threads (cycle slots are host TSC sums). */
void ora_render_seeded(ora_Scene *s, float *rgba, uint32_t x0, uint32_t y0, uint32_t x1,
                       uint32_t y1, uint32_t spp, uint32_t bounces, uint32_t frame,
                       uint32_t threads, uint64_t *metrics);

/* PORT ONLY.  mask[width*height] (u8): 1 for every pixel of the rectangle on whose paths (same
 * seeds as ora_render_seeded) some sp_RayIntersectScene call ended with two candidates of
 * bit-equal closest t -- the one case where the winner depends on visiting order, i.e. on tree
 * topology, which the reference's algorithm does not fix (bvh.cpp:51-200 vs any other builder).
 * Parity tests accept a differing pixel only where this mask is set. */
void ora_tie_mask(ora_Scene *s, uint8_t *mask, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                  uint32_t spp, uint32_t bounces, uint32_t frame, uint32_t threads);

/* The reference's native scheduling (main.cpp:731-759,819-844): tileW x tileH tiles popped from
 * the work queue by `threads` workers, every tile seeded 0xF51C0E49, spp samples per pixel.
 * Returns wall seconds (steady_clock, submit -> last tile done). */
double ora_render_tiles(ora_Scene *s, float *rgba, uint32_t tileW, uint32_t tileH, uint32_t spp,
                        uint32_t bounces, uint32_t threads, uint64_t *metrics);

/* The same native scheduling over a SUBSET of the frame's tiles (a bounded sample of a large
 * workload): tiles = count x 4 u32 (minX minY maxX maxY, as ComputeTiles produced them), every
 * tile seeded 0xF51C0E49.  Returns wall seconds. */
double ora_render_tile_list(ora_Scene *s, float *rgba, const uint32_t *tiles, uint32_t count,
                            uint32_t spp, uint32_t bounces, uint32_t threads, uint64_t *metrics);

/* One sp_PathTraceTile call with the caller's rng state (in/out). */
void ora_path_trace_tile(ora_Scene *s, float *rgba, uint32_t minX, uint32_t minY, uint32_t maxX,
                         uint32_t maxY, uint32_t spp, uint32_t bounces, uint32_t *rngState,
                         uint64_t *metrics);

/* Primary rays of frame `frame`, sample `sample`, generated exactly as sp_PathTraceTile does.
 * Outputs per pixel: triangle index (-1 on miss), object index, world t (sp_scene.cpp:302).
 * Any output pointer may be NULL. */
void ora_primary_hits(ora_Scene *s, int32_t *triId, int32_t *objId, float *t, float *rayDir3,
                      uint32_t sample, uint32_t frame, uint32_t threads);

/* sp_RayIntersectScene on n rays.  out7 per ray: t, materialId (as float bits), n.xyz, uv.xy.
 * triId/objId as in ora_primary_hits (may be NULL). */
void ora_intersect_rays(ora_Scene *s, uint32_t n, const float *origins3, const float *dirs3,
                        float *out7, int32_t *triId, int32_t *objId, uint64_t *metrics);

/* ---- known-answer entry points (pure functions) ---- */
uint32_t ora_xorshift32(uint32_t *state);
float ora_random_unilateral(uint32_t *state);
float ora_random_bilateral(uint32_t *state);
/* out6: t, u, v, n.xyz (RayIntersectTriangleMT, ray_intersection.cpp:156-190) */
void ora_ray_triangle_mt(const float *o3, const float *d3, const float *a3, const float *b3,
                         const float *c3, float *out6);
/* simd_RayIntersectAabb4 (simd.h:198-271); boxMin/boxMax: 4 x float3; returns 4-bit mask */
uint32_t ora_ray_aabb4(const float *boxMin12, const float *boxMax12, const float *o3,
                       const float *invDir3);
/* scalar RayIntersectAabb (ray_intersection.cpp:24-77); returns t or -1 */
float ora_ray_aabb_scalar(const float *boxMin3, const float *boxMax3, const float *o3,
                          const float *d3);
/* RandomDirectionOnHemisphere (math_lib.h:932-946) */
void ora_hemisphere(uint32_t *state, const float *normal3, float *out3);
void ora_to_spherical(const float *v3, float *out2);     /* math_lib.h:849-860 */
void ora_map_equirect(const float *sphere2, float *out2); /* math_lib.h:873-884 */
void ora_spherical_to_cartesian(const float *sphere2, float *out3); /* math_lib.h:864-871 */
/* sp_ConfigureCamera -> out22: right3 up3 forward3 position3 filmCenter3 halfPixelW halfPixelH
 * halfFilmW halfFilmH (then 3 unused) */
void ora_camera_fields(const float *position3, const float *quat4, float filmDistance,
                       uint32_t width, uint32_t height, float *out22);
/* sp_CalculateFilmPositions for n pixel positions (uses the scene's configured camera) */
void ora_film_positions(ora_Scene *s, uint32_t n, const float *pixelPos2, float *out3);
/* TransformAabb (aabb.h:29-58): out6 = min3,max3 */
void ora_transform_aabb(const float *min3, const float *max3, const float *position3,
                        const float *quat4, const float *scale3, float *out6);
/* ComputeRadianceForPath (simd_path_tracer.cpp:107-175).  path: n x 15 floats
 * (materialId bits, worldPosition3, outgoingDir3, incomingDir3, normal3, uv2) */
void ora_radiance_for_path(ora_Scene *s, const float *path15, uint32_t n, float *out3);
/* SampleImageNearest / SampleImageBilinear (image.h:3-18,34-73) */
void ora_sample_nearest(const float *pixels, uint32_t w, uint32_t h, float u, float v, float *out4);
void ora_sample_bilinear(const float *pixels, uint32_t w, uint32_t h, float u, float v, float *out4);
/* Output stage (port only): PerformToneMapping of src/shaders/post_processing.frag.glsl:19-26 --
 * color *= exposure; color = color / (1 + color); pow(color, 1/2.2) -- on the RGB of `count` RGBA
 * f32 pixels, then the 8-bit UNORM store a colour attachment performs (round(clamp(c,0,1) * 255)),
 * alpha 255; out = count x RGBA8 (r in the low byte, like ToColor, src/math_lib.h:523-532).  The
 * shader is not CPU code, so this function is pinned by known-answer values only. */
void ora_tone_map(const float *rgba, uint32_t count, float exposure, uint32_t *out);
/* Environment pre-processing (src/cubemap.cpp).  pixels = w x h RGBA f32 equirectangular map; out =
 * 6 faces (+X -X +Y -Y +Z -Z) of faceW x faceH RGBA f32, layer-major.
 * ora_create_cube_map: CreateCubeMap (cubemap.cpp:237-291).
 * ora_create_irradiance_cube_map: CreateIrradianceCubeMap (cubemap.cpp:108-233); sampling 0 = the
 * uniform grid the reference's config.h:47 selects, 1 = the random branch.  The reference build
 * compiles both branches (the file is included twice) but its sampleDelta is the literal 0.1f:
 * any other value is refused there (returns 0; 1 on success).  The port takes any step. */
void ora_create_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW, uint32_t faceH,
                         float *out);
int ora_create_irradiance_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW,
                                   uint32_t faceH, uint32_t samplesPerPixel, uint32_t sampling,
                                   float sampleDelta, float *out);
/* ComputeTiles (tile.h:11-42): tiles = maxTiles x 4 u32; returns count */
uint32_t ora_compute_tiles(uint32_t w, uint32_t h, uint32_t tw, uint32_t th, uint32_t *tiles,
                           uint32_t maxTiles);
/* bvh_CreateTree + bvh_IntersectRay over n AABBs (bvh.cpp:51-311).  Collects intersected leaf
 * indices (level order) into leaves[maxLeaves]; returns count, sets *error on overflow,
 * *aabbTests = aabbTestCount; *rootBounds6 = root min/max (may be NULL). */
uint32_t ora_bvh_query(const float *aabbMin, const float *aabbMax, uint32_t n, const float *o3,
                       const float *d3, uint32_t *leaves, uint32_t maxLeaves, uint32_t *error,
                       uint32_t *aabbTests, float *rootBounds6);
/* The reference's performance tests as library calls (perf_tests/perf_tests.cpp:51-118 TestBvh,
 * :212-305 TestMeshMidphase), single-threaded like there.  Both return the wall seconds of the
 * query loop alone (steady_clock).
 *   ora_perf_bvh: one tree over `n` boxes (bvh_CreateTree; *buildSeconds if not NULL), then
 *     bvh_IntersectRay for every ray with a `maxLeaves`-entry result array; countXorSum[3 q + 0] = leaves
 *     reported for ray q, [1] = XOR of leafIndex * 0x9E3779B1, [2] = sum of the leaf indices.
 *   ora_perf_mesh: sp_RayIntersectMesh(mesh `mesh` of the scene) for every object-space ray; t[q]
 *     (-1 = miss), tri[q] (-1 = miss; the port reports the triangle, the reference driver -2 = unknown). */
double ora_perf_bvh(const float *aabbMin, const float *aabbMax, uint32_t n, uint32_t rays, const float *origins3,
                    const float *dirs3, uint32_t maxLeaves, uint32_t *countXorSum, double *buildSeconds);
double ora_perf_mesh(ora_Scene *s, uint32_t mesh, uint32_t rays, const float *origins3, const float *dirs3, float *t,
                     int32_t *tri);
/* Tree statistics of mesh `mesh` midphase: out[0]=leafCount out[1]=internalCount
 * out[2]=minDepth out[3]=maxDepth out[4]=allLeavesReachable out[5]=parentsContainChildren */
void ora_mesh_tree_stats(ora_Scene *s, uint32_t mesh, uint32_t *out6);

#ifdef __cplusplus
}
#endif
#endif
