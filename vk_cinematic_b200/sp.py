"""Python mirror of the reference's sp_ interface over libspb200.so's C ABI (include/sp_b200.h).

Thin ctypes bindings: the structures are the C PODs, the functions are the exported symbols with
the reference's names and argument order (reference src/sp_scene.cpp, sp_material_system.cpp,
simd_path_tracer.cpp, tile.h, work_queue.h).  There is no Python or CPU implementation of any
compute entry point here: if the shared library is missing or a symbol is absent, importing
this module raises.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# (development: SPB_B200_LIB points at another build of the same library, e.g. an A/B variant made
# by tools/build_variants.sh; the product path is always the in-tree libspb200.so)
LIB_PATH = os.environ.get("SPB_B200_LIB") or os.path.join(_HERE, "libspb200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "sp_b200.h")

U32_MAX = 0xFFFFFFFF

u32 = C.c_uint32
i32 = C.c_int32
f32 = C.c_float
u64 = C.c_uint64


class vec2(C.Structure):
    _fields_ = [("x", f32), ("y", f32)]


class vec3(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("z", f32)]

    def tuple(self):
        return (self.x, self.y, self.z)


class vec4(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("z", f32), ("w", f32)]

    def tuple(self):
        return (self.x, self.y, self.z, self.w)


quat = vec4


class mat4(C.Structure):
    _fields_ = [("columns", vec4 * 4)]


class MemoryArena(C.Structure):
    _fields_ = [("base", C.c_void_p), ("size", u64), ("capacity", u64)]


class MemoryPool(C.Structure):
    _fields_ = [("storage", C.c_void_p), ("objectSize", u32), ("capacity", u32), ("headIndex", u32)]


class bvh_Tree(C.Structure):
    _fields_ = [("root", C.c_void_p), ("memoryPool", MemoryPool)]


class Aabb(C.Structure):
    _fields_ = [("min", vec3), ("max", vec3)]


class VertexPNT(C.Structure):
    _fields_ = [("position", vec3), ("normal", vec3), ("textureCoord", vec2)]


class HdrImage(C.Structure):
    _fields_ = [("pixels", C.POINTER(f32)), ("width", u32), ("height", u32)]


class Tile(C.Structure):
    _fields_ = [("minX", u32), ("minY", u32), ("maxX", u32), ("maxY", u32)]


class RandomNumberGenerator(C.Structure):
    _fields_ = [("state", u32)]


class WorkQueue(C.Structure):
    _fields_ = [("head", i32), ("tail", i32), ("objectSize", u32), ("maxObjects", u32),
                ("buffer", C.c_void_p)]


SP_MAX_METRICS = 12
(sp_Metric_CyclesElapsed, sp_Metric_PathsTraced, sp_Metric_RaysTraced, sp_Metric_RayHitCount,
 sp_Metric_RayMissCount, sp_Metric_CyclesElapsed_RayIntersectScene,
 sp_Metric_CyclesElapsed_RayIntersectBroadphase, sp_Metric_CyclesElapsed_RayIntersectMesh,
 sp_Metric_CyclesElapsed_RayIntersectMeshMidphase, sp_Metric_CyclesElapsed_RayIntersectTriangle,
 sp_Metric_RayIntersectMesh_MidphaseAabbTestCount,
 sp_Metric_RayIntersectMesh_TestsPerformed) = range(SP_MAX_METRICS)


class sp_b200_BuildInfo(C.Structure):
    _fields_ = [("builder", u32), ("fellBack", u32), ("triangleCount", u32), ("nodeCount", u32),
                ("maxDepth", u32), ("stackNeed", u32), ("deviceMs", f32), ("wallMs", f32)]


BUILDER_HOST_SAH, BUILDER_DEVICE_LBVH = 0, 1
TRIANGLE_MOLLER_TRUMBORE, TRIANGLE_WATERTIGHT = 0, 1


def last_build_info():
    info = sp_b200_BuildInfo()
    lib.sp_b200_GetLastBuildInfo(C.addressof(info))
    return info


class sp_Task(C.Structure):
    """main.cpp:246-250"""
    _fields_ = [("context", C.c_void_p), ("tile", Tile)]


class sp_Metrics(C.Structure):
    _fields_ = [("values", u64 * SP_MAX_METRICS)]


class sp_Mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(VertexPNT)), ("indices", C.POINTER(u32)),
                ("vertexCount", u32), ("indexCount", u32), ("midphaseTree", bvh_Tree),
                ("useSmoothShading", u32)]


SP_SCENE_MAX_OBJECTS = 32


class sp_Scene(C.Structure):
    _fields_ = [("aabbMin", vec3 * SP_SCENE_MAX_OBJECTS), ("aabbMax", vec3 * SP_SCENE_MAX_OBJECTS),
                ("meshes", sp_Mesh * SP_SCENE_MAX_OBJECTS), ("materials", u32 * SP_SCENE_MAX_OBJECTS),
                ("invModelMatrices", mat4 * SP_SCENE_MAX_OBJECTS),
                ("modelMatrices", mat4 * SP_SCENE_MAX_OBJECTS), ("objectCount", u32),
                ("memoryArena", MemoryArena), ("broadphaseTree", bvh_Tree)]


class RayIntersectTriangleResult(C.Structure):
    _fields_ = [("t", f32), ("uv", vec2), ("normal", vec3)]


class sp_RayIntersectMeshResult(C.Structure):
    _fields_ = [("triangleIntersection", RayIntersectTriangleResult)]


class sp_RayIntersectSceneResult(C.Structure):
    _fields_ = [("t", f32), ("materialId", u32), ("normal", vec3), ("uv", vec2)]


class sp_Material(C.Structure):
    _fields_ = [("albedo", vec3), ("albedoTexture", u32), ("emission", vec3),
                ("emissionTexture", u32), ("roughness", f32)]


class sp_MaterialOutput(C.Structure):
    _fields_ = [("albedo", vec3), ("emission", vec3), ("roughness", f32)]


class sp_PathVertex(C.Structure):
    _fields_ = [("materialId", u32), ("worldPosition", vec3), ("outgoingDir", vec3),
                ("incomingDir", vec3), ("normal", vec3), ("uv", vec2)]


SP_MAX_MATERIALS = 32
SP_MAX_IMAGES = 16


class sp_MaterialSystem(C.Structure):
    _fields_ = [("keys", u32 * SP_MAX_MATERIALS), ("materials", sp_Material * SP_MAX_MATERIALS),
                ("count", u32), ("imageKeys", u32 * SP_MAX_IMAGES),
                ("images", HdrImage * SP_MAX_IMAGES), ("imageCount", u32),
                ("backgroundMaterialId", u32)]


class ImagePlane(C.Structure):
    _fields_ = [("pixels", C.POINTER(vec4)), ("width", u32), ("height", u32)]


class Basis(C.Structure):
    _fields_ = [("right", vec3), ("up", vec3), ("forward", vec3)]


class sp_Camera(C.Structure):
    _fields_ = [("basis", Basis), ("position", vec3), ("filmCenter", vec3),
                ("imagePlane", C.POINTER(ImagePlane)), ("halfPixelWidth", f32),
                ("halfPixelHeight", f32), ("halfFilmWidth", f32), ("halfFilmHeight", f32)]


class sp_Context(C.Structure):
    _fields_ = [("camera", C.POINTER(sp_Camera)), ("scene", C.POINTER(sp_Scene)),
                ("materialSystem", C.POINTER(sp_MaterialSystem))]


SP_B200_ENV_NEAREST, SP_B200_ENV_BILINEAR = 0, 1
SP_B200_MATH_F64_ROUNDED, SP_B200_MATH_FAST_F32 = 0, 1
SP_B200_RENDER_WAVEFRONT, SP_B200_RENDER_PER_PIXEL = 0, 1


class sp_b200_Params(C.Structure):
    _fields_ = [("samplesPerPixel", u32), ("bounceCount", u32), ("radianceClamp", f32),
                ("envFilter", u32), ("mathMode", u32), ("cullByDistance", u32),
                ("tileWidth", u32), ("tileHeight", u32), ("renderMode", u32),
                ("samplesPerPass", u32), ("triangleTest", u32)]


class sp_b200_Stats(C.Structure):
    _fields_ = [("rays", u64), ("nodeVisits", u64), ("triangleTests", u64), ("objectTests", u64),
                ("envClampedLookups", u64), ("kernelMs", f32), ("totalMs", f32), ("traceMs", f32),
                ("traceLaunches", u32), ("tracedRays", u64)]


class sp_b200_MeshData(C.Structure):
    _fields_ = [("vertices", C.POINTER(VertexPNT)), ("indices", C.POINTER(u32)), ("vertexCount", u32),
                ("indexCount", u32)]


class sp_b200_TreeInfo(C.Structure):
    _fields_ = [("leafCount", u32), ("nodeCount", u32), ("maxDepth", u32),
                ("allLeavesReachable", u32), ("parentsContainChildren", u32), ("rootMin", vec3),
                ("rootMax", vec3)]


# sizes probed on the reference; the header asserts the same on the C side
_EXPECTED_SIZES = {vec3: 12, vec4: 16, mat4: 64, VertexPNT: 32, sp_Mesh: 64, sp_Scene: 7104,
                   sp_Material: 36, sp_PathVertex: 60, sp_MaterialSystem: 1616, HdrImage: 16,
                   sp_Camera: 88, Tile: 16, sp_Metrics: 96, sp_RayIntersectSceneResult: 28}
for _t, _n in _EXPECTED_SIZES.items():
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)

_P = C.POINTER

_SIGNATURES = {
    "sp_InitializeScene": (None, [_P(sp_Scene), _P(MemoryArena)]),
    "sp_CreateMesh": (sp_Mesh, [_P(VertexPNT), u32, _P(u32), u32, u32]),
    "sp_BuildMeshMidphase": (None, [_P(sp_Mesh), _P(MemoryArena), _P(MemoryArena)]),
    "sp_AddObjectToScene": (None, [_P(sp_Scene), sp_Mesh, u32, vec3, quat, vec3]),
    "sp_b200_AddObjectToScene": (u32, [_P(sp_Scene), sp_Mesh, u32, vec3, quat, vec3]),
    "sp_b200_SceneObjectCount": (u32, [_P(sp_Scene)]),
    "sp_BuildSceneBroadphase": (None, [_P(sp_Scene)]),
    "sp_RayIntersectScene": (sp_RayIntersectSceneResult, [_P(sp_Scene), vec3, vec3, _P(sp_Metrics)]),
    "sp_RayIntersectMesh": (sp_RayIntersectMeshResult, [sp_Mesh, vec3, vec3, _P(sp_Metrics)]),
    "sp_RegisterMaterial": (u32, [_P(sp_MaterialSystem), sp_Material, u32]),
    "sp_FindMaterialById": (_P(sp_Material), [_P(sp_MaterialSystem), u32]),
    "sp_FindTexture": (_P(HdrImage), [_P(sp_MaterialSystem), u32]),
    "sp_RegisterTexture": (u32, [_P(sp_MaterialSystem), HdrImage, u32]),
    "sp_EvaluateMaterial": (sp_MaterialOutput, [_P(sp_MaterialSystem), _P(sp_Material), _P(sp_PathVertex)]),
    "sp_ConfigureCamera": (None, [_P(sp_Camera), _P(ImagePlane), vec3, quat, f32]),
    "sp_CalculateFilmPositions": (u32, [_P(sp_Camera), _P(vec3), _P(vec2), u32]),
    "ComputeRadianceForPath": (vec3, [_P(sp_PathVertex), u32, _P(sp_MaterialSystem)]),
    "sp_PathTraceTile": (None, [_P(sp_Context), Tile, _P(RandomNumberGenerator), _P(sp_Metrics)]),
    "TransformAabb": (Aabb, [vec3, vec3, vec3, quat, vec3]),
    "ComputeTiles": (u32, [u32, u32, u32, u32, _P(Tile), u32]),
    "CreateWorkQueue": (WorkQueue, [_P(MemoryArena), u32, u32]),
    "WorkQueuePush": (u32, [_P(WorkQueue), C.c_void_p, u32]),
    "WorkQueuePop": (C.c_void_p, [_P(WorkQueue), u32]),
    "sp_b200_AddRayTracingWorkQueue": (u32, [_P(WorkQueue), C.c_void_p]),
    "sp_b200_DrainRayTracingWorkQueue": (u32, [_P(WorkQueue), C.c_void_p, u32]),
    "sp_b200_Init": (C.c_int, [C.c_int]),
    "sp_b200_InitDevices": (C.c_int, [u32]),
    "sp_b200_InitDeviceList": (C.c_int, [C.c_void_p, u32]),
    "sp_b200_DeviceCount": (u32, []),
    "sp_b200_RenderFrameToDevice": (C.c_int, [C.c_void_p, u32, C.c_void_p, C.c_void_p]),
    "sp_b200_GetDeviceStats": (C.c_int, [u32, C.c_void_p, _P(u32), _P(u32)]),
    "sp_b200_PartitionRows": (None, [u32, u32, u32, C.c_void_p, C.c_void_p]),
    "sp_b200_RowSeconds": (None, [u32, u32, u32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sp_b200_Shutdown": (None, []),
    "sp_b200_SetLogCallback": (None, [C.c_void_p]),
    "sp_b200_SetStream": (None, [C.c_void_p]),
    "sp_b200_DefaultParams": (None, [_P(sp_b200_Params)]),
    "sp_b200_SetParams": (None, [_P(sp_b200_Params)]),
    "sp_b200_GetParams": (None, [_P(sp_b200_Params)]),
    "sp_b200_GetLastStats": (None, [_P(sp_b200_Stats)]),
    "sp_b200_EnableStats": (None, [C.c_int]),
    "sp_b200_KernelLaunchCount": (u64, []),
    "sp_b200_FlushTextureCache": (None, []),
    "sp_b200_SetPathsPerPass": (None, [u32]),
    "sp_b200_SetSkyCulling": (None, [C.c_int]),
    "sp_b200_SetMissFusion": (None, [C.c_int]),
    "sp_b200_SetRaySorting": (None, [C.c_int]),
    "sp_b200_SetPrimaryCandidates": (None, [C.c_int]),
    "sp_b200_SetCopyOverlap": (None, [C.c_int]),
    "sp_b200_AccumulateFrame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u32, u32, C.c_void_p]),
    "sp_b200_XorShift32Stream": (None, [_P(u32), u32, C.c_void_p]),
    "sp_b200_MeshIntersectedLeavesBatch": (C.c_int, [sp_Mesh, u32, C.c_void_p, C.c_void_p, C.c_void_p, _P(C.c_float)]),
    "sp_b200_RayIntersectMeshBatch": (C.c_int, [sp_Mesh, u32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _P(C.c_float)]),
    "sp_b200_RayIntersectAabb4Batch": (C.c_int, [u32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sp_b200_TileCombinerStats": (None, [_P(C.c_uint64), _P(C.c_uint64)]),
    "sp_b200_SetDeviceTexture": (None, [C.c_void_p, C.c_void_p, u32, u32, C.c_void_p]),
    "sp_b200_SetRefillThresholds": (None, [u32, u32, u32]),
    "sp_b200_SetStragglerEviction": (None, [u32, u32]),
    "sp_b200_Seed": (u32, [u32, u32, u32]),
    "sp_b200_RenderRows": (C.c_int, [_P(sp_Context), u32, u32, u32, C.c_void_p, C.c_void_p,
                                     _P(sp_Metrics), C.c_void_p]),
    "sp_b200_RenderRowsBegin": (C.c_int, [_P(sp_Context), u32, u32, u32, C.c_void_p, C.c_void_p, C.c_int]),
    "sp_b200_RenderRowsEnd": (C.c_int, [C.c_int, _P(sp_Metrics), C.c_void_p]),
    "sp_b200_RenderFrame": (C.c_int, [_P(sp_Context), u32, _P(sp_Metrics)]),
    "LoadExrImage": (C.c_int, [_P(HdrImage), C.c_char_p]),
    "sp_b200_LoadObj": (C.c_int, [C.c_char_p, _P(sp_b200_MeshData)]),
    "sp_b200_FreeMeshData": (None, [_P(sp_b200_MeshData)]),
    "sp_b200_ToneMap": (C.c_int, [C.c_void_p, C.c_void_p, u32, f32, C.c_void_p, C.c_void_p]),
    "sp_b200_SaveExrImage": (C.c_int, [C.c_void_p, C.c_char_p, u32, u32]),
    "sp_b200_SavePpm": (C.c_int, [C.c_void_p, u32, u32, C.c_char_p]),
    "sp_b200_SaveExrImageTiled": (C.c_int, [C.c_void_p, C.c_char_p, u32, u32, u32, u32]),
    "sp_b200_CreateCubeMap": (C.c_int, [C.c_void_p, u32, u32, C.c_void_p, C.c_void_p]),
    "sp_b200_CreateIrradianceCubeMap": (C.c_int, [C.c_void_p, u32, u32, u32, u32, f32, C.c_void_p,
                                                  C.c_void_p]),
    "sp_b200_PrimaryHits": (C.c_int, [_P(sp_Context), u32, u32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sp_b200_RayIntersectSceneBatch": (C.c_int, [_P(sp_Scene), u32, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, _P(sp_Metrics)]),
    "sp_b200_MeshIntersectedLeaves": (u32, [sp_Mesh, vec3, vec3, C.c_void_p, u32, _P(u32)]),
    "sp_b200_MeshTreeInfo": (None, [sp_Mesh, _P(sp_b200_TreeInfo)]),
    "sp_b200_SetMeshBuilder": (None, [u32]),
    "sp_b200_GetLastBuildInfo": (None, [C.c_void_p]),
    "sp_b200_SceneDeviceBytes": (u64, [_P(sp_Scene)]),
    "sp_b200_ReleaseMesh": (None, [_P(sp_Mesh)]),
    "sp_b200_ReleaseScene": (None, [_P(sp_Scene)]),
}


def declared_symbols(header_path=HEADER_PATH):
    """Names of every function include/sp_b200.h declares."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b((?:sp_|sp_b200_)\w+|ComputeRadianceForPath|TransformAabb|ComputeTiles|"
                       r"CreateWorkQueue|WorkQueuePush|WorkQueuePop)\s*\(", text)
    skip = {"sp_b200_LogFn"}
    out = []
    for n in names:
        if n not in out and n not in skip and not n.startswith("sp_Metric"):
            out.append(n)
    return out


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python __graft_entry__.py` (nvcc, sm_100a). "
        "vk_cinematic_b200 has no CPU or pure-Python path.")

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError if the library does not export it
    _fn.restype = _res
    _fn.argtypes = _args
_missing = [n for n in declared_symbols() if not hasattr(lib, n)]
if _missing:
    raise ImportError(f"libspb200.so does not export: {_missing}")
_unbound = [n for n in declared_symbols() if n not in _SIGNATURES]
if _unbound:
    raise ImportError(f"sp.py has no binding for: {_unbound}")


# ------------------------------------------------------------------------------------------------
# small conveniences for host code (no arithmetic of the path lives here)

def V3(v):
    return vec3(float(v[0]), float(v[1]), float(v[2]))


def Q(v):
    return quat(float(v[0]), float(v[1]), float(v[2]), float(v[3]))


def default_params():
    p = sp_b200_Params()
    lib.sp_b200_DefaultParams(C.byref(p))
    return p


def set_params(**kw):
    p = sp_b200_Params()
    lib.sp_b200_GetParams(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    lib.sp_b200_SetParams(C.byref(p))
    return p


def load_obj(path):
    """Wavefront OBJ -> (vertices (n, 8) float32: position, normal, uv; indices uint32) through
    sp_b200_LoadObj.  Returns None if the file is missing or malformed."""
    md = sp_b200_MeshData()
    if lib.sp_b200_LoadObj(os.fsencode(path), C.byref(md)) != 1:
        return None
    vertices = np.ctypeslib.as_array(C.cast(md.vertices, _P(f32)), shape=(md.vertexCount, 8)).copy()
    indices = np.ctypeslib.as_array(md.indices, shape=(md.indexCount,)).copy()
    lib.sp_b200_FreeMeshData(C.byref(md))
    return vertices, indices


def load_exr(path):
    """OpenEXR file -> (H, W, 4) float32 through the library's LoadExrImage; None on failure."""
    img = HdrImage()
    if lib.LoadExrImage(C.byref(img), os.fsencode(path)) != 0:
        return None
    out = np.ctypeslib.as_array(img.pixels, shape=(img.height, img.width, 4)).copy()
    _libc_free(img.pixels)
    return out


def _libc_free(ptr):
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(ptr, C.c_void_p))


def tone_map(rgba, exposure=1.0):
    """(..., 4) float32 host array -> (...,) uint32 RGBA8 through sp_b200_ToneMap."""
    img = np.ascontiguousarray(rgba, dtype=np.float32)
    out = np.zeros(img.size // 4, np.uint32)
    if lib.sp_b200_ToneMap(img.ctypes.data, None, out.size, exposure, out.ctypes.data, None) != 0:
        raise RuntimeError("sp_b200_ToneMap failed")
    return out.reshape(img.shape[:-1])


IRRADIANCE_UNIFORM, IRRADIANCE_RANDOM = 0, 1


def _equirect(env):
    img = np.ascontiguousarray(env, dtype=np.float32)
    assert img.ndim == 3 and img.shape[2] == 4
    hdr = HdrImage(img.ctypes.data_as(C.POINTER(C.c_float)), img.shape[1], img.shape[0])
    return img, hdr


def create_cube_map(env, face_w, face_h):
    """CreateCubeMap (cubemap.cpp:237-291) through sp_b200_CreateCubeMap: (H, W, 4) float32
    equirectangular map -> (6, face_h, face_w, 4) faces (+X -X +Y -Y +Z -Z)."""
    img, hdr = _equirect(env)
    out = np.zeros((6, face_h, face_w, 4), np.float32)
    if lib.sp_b200_CreateCubeMap(C.byref(hdr), face_w, face_h, out.ctypes.data, None) != 0:
        raise RuntimeError("sp_b200_CreateCubeMap failed")
    lib.sp_b200_FlushTextureCache()  # `img` may be a temporary: forget its device copy
    return out


def create_irradiance_cube_map(env, face_w, face_h, spp=32, sampling=IRRADIANCE_UNIFORM, sample_delta=0.1):
    """CreateIrradianceCubeMap (cubemap.cpp:108-233) through sp_b200_CreateIrradianceCubeMap."""
    img, hdr = _equirect(env)
    out = np.zeros((6, face_h, face_w, 4), np.float32)
    if lib.sp_b200_CreateIrradianceCubeMap(C.byref(hdr), face_w, face_h, spp, sampling, sample_delta,
                                           out.ctypes.data, None) != 0:
        raise RuntimeError("sp_b200_CreateIrradianceCubeMap failed")
    lib.sp_b200_FlushTextureCache()
    return out


def write_ppm(path, rgba8):
    """RGBA8 (H, W) uint32 image -> binary PPM through sp_b200_SavePpm."""
    img = np.ascontiguousarray(rgba8, dtype=np.uint32)
    h, w = img.shape
    if lib.sp_b200_SavePpm(img.ctypes.data, w, h, os.fsencode(path)) != 0:
        raise OSError("sp_b200_SavePpm failed: %s" % path)


EXR_HALF, EXR_FLOAT = 1, 2
EXR_NONE, EXR_ZIPS, EXR_ZIP = 0, 2, 3


def save_exr(path, rgba, pixel_type=EXR_FLOAT, compression=EXR_ZIP, tile=None):
    """(H, W, 4) float32 image -> OpenEXR file through sp_b200_SaveExrImage (scanline) or, with
    tile=(w, h), sp_b200_SaveExrImageTiled; False on failure."""
    img, hdr = _equirect(rgba)
    if tile is not None:
        return lib.sp_b200_SaveExrImageTiled(C.byref(hdr), os.fsencode(path), pixel_type, compression,
                                             tile[0], tile[1]) == 0
    return lib.sp_b200_SaveExrImage(C.byref(hdr), os.fsencode(path), pixel_type, compression) == 0


def xorshift_bilateral_stream(seed):
    """draw(n): the next n values of RandomBilateral (math_utils.h:198-214) over one XorShift32 stream."""
    state = u32(seed)

    def draw(n):
        raw = np.zeros(n, np.uint32)
        lib.sp_b200_XorShift32Stream(C.byref(state), n, raw.ctypes.data)
        uni = (raw >> np.uint32(1)).astype(np.float32) / np.float32(2147483648.0)
        return (np.float32(-1.0) + np.float32(2.0) * uni).astype(np.float32)
    return draw


def mesh_leaves_batch(mesh, origins, dirs):
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((len(o), 3), np.uint32)
    ms = C.c_float()
    assert lib.sp_b200_MeshIntersectedLeavesBatch(mesh, len(o), o.ctypes.data, d.ctypes.data, out.ctypes.data, C.byref(ms)) == 0
    return out, ms.value


def mesh_intersect_batch(mesh, origins, dirs):
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    t, tri = np.zeros(len(o), np.float32), np.zeros(len(o), np.int32)
    ms = C.c_float()
    assert lib.sp_b200_RayIntersectMeshBatch(mesh, len(o), o.ctypes.data, d.ctypes.data, t.ctypes.data, tri.ctypes.data, C.byref(ms)) == 0
    return t, tri, ms.value


def last_stats():
    s = sp_b200_Stats()
    lib.sp_b200_GetLastStats(C.byref(s))
    return s


class Renderer:
    """Owns the C structures of one scene + camera + material system and the numpy arrays they
    alias (the reference API aliases caller memory).  Every method is a call into libspb200."""

    def __init__(self, device=0):
        if lib.sp_b200_Init(device) != 0:
            raise RuntimeError("sp_b200_Init failed")
        self.scene = sp_Scene()
        self.materials = sp_MaterialSystem()
        self.camera = sp_Camera()
        self.plane = ImagePlane()
        self.ctx = sp_Context(C.pointer(self.camera), C.pointer(self.scene),
                              C.pointer(self.materials))
        self.meshes = []
        self._keep = []
        self.image = None
        lib.sp_InitializeScene(C.byref(self.scene), None)

    # -- scene -------------------------------------------------------------------------------
    def add_mesh(self, vertices, indices, smooth=False):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 8)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self._keep += [v, i]
        mesh = lib.sp_CreateMesh(v.ctypes.data_as(_P(VertexPNT)), len(v),
                                 i.ctypes.data_as(_P(u32)), len(i), int(bool(smooth)))
        lib.sp_BuildMeshMidphase(C.byref(mesh), None, None)
        self.meshes.append(mesh)
        return len(self.meshes) - 1

    def add_object(self, mesh, material, position=(0, 0, 0), rotation=(0, 0, 0, 1), scale=(1, 1, 1)):
        return lib.sp_b200_AddObjectToScene(C.byref(self.scene), self.meshes[mesh], material, V3(position),
                                            Q(rotation), V3(scale))

    def build(self):
        lib.sp_BuildSceneBroadphase(C.byref(self.scene))

    def register_material(self, id, albedo=(0, 0, 0), albedo_texture=U32_MAX, emission=(0, 0, 0),
                          emission_texture=U32_MAX, roughness=0.0):
        m = sp_Material(V3(albedo), albedo_texture, V3(emission), emission_texture, roughness)
        return bool(lib.sp_RegisterMaterial(C.byref(self.materials), m, id))

    def register_texture(self, id, image):
        img = np.ascontiguousarray(image, dtype=np.float32)
        self._keep.append(img)
        h = HdrImage(img.ctypes.data_as(_P(f32)), img.shape[1], img.shape[0])
        return bool(lib.sp_RegisterTexture(C.byref(self.materials), h, id))

    def set_background(self, material_id):
        self.materials.backgroundMaterialId = material_id

    def configure_camera(self, position, rotation, film_distance, width, height, pixels=None):
        """pixels: optional (height, width, 4) float32 array to render into (e.g. pinned)."""
        if pixels is None:
            pixels = np.zeros((height, width, 4), np.float32)
        assert pixels.shape == (height, width, 4) and pixels.dtype == np.float32
        self.image = pixels
        self.plane.pixels = pixels.ctypes.data_as(_P(vec4))
        self.plane.width = width
        self.plane.height = height
        lib.sp_ConfigureCamera(C.byref(self.camera), C.byref(self.plane), V3(position), Q(rotation),
                               film_distance)

    def load_workload(self, wl, pixels=None):
        for m in wl.meshes:
            self.add_mesh(m.vertices, m.indices, m.smooth)
        for ob in wl.objects:
            self.add_object(ob.mesh, ob.material, ob.position, ob.rotation, ob.scale)
        self.build()
        for tid, img in wl.textures.items():
            self.register_texture(tid, img)
        for m in wl.materials:
            self.register_material(m.id, m.albedo, m.albedo_texture, m.emission,
                                   m.emission_texture, m.roughness)
        self.set_background(wl.background)
        self.configure_camera(wl.camera_position, wl.camera_rotation, wl.film_distance, wl.width,
                              wl.height, pixels)
        return self

    # -- rendering ---------------------------------------------------------------------------
    def render_frame(self, frame=0):
        m = sp_Metrics()
        rc = lib.sp_b200_RenderFrame(C.byref(self.ctx), frame, C.byref(m))
        if rc != 0:
            raise RuntimeError("sp_b200_RenderFrame failed")
        return self.image, np.array(list(m.values), dtype=np.uint64)

    def render_rows(self, row_begin, row_end, frame=0, host=True, device_ptr=None, want_cost=False):
        m = sp_Metrics()
        th = default_tile_height()
        cost = None
        if want_cost:
            rows = (row_end - 1) // th - row_begin // th + 1
            cost = np.zeros(rows, np.uint64)
        host_ptr = self.image.ctypes.data if host else None
        rc = lib.sp_b200_RenderRows(C.byref(self.ctx), row_begin, row_end, frame, host_ptr, device_ptr,
                                    C.byref(m), cost.ctypes.data if cost is not None else None)
        if rc != 0:
            raise RuntimeError("sp_b200_RenderRows failed")
        return np.array(list(m.values), dtype=np.uint64), cost

    def render_rows_begin(self, row_begin, row_end, frame=0, host=True, device_ptr=None, want_cost=False, host_ptr=None):
        """First half of render_rows (sp_b200_RenderRowsBegin): everything is enqueued, nothing waited for.
        Returns the slot to hand to render_rows_end; up to two frames may be in flight."""
        if host_ptr is None:
            host_ptr = self.image.ctypes.data if host else None
        slot = lib.sp_b200_RenderRowsBegin(C.byref(self.ctx), row_begin, row_end, frame, host_ptr, device_ptr,
                                           1 if want_cost else 0)
        if slot < 0:
            raise RuntimeError("sp_b200_RenderRowsBegin failed (two frames in flight already?)")
        if not hasattr(self, "_begun"):
            self._begun = {}
        self._begun[slot] = (row_begin, row_end, want_cost)
        return slot

    def render_rows_end(self, slot):
        """Second half: waits for the frame begun in `slot`; returns (metrics, per-tile-row cost or None)."""
        row_begin, row_end, want_cost = self._begun.pop(slot)
        m = sp_Metrics()
        cost = None
        if want_cost and row_end > row_begin:
            th = default_tile_height()
            cost = np.zeros((row_end - 1) // th - row_begin // th + 1, np.uint64)
        rc = lib.sp_b200_RenderRowsEnd(slot, C.byref(m), cost.ctypes.data if cost is not None else None)
        if rc != 0:
            raise RuntimeError("sp_b200_RenderRowsEnd failed")
        return np.array(list(m.values), dtype=np.uint64), cost

    def path_trace_tile(self, tile, rng_state):
        m = sp_Metrics()
        rng = RandomNumberGenerator(rng_state)
        lib.sp_PathTraceTile(C.byref(self.ctx), Tile(*tile), C.byref(rng), C.byref(m))
        return rng.state, np.array(list(m.values), dtype=np.uint64)

    def render_work_queue(self, capacity=1024):
        """The reference's own frame loop (main.cpp:1380-1381, 1545-1557, 728-759): a WorkQueue of
        `capacity` sp_Tasks (1024 in the app), one task per tile pushed by AddRayTracingWorkQueue,
        drained the way the worker threads drain it.  Returns (tile count, per-tile metrics
        (n, 12) uint64 in pop order); pixels land in the image plane."""
        storage = np.zeros(capacity * C.sizeof(sp_Task), np.uint8)
        arena = MemoryArena(storage.ctypes.data, 0, storage.nbytes)
        queue = lib.CreateWorkQueue(C.byref(arena), C.sizeof(sp_Task), capacity)
        pushed = lib.sp_b200_AddRayTracingWorkQueue(C.byref(queue), C.addressof(self.ctx))
        metrics = (sp_Metrics * max(pushed, 1))()
        done = lib.sp_b200_DrainRayTracingWorkQueue(C.byref(queue), C.addressof(metrics), pushed)
        assert done == pushed and queue.head == queue.tail
        per_tile = np.array([list(metrics[i].values) for i in range(done)], dtype=np.uint64).reshape(done, -1)
        return pushed, per_tile

    def primary_hits(self, sample=0, frame=0):
        n = self.plane.width * self.plane.height
        tri = np.zeros(n, np.int32)
        obj = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        rc = lib.sp_b200_PrimaryHits(C.byref(self.ctx), sample, frame, tri.ctypes.data, obj.ctypes.data,
                                     t.ctypes.data)
        if rc != 0:
            raise RuntimeError("sp_b200_PrimaryHits failed")
        shape = (self.plane.height, self.plane.width)
        return {"tri": tri.reshape(shape), "obj": obj.reshape(shape), "t": t.reshape(shape)}

    def intersect_rays(self, origins, dirs):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        n = len(o)
        res = (sp_RayIntersectSceneResult * n)()
        tri = np.zeros(n, np.int32)
        obj = np.zeros(n, np.int32)
        m = sp_Metrics()
        lib.sp_b200_RayIntersectSceneBatch(C.byref(self.scene), n, o.ctypes.data, d.ctypes.data,
                                           C.addressof(res), tri.ctypes.data, obj.ctypes.data, C.byref(m))
        arr = np.frombuffer(res, dtype=np.float32).reshape(n, 7)
        return {"t": arr[:, 0].copy(), "material": arr[:, 1].copy().view(np.uint32),
                "normal": arr[:, 2:5].copy(), "uv": arr[:, 5:7].copy(), "tri": tri, "obj": obj,
                "metrics": np.array(list(m.values), dtype=np.uint64)}

    def close(self):
        lib.sp_b200_ReleaseScene(C.byref(self.scene))
        for m in self.meshes:
            lib.sp_b200_ReleaseMesh(C.byref(m))
        self.meshes = []


def default_tile_height():
    p = sp_b200_Params()
    lib.sp_b200_GetParams(C.byref(p))
    return p.tileHeight
