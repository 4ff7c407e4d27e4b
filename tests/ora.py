"""ctypes wrapper over the oracle harness ABI (oracle/ora_api.h).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(_ROOT, "oracle")
PORT_LIB = os.path.join(ORACLE_DIR, "libsporacle.so")
PORT_DM_LIB = os.path.join(ORACLE_DIR, "libsporacle_dm.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libspref.so")
REF_DM_LIB = os.path.join(ORACLE_DIR, "_ref", "libspref_dm.so")

METRIC_NAMES = ["cycles", "paths", "rays", "hits", "misses", "cyc_scene", "cyc_broadphase",
                "cyc_mesh", "cyc_midphase", "cyc_triangle", "midphase_aabb_tests", "mesh_tests"]

_f = C.POINTER(C.c_float)
_u = C.POINTER(C.c_uint32)
_i = C.POINTER(C.c_int32)
_u64 = C.POINTER(C.c_uint64)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f)


def _up(a):
    return None if a is None else a.ctypes.data_as(_u)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracles(verbose=False):
    """Compile oracle/libsporacle.so (always) and oracle/_ref/libspref.so (when the reference
    sources are mounted).  Building the checker is not using it."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=out)
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=out)


def have_ref():
    return os.path.exists(REF_LIB)


def have_port():
    return os.path.exists(PORT_LIB)


class _Missing:
    """Placeholder for an entry point a library does not export (hostsim exports a subset)."""

    def __init__(self, name):
        self.name = name
        self.argtypes = None
        self.restype = None

    def __call__(self, *a):
        raise NotImplementedError(self.name)


class _Tolerant:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_missing", {})

    def __getattr__(self, name):
        try:
            return getattr(self._lib, name)
        except AttributeError:
            return self._missing.setdefault(name, _Missing(name))


class OracleLib:
    def __init__(self, path):
        self.path = path
        lib = _Tolerant(C.CDLL(path))
        self.lib = lib
        lib.ora_name.restype = C.c_char_p
        lib.ora_max_bounces.restype = C.c_uint32
        lib.ora_create.restype = C.c_void_p
        lib.ora_destroy.argtypes = [C.c_void_p]
        lib.ora_add_mesh.argtypes = [C.c_void_p, _f, C.c_uint32, _u, C.c_uint32, C.c_uint32]
        lib.ora_add_object.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _f, _f, _f]
        lib.ora_build.argtypes = [C.c_void_p]
        lib.ora_register_material.argtypes = [C.c_void_p, C.c_uint32, _f, C.c_uint32, _f,
                                              C.c_uint32, C.c_float]
        lib.ora_register_texture.argtypes = [C.c_void_p, C.c_uint32, _f, C.c_uint32, C.c_uint32]
        lib.ora_set_background.argtypes = [C.c_void_p, C.c_uint32]
        lib.ora_configure_camera.argtypes = [C.c_void_p, _f, _f, C.c_float, C.c_uint32, C.c_uint32]
        lib.ora_seed.argtypes = [C.c_uint32] * 3
        lib.ora_seed.restype = C.c_uint32
        lib.ora_render_seeded.argtypes = [C.c_void_p, _f] + [C.c_uint32] * 8 + [_u64]
        lib.ora_tie_mask.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)] + [C.c_uint32] * 8
        lib.ora_render_tiles.argtypes = [C.c_void_p, _f] + [C.c_uint32] * 5 + [_u64]
        lib.ora_render_tiles.restype = C.c_double
        lib.ora_render_tile_list.argtypes = [C.c_void_p, _f, _u] + [C.c_uint32] * 4 + [_u64]
        lib.ora_render_tile_list.restype = C.c_double
        lib.ora_path_trace_tile.argtypes = [C.c_void_p, _f] + [C.c_uint32] * 6 + [_u, _u64]
        lib.ora_primary_hits.argtypes = [C.c_void_p, _i, _i, _f, _f, C.c_uint32, C.c_uint32,
                                         C.c_uint32]
        lib.ora_intersect_rays.argtypes = [C.c_void_p, C.c_uint32, _f, _f, _f, _i, _i, _u64]
        lib.ora_xorshift32.argtypes = [_u]
        lib.ora_xorshift32.restype = C.c_uint32
        lib.ora_random_unilateral.argtypes = [_u]
        lib.ora_random_unilateral.restype = C.c_float
        lib.ora_random_bilateral.argtypes = [_u]
        lib.ora_random_bilateral.restype = C.c_float
        lib.ora_ray_triangle_mt.argtypes = [_f] * 6
        lib.ora_ray_aabb4.argtypes = [_f] * 4
        lib.ora_ray_aabb4.restype = C.c_uint32
        lib.ora_ray_aabb_scalar.argtypes = [_f] * 4
        lib.ora_ray_aabb_scalar.restype = C.c_float
        lib.ora_hemisphere.argtypes = [_u, _f, _f]
        lib.ora_to_spherical.argtypes = [_f, _f]
        lib.ora_map_equirect.argtypes = [_f, _f]
        lib.ora_spherical_to_cartesian.argtypes = [_f, _f]
        lib.ora_camera_fields.argtypes = [_f, _f, C.c_float, C.c_uint32, C.c_uint32, _f]
        lib.ora_film_positions.argtypes = [C.c_void_p, C.c_uint32, _f, _f]
        lib.ora_transform_aabb.argtypes = [_f] * 6
        lib.ora_radiance_for_path.argtypes = [C.c_void_p, _f, C.c_uint32, _f]
        if hasattr(lib, "ora_tone_map"):
            lib.ora_tone_map.argtypes = [_f, C.c_uint32, C.c_float, _u]
        lib.ora_create_cube_map.argtypes = [_f] + [C.c_uint32] * 4 + [_f]
        lib.ora_create_irradiance_cube_map.argtypes = [_f] + [C.c_uint32] * 6 + [C.c_float, _f]
        lib.ora_create_irradiance_cube_map.restype = C.c_int
        lib.hostsim_set_builder.argtypes = [C.c_int]
        lib.hostsim_build_info.argtypes = [_f, C.c_uint32, _u, C.c_uint32, C.c_int, _u]
        lib.hostsim_create_cube_map.argtypes = [_f] + [C.c_uint32] * 4 + [_f]
        lib.hostsim_create_irradiance_cube_map.argtypes = [_f] + [C.c_uint32] * 6 + [C.c_float, _f]
        lib.ora_sample_nearest.argtypes = [_f, C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f]
        lib.ora_sample_bilinear.argtypes = [_f, C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f]
        lib.ora_compute_tiles.argtypes = [C.c_uint32] * 4 + [_u, C.c_uint32]
        lib.ora_compute_tiles.restype = C.c_uint32
        lib.ora_bvh_query.argtypes = [_f, _f, C.c_uint32, _f, _f, _u, C.c_uint32, _u, _u, _f]
        lib.ora_bvh_query.restype = C.c_uint32
        if hasattr(lib, "ora_perf_bvh"):
            lib.ora_perf_bvh.argtypes = [_f, _f, C.c_uint32, C.c_uint32, _f, _f, C.c_uint32, _u, C.POINTER(C.c_double)]
            lib.ora_perf_bvh.restype = C.c_double
            lib.ora_perf_mesh.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _f, _f, _f, C.POINTER(C.c_int32)]
            lib.ora_perf_mesh.restype = C.c_double
        lib.ora_mesh_tree_stats.argtypes = [C.c_void_p, C.c_uint32, _u]
        self.name = lib.ora_name().decode()
        self.max_bounces = lib.ora_max_bounces()

    # ---- pure functions ----
    def seed(self, pixel, sample, frame):
        return self.lib.ora_seed(pixel, sample, frame)

    def xorshift32(self, state):
        s = C.c_uint32(state)
        r = self.lib.ora_xorshift32(C.byref(s))
        return r, s.value

    def random_unilateral(self, state):
        s = C.c_uint32(state)
        r = self.lib.ora_random_unilateral(C.byref(s))
        return np.float32(r), s.value

    def random_bilateral(self, state):
        s = C.c_uint32(state)
        r = self.lib.ora_random_bilateral(C.byref(s))
        return np.float32(r), s.value

    def ray_triangle_mt(self, o, d, a, b, c):
        out = np.zeros(6, np.float32)
        self.lib.ora_ray_triangle_mt(_fp(_f32(o)), _fp(_f32(d)), _fp(_f32(a)), _fp(_f32(b)),
                                     _fp(_f32(c)), _fp(out))
        return out

    def ray_aabb4(self, box_min, box_max, o, inv):
        return self.lib.ora_ray_aabb4(_fp(_f32(box_min).reshape(-1)), _fp(_f32(box_max).reshape(-1)),
                                      _fp(_f32(o)), _fp(_f32(inv)))

    def ray_aabb_scalar(self, mn, mx, o, d):
        return np.float32(self.lib.ora_ray_aabb_scalar(_fp(_f32(mn)), _fp(_f32(mx)), _fp(_f32(o)),
                                                       _fp(_f32(d))))

    # ---- hostsim only: which builder makes the mesh trees ----
    def set_builder(self, builder):
        self.lib.hostsim_set_builder(int(builder))

    def build_info(self, vertices, indices, builder):
        """nodeCount, maxDepth, stackNeed, leafCount, fellBack, area proxy of the tree `builder`
        (0 host SAH, 1 host emulation of the device LBVH) makes for this mesh."""
        v = _f32(vertices)
        i = np.ascontiguousarray(indices, dtype=np.uint32)
        out = np.zeros(8, np.uint32)
        self.lib.hostsim_build_info(_fp(v), len(v), _up(i), len(i), int(builder), _up(out))
        return dict(zip(["nodeCount", "maxDepth", "stackNeed", "leafCount", "fellBack", "area"], (int(x) for x in out[:6])))

    # ---- environment pre-processing (src/cubemap.cpp) ----
    def create_cube_map(self, env, face_w, face_h):
        """env: (H, W, 4) float32 equirect map -> (6, face_h, face_w, 4)"""
        env = _f32(env)
        out = np.zeros((6, face_h, face_w, 4), np.float32)
        fn = self.lib.hostsim_create_cube_map if self.name == "hostsim" else self.lib.ora_create_cube_map
        fn(_fp(env), env.shape[1], env.shape[0], face_w, face_h, _fp(out))
        return out

    def create_irradiance_cube_map(self, env, face_w, face_h, spp=32, sampling=0, sample_delta=0.1):
        env = _f32(env)
        out = np.zeros((6, face_h, face_w, 4), np.float32)
        if self.name == "hostsim":
            self.lib.hostsim_create_irradiance_cube_map(_fp(env), env.shape[1], env.shape[0], face_w, face_h,
                                                        spp, sampling, sample_delta, _fp(out))
        elif not self.lib.ora_create_irradiance_cube_map(_fp(env), env.shape[1], env.shape[0], face_w, face_h,
                                                         spp, sampling, sample_delta, _fp(out)):
            raise ValueError("%s cannot bake with sample_delta %r" % (self.name, sample_delta))
        return out

    def hemisphere(self, state, normal):
        s = C.c_uint32(state)
        out = np.zeros(3, np.float32)
        self.lib.ora_hemisphere(C.byref(s), _fp(_f32(normal)), _fp(out))
        return out, s.value

    def to_spherical(self, v):
        out = np.zeros(2, np.float32)
        self.lib.ora_to_spherical(_fp(_f32(v)), _fp(out))
        return out

    def map_equirect(self, sc):
        out = np.zeros(2, np.float32)
        self.lib.ora_map_equirect(_fp(_f32(sc)), _fp(out))
        return out

    def spherical_to_cartesian(self, sc):
        out = np.zeros(3, np.float32)
        self.lib.ora_spherical_to_cartesian(_fp(_f32(sc)), _fp(out))
        return out

    def camera_fields(self, position, rotation, film_distance, width, height):
        out = np.zeros(22, np.float32)
        self.lib.ora_camera_fields(_fp(_f32(position)), _fp(_f32(rotation)), film_distance, width,
                                   height, _fp(out))
        return {"right": out[0:3], "up": out[3:6], "forward": out[6:9], "position": out[9:12],
                "filmCenter": out[12:15], "halfPixelWidth": out[15], "halfPixelHeight": out[16],
                "halfFilmWidth": out[17], "halfFilmHeight": out[18]}

    def transform_aabb(self, mn, mx, position, rotation, scale):
        out = np.zeros(6, np.float32)
        self.lib.ora_transform_aabb(_fp(_f32(mn)), _fp(_f32(mx)), _fp(_f32(position)),
                                    _fp(_f32(rotation)), _fp(_f32(scale)), _fp(out))
        return out[0:3], out[3:6]

    def sample_nearest(self, image, u, v):
        img = _f32(image)
        out = np.zeros(4, np.float32)
        self.lib.ora_sample_nearest(_fp(img), img.shape[1], img.shape[0], u, v, _fp(out))
        return out

    def tone_map(self, rgba, exposure=1.0):
        """(..., 4) float32 -> (...,) uint32 RGBA8 (port only)."""
        img = np.ascontiguousarray(rgba, dtype=np.float32)
        out = np.zeros(img.size // 4, np.uint32)
        self.lib.ora_tone_map(_fp(img), out.size, exposure, _up(out))
        return out.reshape(img.shape[:-1])

    def sample_bilinear(self, image, u, v):
        img = _f32(image)
        out = np.zeros(4, np.float32)
        self.lib.ora_sample_bilinear(_fp(img), img.shape[1], img.shape[0], u, v, _fp(out))
        return out

    def compute_tiles(self, w, h, tw, th, max_tiles):
        tiles = np.zeros((max_tiles, 4), np.uint32)
        n = self.lib.ora_compute_tiles(w, h, tw, th, _up(tiles), max_tiles)
        return n, tiles

    def bvh_query(self, aabb_min, aabb_max, o, d, max_leaves):
        mn = _f32(aabb_min).reshape(-1, 3)
        mx = _f32(aabb_max).reshape(-1, 3)
        leaves = np.zeros(max(1, max_leaves), np.uint32)
        err, tests = C.c_uint32(0), C.c_uint32(0)
        root = np.zeros(6, np.float32)
        n = self.lib.ora_bvh_query(_fp(mn), _fp(mx), len(mn), _fp(_f32(o)), _fp(_f32(d)),
                                   _up(leaves), max_leaves, C.byref(err), C.byref(tests), _fp(root))
        return {"count": n, "leaves": leaves[:n].copy(), "error": bool(err.value),
                "aabb_tests": tests.value, "root_min": root[0:3], "root_max": root[3:6]}

    def perf_bvh(self, aabb_min, aabb_max, origins, dirs, max_leaves):
        """TestBvh's loop (perf_tests.cpp:101-108): (count, xor, sum) per ray, build seconds, query seconds."""
        mn, mx = _f32(aabb_min).reshape(-1, 3), _f32(aabb_max).reshape(-1, 3)
        o, d = _f32(origins).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        out = np.zeros((len(o), 3), np.uint32)
        build = C.c_double()
        secs = self.lib.ora_perf_bvh(_fp(mn), _fp(mx), len(mn), len(o), _fp(o), _fp(d), max_leaves, _up(out), C.byref(build))
        return out, build.value, secs

    def scene(self):
        return OracleScene(self)


class OracleScene:
    def __init__(self, olib):
        self.o = olib
        self.lib = olib.lib
        self.h = self.lib.ora_create()
        self._keep = []
        self.width = self.height = 0

    def close(self):
        if self.h:
            self.lib.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_mesh(self, vertices, indices, smooth=False):
        v = _f32(vertices)
        i = np.ascontiguousarray(indices, dtype=np.uint32)
        return self.lib.ora_add_mesh(self.h, _fp(v), len(v), _up(i), len(i), int(bool(smooth)))

    def add_object(self, mesh, material, position=(0, 0, 0), rotation=(0, 0, 0, 1),
                   scale=(1, 1, 1)):
        return self.lib.ora_add_object(self.h, mesh, material, _fp(_f32(position)),
                                       _fp(_f32(rotation)), _fp(_f32(scale)))

    def perf_mesh(self, mesh, origins, dirs):
        """TestMeshMidphase's loop (perf_tests.cpp:240-276): t and triangle per ray, wall seconds."""
        o, d = _f32(origins).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        t, tri = np.zeros(len(o), np.float32), np.zeros(len(o), np.int32)
        secs = self.lib.ora_perf_mesh(self.h, mesh, len(o), _fp(o), _fp(d), _fp(t), tri.ctypes.data_as(C.POINTER(C.c_int32)))
        return t, tri, secs

    def build(self):
        self.lib.ora_build(self.h)

    def register_material(self, id, albedo=(0, 0, 0), albedo_texture=0xFFFFFFFF,
                          emission=(0, 0, 0), emission_texture=0xFFFFFFFF, roughness=0.0):
        return self.lib.ora_register_material(self.h, id, _fp(_f32(albedo)), albedo_texture,
                                              _fp(_f32(emission)), emission_texture, roughness)

    def register_texture(self, id, image):
        img = _f32(image)
        self._keep.append(img)
        return self.lib.ora_register_texture(self.h, id, _fp(img), img.shape[1], img.shape[0])

    def set_background(self, material_id):
        self.lib.ora_set_background(self.h, material_id)

    def configure_camera(self, position, rotation, film_distance, width, height):
        self.width, self.height = width, height
        self.lib.ora_configure_camera(self.h, _fp(_f32(position)), _fp(_f32(rotation)),
                                      film_distance, width, height)

    def load_workload(self, wl):
        for m in wl.meshes:
            self.add_mesh(m.vertices, m.indices, m.smooth)
        for ob in wl.objects:
            if self.add_object(ob.mesh, ob.material, ob.position, ob.rotation, ob.scale) < 0:
                raise RuntimeError("object cap reached in " + self.o.name)
        self.build()
        for tid, img in wl.textures.items():
            self.register_texture(tid, img)
        for m in wl.materials:
            self.register_material(m.id, m.albedo, m.albedo_texture, m.emission,
                                   m.emission_texture, m.roughness)
        self.set_background(wl.background)
        self.configure_camera(wl.camera_position, wl.camera_rotation, wl.film_distance, wl.width,
                              wl.height)
        return self

    def render_seeded(self, spp=1, bounces=3, frame=0, threads=None, rect=None, image=None):
        threads = threads or os.cpu_count() or 1
        if image is None:
            image = np.zeros((self.height, self.width, 4), np.float32)
        x0, y0, x1, y1 = rect or (0, 0, self.width, self.height)
        metrics = np.zeros(12, np.uint64)
        self.lib.ora_render_seeded(self.h, _fp(image), x0, y0, x1, y1, spp, bounces, frame,
                                   threads, metrics.ctypes.data_as(_u64))
        return image, metrics

    def tie_mask(self, spp=1, bounces=3, frame=0, threads=None, rect=None):
        """uint8 (H, W): 1 where some scene query on the pixel's paths ended in an exact-t tie (port
        only; the verbatim reference cannot report it)."""
        threads = threads or os.cpu_count() or 1
        mask = np.zeros((self.height, self.width), np.uint8)
        x0, y0, x1, y1 = rect or (0, 0, self.width, self.height)
        self.lib.ora_tie_mask(self.h, mask.ctypes.data_as(C.POINTER(C.c_uint8)), x0, y0, x1, y1, spp,
                              bounces, frame, threads)
        return mask

    def render_tiles(self, tile_w=64, tile_h=64, spp=1, bounces=3, threads=None):
        threads = threads or os.cpu_count() or 1
        image = np.zeros((self.height, self.width, 4), np.float32)
        metrics = np.zeros(12, np.uint64)
        secs = self.lib.ora_render_tiles(self.h, _fp(image), tile_w, tile_h, spp, bounces, threads,
                                         metrics.ctypes.data_as(_u64))
        return image, metrics, secs

    def render_tile_list(self, tiles, spp=1, bounces=3, threads=None, image=None):
        """tiles: (n, 4) uint32 (minX minY maxX maxY).  Returns image, metrics, wall seconds."""
        threads = threads or os.cpu_count() or 1
        if image is None:
            image = np.zeros((self.height, self.width, 4), np.float32)
        t = np.ascontiguousarray(tiles, dtype=np.uint32).reshape(-1, 4)
        metrics = np.zeros(12, np.uint64)
        secs = self.lib.ora_render_tile_list(self.h, _fp(image), _up(t), len(t), spp, bounces,
                                             threads, metrics.ctypes.data_as(_u64))
        return image, metrics, secs

    def path_trace_tile(self, image, tile, spp, bounces, rng_state):
        state = C.c_uint32(rng_state)
        metrics = np.zeros(12, np.uint64)
        self.lib.ora_path_trace_tile(self.h, _fp(image), tile[0], tile[1], tile[2], tile[3], spp,
                                     bounces, C.byref(state), metrics.ctypes.data_as(_u64))
        return state.value, metrics

    def primary_hits(self, sample=0, frame=0, threads=None, want_dirs=False):
        threads = threads or os.cpu_count() or 1
        n = self.width * self.height
        tri = np.zeros(n, np.int32)
        obj = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        dirs = np.zeros((n, 3), np.float32) if want_dirs else None
        self.lib.ora_primary_hits(self.h, _ip(tri), _ip(obj), _fp(t), _fp(dirs), sample, frame,
                                  threads)
        shape = (self.height, self.width)
        out = {"tri": tri.reshape(shape), "obj": obj.reshape(shape), "t": t.reshape(shape)}
        if want_dirs:
            out["dir"] = dirs.reshape(shape + (3,))
        return out

    def intersect_rays(self, origins, dirs, want_ids=True):
        o = _f32(origins).reshape(-1, 3)
        d = _f32(dirs).reshape(-1, 3)
        n = len(o)
        out = np.zeros((n, 7), np.float32)
        tri = np.zeros(n, np.int32) if want_ids else None
        obj = np.zeros(n, np.int32) if want_ids else None
        metrics = np.zeros(12, np.uint64)
        self.lib.ora_intersect_rays(self.h, n, _fp(o), _fp(d), _fp(out), _ip(tri), _ip(obj),
                                    metrics.ctypes.data_as(_u64))
        return {"t": out[:, 0].copy(), "material": out[:, 1].copy().view(np.uint32),
                "normal": out[:, 2:5].copy(), "uv": out[:, 5:7].copy(), "tri": tri, "obj": obj,
                "metrics": metrics}

    def film_positions(self, pixel_positions):
        p = _f32(pixel_positions).reshape(-1, 2)
        out = np.zeros((len(p), 3), np.float32)
        self.lib.ora_film_positions(self.h, len(p), _fp(p), _fp(out))
        return out

    def radiance_for_path(self, path):
        """path: list of dicts(materialId, worldPosition, outgoingDir, incomingDir, normal, uv)."""
        arr = np.zeros((len(path), 15), np.float32)
        for k, v in enumerate(path):
            arr[k, 0] = np.array([v.get("materialId", 0)], np.uint32).view(np.float32)[0]
            arr[k, 1:4] = v.get("worldPosition", (0, 0, 0))
            arr[k, 4:7] = v.get("outgoingDir", (0, 0, 0))
            arr[k, 7:10] = v.get("incomingDir", (0, 0, 0))
            arr[k, 10:13] = v.get("normal", (0, 0, 0))
            arr[k, 13:15] = v.get("uv", (0, 0))
        out = np.zeros(3, np.float32)
        self.lib.ora_radiance_for_path(self.h, _fp(arr), len(path), _fp(out))
        return out

    def mesh_tree_stats(self, mesh):
        out = np.zeros(6, np.uint32)
        self.lib.ora_mesh_tree_stats(self.h, mesh, _up(out))
        return {"leaves": int(out[0]), "internal": int(out[1]), "min_depth": int(out[2]),
                "max_depth": int(out[3]), "all_reachable": bool(out[4]),
                "parents_contain_children": bool(out[5])}


HOSTSIM_LIB = os.path.join(_ROOT, "tests", "hostsim", "libhostsim.so")


def build_hostsim():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "tests", "hostsim")])


def load_hostsim():
    o = OracleLib(HOSTSIM_LIB)
    o.lib.hostsim_set_cull.argtypes = [C.c_int]
    o.lib.hostsim_check_shortcuts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _u64]
    o.lib.hostsim_check_coverage.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _u64]
    o.lib.hostsim_set_stepped.argtypes = [C.c_int]
    o.lib.hostsim_check_wide_slab.argtypes = [C.c_uint32, C.c_uint32, _u64]
    o.lib.hostsim_flat_info.argtypes = [C.c_void_p, _u]
    o.lib.hostsim_set_triangle_test.argtypes = [C.c_uint32]
    return o


def load_ref():
    return OracleLib(REF_LIB)


def load_ref_dm():
    return OracleLib(REF_DM_LIB)


def load_port():
    return OracleLib(PORT_LIB)


def load_port_dm():
    return OracleLib(PORT_DM_LIB)
