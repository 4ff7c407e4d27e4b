"""Synthetic, seeded inputs for the BASELINE.json configurations.

Pure numpy host data: meshes (from assets/*.npz, made by tools/make_mesh_fixtures.py),
procedural equirectangular environment maps standing in for the reference's git-LFS EXR files
(studio_garden_4k.exr / kiara_4_mid-morning_4k.exr are 133-byte pointers in the reference
checkout), camera framing and materials.  The same arrays are handed to the CUDA library and to
the CPU checkers, so nothing here is on the measured path.
"""
import os
from dataclasses import dataclass, field

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSET_DIR = os.path.join(_ROOT, "assets")

U32_MAX = 0xFFFFFFFF

# ids used by every workload (reference convention: enum index == id, main.cpp:1167-1189)
MATERIAL_BACKGROUND = 0
MATERIAL_SURFACE = 1
MATERIAL_CHECKER = 2
IMAGE_ENV = 7          # any value but 0: zero-initialised materials look up image key 0
IMAGE_CHECKER = 3


@dataclass
class Mesh:
    vertices: np.ndarray  # (n, 8) float32: position3 normal3 uv2  == VertexPNT (mesh.h:11-16)
    indices: np.ndarray   # (m,) uint32, m % 3 == 0
    smooth: bool = False

    @property
    def triangle_count(self):
        return len(self.indices) // 3

    def bounds(self):
        p = self.vertices[:, 0:3]
        return p.min(axis=0), p.max(axis=0)


@dataclass
class SceneObject:
    mesh: int
    material: int
    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (0.0, 0.0, 0.0, 1.0)   # quat x y z w
    scale: tuple = (1.0, 1.0, 1.0)


@dataclass
class Material:
    id: int
    albedo: tuple = (0.0, 0.0, 0.0)
    albedo_texture: int = U32_MAX
    emission: tuple = (0.0, 0.0, 0.0)
    emission_texture: int = U32_MAX
    roughness: float = 0.0


@dataclass
class Workload:
    name: str
    meshes: list
    objects: list
    materials: list
    textures: dict            # id -> (h, w, 4) float32 array
    background: int
    camera_position: tuple
    camera_rotation: tuple
    film_distance: float
    width: int
    height: int
    spp: int = 1
    bounces: int = 3
    notes: dict = field(default_factory=dict)


def load_mesh(name, smooth=False):
    data = np.load(os.path.join(ASSET_DIR, name + ".npz"))
    return Mesh(np.ascontiguousarray(data["vertices"], dtype=np.float32),
                np.ascontiguousarray(data["indices"], dtype=np.uint32), smooth)


# --------------------------------------------------------------------------------------------
# procedural meshes (reference src/mesh_generation.cpp; counts match, construction is ours)

def plane_mesh():
    v = np.array([[-0.5, -0.5, 0, 0, 0, 1, 0, 0], [0.5, -0.5, 0, 0, 0, 1, 1, 0],
                  [0.5, 0.5, 0, 0, 0, 1, 1, 1], [-0.5, 0.5, 0, 0, 0, 1, 0, 1]], dtype=np.float32)
    return Mesh(v, np.array([0, 1, 2, 2, 3, 0], dtype=np.uint32))


def triangle_mesh():
    v = np.array([[-0.5, -0.5, 0, 0, 0, 1, 0, 0], [0.5, -0.5, 0, 0, 0, 1, 1, 0],
                  [0.0, 0.5, 0, 0, 0, 1, 1, 1]], dtype=np.float32)
    return Mesh(v, np.array([0, 1, 2], dtype=np.uint32))


def icosphere_mesh(level, smooth=True):
    """Unit icosphere, 20 * 4**level triangles, outward winding (shared midpoints)."""
    t = (1.0 + 5.0 ** 0.5) * 0.5
    pts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t),
           (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in pts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
             (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
             (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache, new_faces = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for (i, j, k) in faces:
            a, b, c = mid(i, j), mid(j, k), mid(k, i)
            new_faces += [(i, a, c), (j, b, a), (k, c, b), (a, b, c)]
        faces = new_faces
    p = np.asarray(verts, dtype=np.float32)
    v = np.zeros((len(p), 8), dtype=np.float32)
    v[:, 0:3] = p
    v[:, 3:6] = p
    return Mesh(v, np.asarray(faces, dtype=np.uint32).reshape(-1), smooth)


# --------------------------------------------------------------------------------------------
# environment maps

def _hash32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


_ENV_CACHE = {}


def make_env_map(width=4096, height=2048, variant="studio_garden"):
    """Seeded procedural equirect RGBA-f32 map: vertical sky gradient, one sun disc whose peak
    exceeds RADIANCE_CLAMP (10) so the clamp is exercised, and per-texel value noise
    (seed 0x45BA12F3, the reference's cubemap seed, cubemap.cpp:123).  Row 0 is v = 0."""
    key = (width, height, variant)
    if key in _ENV_CACHE:
        return _ENV_CACHE[key]
    sun_az, sun_el, tint, seed = {
        "studio_garden": (0.9, 0.55, (1.0, 0.95, 0.85), 0x45BA12F3),
        "kiara": (2.4, 0.35, (1.0, 0.85, 0.7), 0x45BA12F4),
    }[variant]
    with np.errstate(over="ignore"):
        ys = (np.arange(height, dtype=np.float32) + 0.5) / np.float32(height)
        xs = (np.arange(width, dtype=np.float32) + 0.5) / np.float32(width)
        # image row y <-> uv.y = y/H after the reference's flip: uv.y = 1 - (cos(inc)/2 + 1/2)
        cos_inc = (1.0 - 2.0 * ys).astype(np.float32)            # +1 = straight up
        sin_inc = np.sqrt(np.maximum(0.0, 1.0 - cos_inc * cos_inc)).astype(np.float32)
        az = (xs * np.float32(2.0 * np.pi)).astype(np.float32)
        sky = 0.35 + 0.65 * np.clip(cos_inc * 0.5 + 0.5, 0, 1) ** 2
        ground = 0.12 + 0.1 * np.clip(-cos_inc, 0, 1)
        base = np.where(cos_inc >= 0, sky, ground).astype(np.float32)[:, None]
        sky_rgb = np.array([0.55, 0.72, 1.0], dtype=np.float32)
        gnd_rgb = np.array([0.42, 0.36, 0.30], dtype=np.float32)
        rgb = np.where((cos_inc >= 0)[:, None, None], sky_rgb[None, None, :],
                       gnd_rgb[None, None, :]) * base[:, :, None]
        rgb = np.broadcast_to(rgb, (height, width, 3)).astype(np.float32).copy()
        # sun disc
        sd = np.array([np.cos(sun_el) * np.cos(sun_az), np.sin(sun_el),
                       np.cos(sun_el) * np.sin(sun_az)], dtype=np.float32)
        dx = sin_inc[:, None] * np.cos(az)[None, :]
        dz = sin_inc[:, None] * np.sin(az)[None, :]
        dy = np.broadcast_to(cos_inc[:, None], dx.shape)
        cosang = dx * sd[0] + dy * sd[1] + dz * sd[2]
        sun = np.clip((cosang - np.float32(0.9985)) / np.float32(0.0015), 0, 1).astype(np.float32)
        glow = np.clip((cosang - np.float32(0.9)) / np.float32(0.1), 0, 1).astype(np.float32) ** 4
        rgb += (sun[:, :, None] * np.float32(60.0) + glow[:, :, None] * np.float32(1.5)) * \
            np.asarray(tint, dtype=np.float32)[None, None, :]
        # value noise, +-12 %
        idx = np.arange(width * height, dtype=np.uint32).reshape(height, width)
        n = _hash32(idx ^ np.uint32(seed)).astype(np.float32) / np.float32(4294967296.0)
        rgb *= (np.float32(0.88) + np.float32(0.24) * n)[:, :, None]
    out = np.ones((height, width, 4), dtype=np.float32)
    out[:, :, 0:3] = rgb
    out = np.ascontiguousarray(out)
    _ENV_CACHE[key] = out
    return out


def make_checkerboard(size=256, cells=8):
    """RGBA-f32 checkerboard (stands in for Image_CheckerBoard, main.cpp:1304)."""
    y, x = np.mgrid[0:size, 0:size]
    c = (((x * cells) // size + (y * cells) // size) & 1).astype(np.float32)
    out = np.ones((size, size, 4), dtype=np.float32)
    out[:, :, 0] = 0.1 + 0.8 * c
    out[:, :, 1] = 0.1 + 0.8 * c
    out[:, :, 2] = 0.15 + 0.7 * c
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------------------------
# workloads

def frame_camera(bounds_min, bounds_max, distance_factor=1.2):
    """Camera on +z looking down -z at aabbCentre + (0, 0, 1.2 |diag|)."""
    lo = np.asarray(bounds_min, dtype=np.float64)
    hi = np.asarray(bounds_max, dtype=np.float64)
    centre = (lo + hi) * 0.5
    diag = float(np.linalg.norm(hi - lo))
    pos = centre + np.array([0.0, 0.0, distance_factor * diag])
    return tuple(float(np.float32(v)) for v in pos)


def _standard_materials(env_id=IMAGE_ENV):
    return [
        Material(MATERIAL_BACKGROUND, emission_texture=env_id),
        Material(MATERIAL_SURFACE, albedo=(0.18, 0.18, 0.18), roughness=0.6),
    ]


def single_mesh_workload(mesh_name, width, height, spp=1, bounces=3, smooth=True,
                         env_variant="studio_garden", env_size=(4096, 2048), name=None):
    mesh = load_mesh(mesh_name, smooth=smooth)
    lo, hi = mesh.bounds()
    return Workload(
        name=name or f"{mesh_name}_{width}x{height}_{spp}spp_{bounces}b",
        meshes=[mesh],
        objects=[SceneObject(0, MATERIAL_SURFACE)],
        materials=_standard_materials(),
        textures={IMAGE_ENV: make_env_map(env_size[0], env_size[1], env_variant)},
        background=MATERIAL_BACKGROUND,
        camera_position=frame_camera(lo, hi),
        camera_rotation=(0.0, 0.0, 0.0, 1.0),
        film_distance=0.8,
        width=width, height=height, spp=spp, bounces=bounces,
        notes={"mesh": mesh_name, "triangles": mesh.triangle_count, "env": env_variant},
    )


def config1(width=1024, height=768, **kw):
    """BASELINE configs[0]: bunny + studio_garden, 1024x768, 1 spp (reference: 3 bounces)."""
    return single_mesh_workload("bunny", width, height, spp=1, bounces=3, smooth=True,
                                env_variant="studio_garden", name="C1_bunny_1024x768_1spp", **kw)


def config2(width=1920, height=1080, **kw):
    """BASELINE configs[1]: monkey primary rays, flat shading (triangle-ID check)."""
    return single_mesh_workload("monkey", width, height, spp=1, bounces=1, smooth=False,
                                env_variant="studio_garden", name="C2_monkey_1920x1080_primary",
                                **kw)


def config3(width=3840, height=2160, spp=64, bounces=5, **kw):
    """BASELINE configs[2]: bunny + kiara, 4K, 64 spp, 5 bounces."""
    return single_mesh_workload("bunny", width, height, spp=spp, bounces=bounces, smooth=True,
                                env_variant="kiara", name="C3_bunny_3840x2160_64spp_5b", **kw)


def quat_axis_angle(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    s = np.sin(angle * 0.5)
    return (float(a[0] * s), float(a[1] * s), float(a[2] * s), float(np.cos(angle * 0.5)))


def multi_object_workload(width=320, height=240, spp=1, bounces=3, count=12, seed=0x1A34C249,
                          env_size=(512, 256)):
    """Small instanced scene (<= 32 objects so the verbatim reference can run it): bunnies,
    icospheres and a textured ground plane with rotations and non-unit uniform scales."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    bunny = load_mesh("bunny", smooth=False)
    sphere = icosphere_mesh(2, smooth=True)
    plane = plane_mesh()
    meshes = [bunny, sphere, plane]
    objects = [SceneObject(2, MATERIAL_CHECKER, position=(0.0, -0.6, 0.0),
                           rotation=quat_axis_angle((1, 0, 0), -np.pi / 2), scale=(8.0, 8.0, 8.0))]
    for i in range(count):
        kind = i % 2
        pos = (float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-0.4, 0.8)),
               float(rng.uniform(-1.5, 0.5)))
        rot = quat_axis_angle(rng.uniform(-1, 1, 3) + 1e-3, rng.uniform(0, 2 * np.pi))
        if kind == 0:
            s = float(rng.uniform(2.0, 4.0))
        else:
            s = float(rng.uniform(0.15, 0.35))
        objects.append(SceneObject(kind, MATERIAL_SURFACE, pos, rot, (s, s, s)))
    materials = _standard_materials() + [
        Material(MATERIAL_CHECKER, albedo=(0.5, 0.5, 0.5), albedo_texture=IMAGE_CHECKER,
                 roughness=0.3)]
    return Workload(
        name=f"multi_{count}obj_{width}x{height}", meshes=meshes, objects=objects,
        materials=materials,
        textures={IMAGE_ENV: make_env_map(env_size[0], env_size[1], "studio_garden"),
                  IMAGE_CHECKER: make_checkerboard()},
        background=MATERIAL_BACKGROUND,
        camera_position=(0.0, 0.4, 4.0), camera_rotation=(0.0, 0.0, 0.0, 1.0),
        film_distance=0.8, width=width, height=height, spp=spp, bounces=bounces,
        notes={"objects": len(objects)})


def nested_objects_workload(count=48, width=96, height=64, spp=1, bounces=3, env_size=(128, 64)):
    """Pathological TLAS input: `count` quads, each 1.5x the size of the one in front
    of it, so every object's box contains all the smaller ones' in x and y and has 2.25x their
    area.  A surface-area split peels such boxes off one at a time -- cost 1 + (n - 1) / 2.25
    against n / 2 for the middle split -- giving a tree as deep as the object count: more than the
    traversal stack's TLAS share holds.  The builder must notice and rebuild balanced."""
    plane = plane_mesh()
    objects = []
    for i in range(count):
        s = 1.0e-3 * (1.5 ** i)
        objects.append(SceneObject(0, MATERIAL_SURFACE if i % 2 else MATERIAL_CHECKER,
                                   position=(0.0, 0.0, -0.01 * i), scale=(s, s, s)))
    materials = _standard_materials() + [Material(MATERIAL_CHECKER, albedo=(0.6, 0.3, 0.2), roughness=0.35)]
    return Workload(
        name=f"nested_{count}obj_{width}x{height}", meshes=[plane], objects=objects, materials=materials,
        textures={IMAGE_ENV: make_env_map(env_size[0], env_size[1], "studio_garden")},
        background=MATERIAL_BACKGROUND, camera_position=(0.0, 0.0, 3.0),
        camera_rotation=(0.0, 0.0, 0.0, 1.0), film_distance=0.8, width=width, height=height,
        spp=spp, bounces=bounces, notes={"objects": count})


def crack_test_inputs(n=24, rays=200000, seed=0x45BA12F3):
    """A wavy sheet of 2 n^2 triangles facing +z and `rays` rays from seeded origins above it, each
    aimed at a point that lies EXACTLY on an interior shared edge or vertex of the sheet (vertices,
    edge midpoints, random points of edges).  The direction is rounded to f32, so each ray passes within
    an ulp of the edge: a test that evaluates the two neighbours independently can reject the ray on
    both sides (a crack); a watertight test cannot.  Returns (mesh, origins, dirs)."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    xs = np.linspace(-1.0, 1.0, n + 1).astype(np.float32)
    gx, gy = np.meshgrid(xs, xs)
    gz = (0.15 * np.sin(3.0 * gx) * np.cos(3.0 * gy)).astype(np.float32)
    v = np.zeros(((n + 1) * (n + 1), 8), np.float32)
    v[:, 0], v[:, 1], v[:, 2] = gx.ravel(), gy.ravel(), gz.ravel()
    v[:, 5] = 1.0
    idx = []
    for j in range(n):
        for i in range(n):
            a, b = j * (n + 1) + i, j * (n + 1) + i + 1
            c, d = a + n + 1, b + n + 1
            idx += [a, b, d, a, d, c]
    mesh = Mesh(v, np.asarray(idx, np.uint32))
    pos = v[:, 0:3].reshape(n + 1, n + 1, 3)
    j = rng.randint(1, n - 1, rays)
    i = rng.randint(1, n - 1, rays)
    kind = rng.randint(0, 4, rays)
    p0 = pos[j, i]
    # the other end of an edge leaving (i, j): +x, +y or the diagonal of the cell (a -> d)
    di = np.where(kind == 2, 0, 1)
    dj = np.where(kind == 1, 0, 1)
    dj = np.where(kind == 0, 0, dj)
    di = np.where(kind == 0, 0, di)
    p1 = pos[j + dj, i + di]
    w = np.where(kind == 0, 0.0, np.where(rng.rand(rays) < 0.5, 0.5, rng.rand(rays))).astype(np.float32)
    target = (p0 + (p1 - p0) * w[:, None]).astype(np.float32)
    origins = np.stack([rng.uniform(-1.5, 1.5, rays), rng.uniform(-1.5, 1.5, rays), rng.uniform(1.0, 4.0, rays)], axis=1).astype(np.float32)
    d = (target - origins).astype(np.float32)
    d = (d / np.sqrt((d * d).sum(axis=1, dtype=np.float32))[:, None]).astype(np.float32)
    return mesh, origins, d


# ------------------------------------------------------------------------------------------------
# the reference's own performance tests (perf_tests/perf_tests.cpp), inputs restated draw for draw.
# `draw(n)` returns the next n RandomBilateral values of ONE XorShift32 stream (seed 0x1A34C249,
# perf_tests.cpp:57,214): sp.xorshift_bilateral_stream on the product side, ora's on the checker's.

def _lerp(a, b, t):
    return (np.float32(a) * (np.float32(1.0) - t) + np.float32(b) * t).astype(np.float32)


def _perf_rays(draw, lo, hi, count):
    """perf_tests.cpp:80-99 / :243-258: per ray six draws in statement order x0 y0 z0 x1 y1 z1,
    Lerp(min, max, RandomBilateral) (t in [-1, 1]: points may lie outside the box), Normalize(q - p)."""
    r = draw(count * 6).reshape(count, 6)
    p = np.stack([_lerp(lo[k], hi[k], r[:, k]) for k in range(3)], axis=1)
    q = np.stack([_lerp(lo[k], hi[k], r[:, 3 + k]) for k in range(3)], axis=1)
    d = (q - p).astype(np.float32)
    # Normalize (math_lib.h:503-512): sqrt of the dot product in f32, multiply by the reciprocal
    length = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
    inv = (np.float32(1.0) / length).astype(np.float32)
    return p.astype(np.float32), (d * inv[:, None]).astype(np.float32)


def perf_bvh_inputs(draw, boxes=2048, rays=8192, proper=False):
    """TestBvh (perf_tests.cpp:51-118): 2048 boxes -- centre 2000 * (B, B, B) with the THREE draws of the
    Vec3(...) call made right to left (g++ evaluates call arguments last to first: z, y, x), then radius
    500 * B, min = centre - radius, max = centre + radius -- and 8192 rays inside the root's bounds.
    Half the radii are NEGATIVE, i.e. half the boxes have min > max.  The slab test orders each axis
    itself, so such a leaf is tested like its reordered box; but the reference's tree unions boxes with
    Min(min) / Max(max), which for inverted children is not a union at all, so which of those leaves
    bvh_IntersectRay reports depends on its tree.  `proper=True` uses |radius| (same draws): the input
    on which "every leaf whose own box the ray passes" is well defined and tree-independent."""
    r = draw(boxes * 4).reshape(boxes, 4)
    s = np.float32(2000.0)
    centre = np.stack([s * r[:, 2], s * r[:, 1], s * r[:, 0]], axis=1).astype(np.float32)
    radius = (np.float32(500.0) * r[:, 3]).astype(np.float32)
    if proper:
        radius = np.abs(radius)
    mn = (centre - radius[:, None]).astype(np.float32)
    mx = (centre + radius[:, None]).astype(np.float32)
    lo, hi = mn.min(axis=0), mx.max(axis=0)     # tree.root->min / max: Min over the mins, Max over the maxes
    origins, dirs = _perf_rays(draw, lo, hi, rays)
    return mn, mx, origins, dirs


def boxes_as_triangles(mn, mx):
    """A mesh whose triangle i has exactly box i as its AABB (vertices at min, max and (min.x, max.y,
    min.z)): bvh_IntersectRay over boxes becomes the leaf query of a mesh tree.  The per-axis order of
    min / max does not matter to the slab test (simd.h:226-239 orders t0, t1 itself)."""
    n = len(mn)
    v = np.zeros((n * 3, 8), np.float32)
    v[0::3, 0:3] = mn
    v[1::3, 0:3] = mx
    v[2::3, 0] = mn[:, 0]
    v[2::3, 1] = mx[:, 1]
    v[2::3, 2] = mn[:, 2]
    v[:, 5] = 1.0
    return Mesh(v, np.arange(n * 3, dtype=np.uint32))


def perf_mesh_inputs(draw, rays, level=3):
    """TestMeshMidphase (perf_tests.cpp:212-305): a level-3 icosphere (1280 triangles) and `rays` rays
    (the reference: 1024 * 8192) between seeded points of its bounding box."""
    mesh = icosphere_mesh(level, smooth=False)
    lo, hi = mesh.bounds()
    origins, dirs = _perf_rays(draw, lo.astype(np.float32), hi.astype(np.float32), rays)
    return mesh, origins, dirs


def config5(width=3840, height=2160, spp=16, bounces=5, bunnies=64, spheres=118, sphere_level=6,
            unique_spheres=False, seed=0x1A34C249, env_size=(4096, 2048)):
    """BASELINE configs[4]: procedural instanced scene, ~10 M triangles at the defaults
    (64 bunnies x 4 968 + 118 icospheres x 81 920 = 9.98 M), on a seeded jittered grid
    (XorShift-style seed 0x1A34C249, the reference's perf-test seed, perf_tests.cpp:57).
    The reference cannot run it (32-object table, sp_scene.h:15; O(n^2 log n) build), so the
    checker is the port.  unique_spheres=True gives every sphere its own mesh (no instancing):
    the BVH then no longer fits L2 -- the HBM-bound variant."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    bunny = load_mesh("bunny", smooth=True)
    meshes = [bunny]
    count = bunnies + spheres
    side = int(np.ceil(count ** (1.0 / 3.0)))
    cells = [(i, j, k) for i in range(side) for j in range(side) for k in range(side)]
    order = rng.permutation(len(cells))[:count]
    objects = []
    sphere_meshes = 0
    for n, ci in enumerate(order):
        cell = np.array(cells[ci], dtype=np.float64)
        pos = (cell - (side - 1) / 2.0) * 1.0 + rng.uniform(-0.18, 0.18, 3)
        rot = quat_axis_angle(rng.uniform(-1, 1, 3) + 1e-3, rng.uniform(0, 2 * np.pi))
        if n < bunnies:
            sc = float(rng.uniform(2.2, 3.2))
            objects.append(SceneObject(0, MATERIAL_SURFACE, tuple(pos), rot, (sc, sc, sc)))
        else:
            if unique_spheres or sphere_meshes == 0:
                meshes.append(icosphere_mesh(sphere_level, smooth=True))
                sphere_meshes += 1
            sc = float(rng.uniform(0.25, 0.38))
            objects.append(SceneObject(len(meshes) - 1, MATERIAL_CHECKER if n % 3 == 0 else MATERIAL_SURFACE,
                                       tuple(pos), rot, (sc, sc, sc)))
    tris = sum(meshes[o.mesh].triangle_count for o in objects)
    materials = _standard_materials() + [
        Material(MATERIAL_CHECKER, albedo=(0.6, 0.3, 0.2), roughness=0.35)]
    dist = 1.15 * side
    return Workload(
        name=f"C5_{count}obj_{tris}tris_{width}x{height}_{spp}spp", meshes=meshes, objects=objects,
        materials=materials, textures={IMAGE_ENV: make_env_map(env_size[0], env_size[1], "studio_garden")},
        background=MATERIAL_BACKGROUND, camera_position=(0.0, 0.0, float(dist)),
        camera_rotation=(0.0, 0.0, 0.0, 1.0), film_distance=0.8, width=width, height=height,
        spp=spp, bounces=bounces,
        notes={"objects": count, "instanced_triangles": int(tris), "unique_meshes": len(meshes)})
