"""Parity of the CUDA path with the oracle, through libspb200.so's C ABI (needs a B200).

Oracles: oracle/_ref/libspref*.so (the reference's unmodified sources; prebuilt .so files travel
with the snapshot) when present, else the port (oracle/libsporacle*.so, shown bit-equal to the
reference by tests/test_oracle_kat.py).  Bars:
  * integer / index work (triangle ids, object ids, hit counters, rng state): exact;
  * floating point in deterministic-math mode (mathMode 0 vs the *_dm checkers, which compute
    sin/cos/atan2/pow as "double, rounded once" on both sides): bit-exact images;
  * floating point against glibc's libm (the plain reference): per-pixel and RMSE tolerances
    written in the tests below.
"""
import ctypes as C
import os

import numpy as np
import pytest

import ora
from vk_cinematic_b200 import workloads as W
from vk_cinematic_b200.fixtures import lattice, relative_error_report, tile_crcs

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS = np.finfo(np.float32).eps


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


def assert_same_image_up_to_ties(img, cimg, chk, spp, bounces, frame, bound=1e-4):
    """Bit-equal images except at pixels where the checker itself met two candidates with bit-equal
    closest t on one of the pixel's paths (ora_tie_mask): there the winner depends on the order
    the tree presents them in, which differs between the reference's agglomerative tree and any
    other builder.  The number of such differing pixels is bounded (BASELINE.json: < 1e-4)."""
    diff = np.any(bits(img) != bits(cimg), axis=-1)
    if not diff.any():
        return 0
    ties = chk.tie_mask(spp=spp, bounces=bounces, frame=frame).astype(bool)
    unexplained = diff & ~ties
    assert not unexplained.any(), f"{int(unexplained.sum())} differing pixels without an exact-t tie: " \
                                  f"{list(zip(*np.nonzero(unexplained)))[:8]}"
    assert diff.sum() <= max(1, int(bound * diff.size)), f"{int(diff.sum())} tie pixels differ (bound {bound})"
    return int(diff.sum())


def assert_tiles_match(image, metrics, gold, key, ties_key, bound=1e-4):
    """Whole-frame comparison against per-tile CRC-32s of a checker's render (fixtures.tile_crcs):
    every 64x64 tile must hash to the checker's value, except tiles holding a pixel the checker
    recorded an exact-t tie for; ray / hit / miss counters equal up to those ties."""
    crc = tile_crcs(image)
    want = gold["tile_crc_" + key]
    assert crc.shape == want.shape
    bad = set(np.nonzero(crc != want)[0].tolist())
    tx = (image.shape[1] + 63) // 64
    tie_tiles = set(int(y // 64) * tx + int(x // 64) for x, y in gold[ties_key])
    assert bad <= tie_tiles, f"tiles differing without a tie pixel: {sorted(bad - tie_tiles)[:8]}"
    assert len(bad) <= max(1, int(bound * image.shape[0] * image.shape[1]))
    gm = gold["metrics_" + key].astype(np.int64)
    slack = 8 * len(gold[ties_key]) * 64
    assert int(metrics[1]) == int(gm[0]) and np.all(np.abs(metrics[2:5].astype(np.int64) - gm[1:4]) <= slack)
    return len(bad)


def best(dm):
    """The strongest available checker: the reference itself, else the port."""
    if ora.have_ref():
        return ora.load_ref_dm() if dm else ora.load_ref()
    return ora.load_port_dm() if dm else ora.load_port()


@pytest.fixture(params=["wavefront", "per_pixel"])
def params(gpu_sp, request):
    """Every test runs under both schedulers of sp_b200_Render*: the wavefront kernels (default)
    and the one-thread-per-pixel kernel."""
    mode = 0 if request.param == "wavefront" else 1
    gpu_sp.set_params(samplesPerPixel=1, bounceCount=3, radianceClamp=10.0, envFilter=0, mathMode=0,
                      cullByDistance=1, tileWidth=64, tileHeight=64, renderMode=mode, samplesPerPass=0)
    gpu_sp.lib.sp_b200_EnableStats(0)
    yield gpu_sp
    gpu_sp.set_params(samplesPerPixel=1, bounceCount=3, radianceClamp=10.0, envFilter=0, mathMode=0,
                      cullByDistance=1, tileWidth=64, tileHeight=64, renderMode=0, samplesPerPass=0)
    gpu_sp.lib.sp_b200_EnableStats(0)


# ---------------------------------------------------------------------------------------------
# the reference's unit tests, restated against the C ABI

TRI_VERTS = np.array([[-0.5, -0.5, 0, 0, 0, 1, 0, 0], [0.5, -0.5, 0, 0, 0, 1, 0, 0],
                      [0.0, 0.5, 0, 0, 0, 1, 0, 0]], np.float32)


def test_ray_intersect_scene_two_objects(params):
    """test_simd_path_tracer.cpp:216-276"""
    sp = params
    r = sp.Renderer()
    m = r.add_mesh(TRI_VERTS, [0, 1, 2])
    r.add_object(m, 53, (0, 2, -5))
    r.add_object(m, 53, (0, 2, -15), W.quat_axis_angle((0, 1, 0), 3.14159265359 * 0.25), (2, 2, 2))
    r.build()
    met = sp.sp_Metrics()
    res = sp.lib.sp_RayIntersectScene(C.byref(r.scene), sp.V3((0, 2, 0)), sp.V3((0, 0, -1)), C.byref(met))
    assert res.t >= 0.0 and res.materialId == 53 and abs(res.t - 5.0) < 1e-5
    chk = best(False).scene()
    cm = chk.add_mesh(TRI_VERTS, [0, 1, 2])
    chk.add_object(cm, 53, (0, 2, -5))
    chk.add_object(cm, 53, (0, 2, -15), W.quat_axis_angle((0, 1, 0), 3.14159265359 * 0.25), (2, 2, 2))
    chk.build()
    rng = np.random.RandomState(2)
    o = np.tile(np.float32((0, 2, 0)), (256, 1)) + rng.uniform(-0.3, 0.3, (256, 3)).astype(np.float32)
    d = np.float32((0, 0, -1)) + rng.uniform(-0.08, 0.08, (256, 3)).astype(np.float32)
    d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    g, e = r.intersect_rays(o, d), chk.intersect_rays(o, d)
    assert same_bits(g["t"], e["t"]) and np.array_equal(g["obj"], e["obj"])
    assert same_bits(g["normal"], e["normal"]) and np.array_equal(g["material"], e["material"])
    assert (e["obj"] == 0).any() and (e["obj"] == 1).any()
    chk.close()
    r.close()


def test_ray_intersect_mesh_single_triangle(params):
    """test_simd_path_tracer.cpp:370-396 (+ :398-423: one triangle -> root bounds = its AABB)"""
    sp = params
    r = sp.Renderer()
    r.add_mesh(TRI_VERTS, [0, 1, 2])
    res = sp.lib.sp_RayIntersectMesh(r.meshes[0], sp.V3((0, 0, 10)), sp.V3((0, 0, -1)), None)
    ti = res.triangleIntersection
    assert ti.t == 10.0 and ti.normal.tuple() == (0.0, 0.0, 1.0)
    res = sp.lib.sp_RayIntersectMesh(r.meshes[0], sp.V3((2, 0, 10)), sp.V3((0, 0, -1)), None)
    assert res.triangleIntersection.t == -1.0
    info = sp.sp_b200_TreeInfo()
    sp.lib.sp_b200_MeshTreeInfo(r.meshes[0], C.byref(info))
    assert info.leafCount == 1 and info.rootMin.tuple() == (-0.5, -0.5, 0.0)
    assert info.rootMax.tuple() == (0.5, 0.5, 0.0)
    r.close()


def test_path_trace_tile_magenta_bounds_metrics(params):
    """test_simd_path_tracer.cpp:49-135, 333-368: zero-initialised context, rng state 0"""
    sp = params
    pixels = np.zeros((4, 4, 4), np.float32)
    plane = sp.ImagePlane(pixels.ctypes.data_as(C.POINTER(sp.vec4)), 4, 4)
    cam = sp.sp_Camera()
    cam.imagePlane = C.pointer(plane)
    scene, ms = sp.sp_Scene(), sp.sp_MaterialSystem()
    ctx = sp.sp_Context(C.pointer(cam), C.pointer(scene), C.pointer(ms))
    rng, met = sp.RandomNumberGenerator(0), sp.sp_Metrics()
    sp.lib.sp_PathTraceTile(C.byref(ctx), sp.Tile(0, 0, 4, 4), C.byref(rng), C.byref(met))
    assert np.abs(pixels - np.float32((1, 0, 1, 1))).max() <= EPS
    pixels[:] = 0
    met = sp.sp_Metrics()
    sp.lib.sp_PathTraceTile(C.byref(ctx), sp.Tile(1, 1, 3, 3), C.byref(rng), C.byref(met))
    exp = np.zeros((4, 4, 4), np.float32)
    exp[1:3, 1:3] = (1, 0, 1, 1)
    assert np.array_equal(pixels, exp)
    assert met.values[sp.sp_Metric_CyclesElapsed] > 0
    assert met.values[sp.sp_Metric_PathsTraced] == 4 and met.values[sp.sp_Metric_RayMissCount] == 4
    # a tile reaching past the image is clamped (simd_path_tracer.cpp:188-191)
    pixels[:] = 0
    sp.lib.sp_PathTraceTile(C.byref(ctx), sp.Tile(2, 2, 9, 9), C.byref(rng), C.byref(met))
    assert pixels[2:, 2:, 3].min() == 1.0 and pixels[:2].max() == 0.0


def test_material_albedo_texture_and_light_path(params):
    """test_simd_path_tracer.cpp:278-331 (light path restated with roughness > 0)"""
    sp = params
    img = np.zeros((1, 1, 4), np.float32)
    img[0, 0] = (1, 0, 0, 1)
    ms = sp.sp_MaterialSystem()
    sp.lib.sp_RegisterTexture(C.byref(ms), sp.HdrImage(img.ctypes.data_as(C.POINTER(C.c_float)), 1, 1), 1)
    mat = sp.sp_Material(sp.V3((0, 0, 0)), 1, sp.V3((0, 0, 0)), sp.U32_MAX, 0.0)
    vert = sp.sp_PathVertex()
    out = sp.lib.sp_EvaluateMaterial(C.byref(ms), C.byref(mat), C.byref(vert))
    assert out.albedo.tuple() == (1.0, 0.0, 0.0)

    chk = best(True).scene()
    ms = sp.sp_MaterialSystem()
    for mid, kw in ((0, dict(emission=(1, 1, 1))), (1, dict(albedo=(0.18, 0.18, 0.18), roughness=0.6)),
                    (2, dict(albedo=(0.7, 0.2, 0.1), emission=(0.3, 0.0, 0.2), roughness=0.25))):
        m = sp.sp_Material(sp.V3(kw.get("albedo", (0, 0, 0))), sp.U32_MAX, sp.V3(kw.get("emission", (0, 0, 0))),
                           sp.U32_MAX, kw.get("roughness", 0.0))
        sp.lib.sp_RegisterMaterial(C.byref(ms), m, mid)
        chk.register_material(mid, **kw)
    rng = np.random.RandomState(9)
    for _ in range(40):
        n = int(rng.randint(1, 5))
        path, cpath = (sp.sp_PathVertex * n)(), []
        for i in range(n):
            v = {k: tuple(float(x) for x in (lambda a: a / np.linalg.norm(a))(rng.normal(size=3)))
                 for k in ("outgoingDir", "incomingDir", "normal")}
            v["materialId"] = int(rng.choice([1, 2, 7])) if i < n - 1 else 0   # 7: unregistered -> magenta
            path[i].materialId = v["materialId"]
            path[i].outgoingDir, path[i].incomingDir, path[i].normal = (sp.V3(v[k]) for k in
                                                                       ("outgoingDir", "incomingDir", "normal"))
            cpath.append(v)
        got = sp.lib.ComputeRadianceForPath(path, n, C.byref(ms))
        assert same_bits(np.float32(got.tuple()), chk.radiance_for_path(cpath))
    chk.close()


def test_bilinear_environment_lookup(params):
    """image.h:34-73 (SampleImageBilinear) as the env filter the north star asks for: emission of
    a material with an emission texture under envFilter = bilinear equals the oracle's bilinear
    sample at the oracle's equirect uv (v flipped, sp_material_system.cpp:88-93)."""
    sp = params
    chk = best(True)
    env = W.make_env_map(64, 32, "kiara")
    ms = sp.sp_MaterialSystem()
    sp.lib.sp_RegisterTexture(C.byref(ms), sp.HdrImage(env.ctypes.data_as(C.POINTER(C.c_float)), 64, 32), 4)
    mat = sp.sp_Material(sp.V3((0, 0, 0)), sp.U32_MAX, sp.V3((0, 0, 0)), 4, 0.0)
    rng = np.random.RandomState(4)
    for flt in (0, 1):
        sp.set_params(envFilter=flt)
        for _ in range(64):
            d = rng.normal(size=3)
            d = np.float32(d / np.linalg.norm(d))
            vert = sp.sp_PathVertex()
            vert.outgoingDir = sp.V3(d)
            out = sp.lib.sp_EvaluateMaterial(C.byref(ms), C.byref(mat), C.byref(vert))
            uv = chk.map_equirect(chk.to_spherical(-d))
            u, v = np.float32(uv[0]), np.float32(1.0) - np.float32(uv[1])
            exp = (chk.sample_bilinear if flt else chk.sample_nearest)(env, u, v)[0:3]
            assert same_bits(np.float32(out.emission.tuple()), exp), (flt, d)
    sp.lib.sp_b200_FlushTextureCache()


def test_intersected_leaves_match_bvh_query(params):
    """bvh_IntersectRay semantics (bvh.cpp:203-311; test_bvh.cpp:99-122,265-292): the set of
    leaves reported for a ray equals the reference's, independent of tree topology; too small a
    buffer sets errorOccurred."""
    sp = params
    chk = best(False)
    mesh = W.icosphere_mesh(2)
    r = sp.Renderer()
    r.add_mesh(mesh.vertices, mesh.indices)
    p = mesh.vertices[:, 0:3][mesh.indices.reshape(-1, 3)]
    mn, mx = p.min(axis=1), p.max(axis=1)
    rng = np.random.RandomState(6)
    for k in range(24):
        o = rng.uniform(-2, 2, 3).astype(np.float32)
        d = -o + rng.uniform(-0.5, 0.5, 3).astype(np.float32)
        d = np.float32(d / np.linalg.norm(d))
        if k == 0:
            o, d = np.float32((0, 0, 3)), np.float32((0, 0, -1))
        leaves = np.zeros(128, np.uint32)
        err = C.c_uint32(0)
        n = sp.lib.sp_b200_MeshIntersectedLeaves(r.meshes[0], sp.V3(o), sp.V3(d), leaves.ctypes.data, 128,
                                                 C.byref(err))
        e = chk.bvh_query(mn, mx, o, d, 128)
        assert not err.value and sorted(leaves[:n].tolist()) == sorted(e["leaves"].tolist())
        if n > 1:
            n2 = sp.lib.sp_b200_MeshIntersectedLeaves(r.meshes[0], sp.V3(o), sp.V3(d), leaves.ctypes.data, 1,
                                                      C.byref(err))
            assert n2 == 1 and err.value
    r.close()


def test_empty_and_unbuilt_scenes(params):
    """bvh.cpp:218-229 NULL root: every ray misses; background evaluated per pixel"""
    sp = params
    wl = W.config1(64, 48, env_size=(64, 32))
    wl.objects = []
    r = sp.Renderer().load_workload(wl)
    img, m = r.render_frame()
    chk = best(True).scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=1, bounces=3)
    assert same_bits(img, cimg) and m[sp.sp_Metric_RayMissCount] == 64 * 48
    chk.close()
    r.close()


# ---------------------------------------------------------------------------------------------
# golden fixtures made from the unmodified reference (tools/make_golden.py)

def test_golden_fixtures(params):
    sp = params
    g = np.load(os.path.join(GOLD, "g1_bunny_96x64.npz"))
    r = sp.Renderer().load_workload(W.config1(96, 64, env_size=(512, 256)))
    for cull in (1, 0):
        sp.set_params(samplesPerPixel=2, cullByDistance=cull)
        img, m = r.render_frame(frame=1)
        assert same_bits(img, g["image_ref_dm"])
        assert np.array_equal(m[1:5], g["metrics_ref_dm"])
        ph = r.primary_hits(sample=0, frame=1)
        assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
        r.image[:] = 0
        state, tm = r.path_trace_tile((16, 8, 48, 40), 0xF51C0E49)
        assert state == int(g["tile_state_ref_dm"]) and same_bits(r.image, g["tile_image_ref_dm"])
        assert np.array_equal(tm[1:5], g["tile_metrics_ref_dm"])
    r.close()
    sp.set_params(samplesPerPixel=1, cullByDistance=1)
    g = np.load(os.path.join(GOLD, "g2_monkey_160x90_primary.npz"))
    r = sp.Renderer().load_workload(W.config2(160, 90, env_size=(64, 32)))
    ph = r.primary_hits()
    assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
    r.close()
    g = np.load(os.path.join(GOLD, "g3_multi_80x60.npz"))
    r = sp.Renderer().load_workload(W.multi_object_workload(width=80, height=60, spp=2, env_size=(256, 128)))
    for cull in (1, 0):
        sp.set_params(samplesPerPixel=2, cullByDistance=cull)
        q = r.intersect_rays(g["origins"], g["dirs"])
        assert same_bits(q["t"], g["rays_t"]) and np.array_equal(q["tri"], g["rays_tri"])
        assert np.array_equal(q["obj"], g["rays_obj"]) and np.array_equal(q["material"], g["rays_material"])
        assert same_bits(q["normal"], g["rays_normal"]) and same_bits(q["uv"], g["rays_uv"])
        img, m = r.render_frame(frame=5)
        assert same_bits(img, g["image_ref_dm"]) and np.array_equal(m[1:5], g["metrics_ref_dm"])
    r.close()


# ---------------------------------------------------------------------------------------------
# BASELINE configurations at full size

def test_c2_monkey_primary_triangle_ids_full_size(params):
    """BASELINE configs[1]: monkey 1920x1080 primary rays.  Closest-hit triangle ids must match
    the reference except on equal-t ties; stated bound on the mismatch rate: < 1e-4."""
    sp = params
    wl = W.config2(1920, 1080, env_size=(64, 32))
    r = sp.Renderer().load_workload(wl)
    chk = best(False).scene().load_workload(wl)
    e = chk.primary_hits()
    for cull in (1, 0):
        sp.set_params(cullByDistance=cull)
        g = r.primary_hits()
        assert same_bits(g["t"], e["t"])                      # hit / miss and distance: exact
        mism = g["tri"] != e["tri"]
        assert mism.mean() < 1e-4
        assert (e["tri"] >= 0).sum() > 200000
        # every mismatch is a tie: same t bits were already asserted; both must be hits
        assert np.all(g["tri"][mism] >= 0) and np.all(e["tri"][mism] >= 0)
    chk.close()
    r.close()


def test_c1_bunny_image_full_size(params):
    """BASELINE configs[0]: bunny + studio_garden env (4096x2048), 1024x768, 1 spp, 3 bounces.
    vs reference-dm: bit-exact.  vs the plain reference (glibc sinf/cosf/atan2f/powf, which are
    not correctly rounded): stated tolerance -- at most 2 % of pixels may differ by more than
    1e-3 relative (a last-ulp difference in a bounce direction can change which texel of the
    noisy env map a path ends on), and whole-image RMSE <= 2 % of the mean radiance."""
    sp = params
    wl = W.config1(1024, 768)
    r = sp.Renderer().load_workload(wl)
    img, m = r.render_frame(frame=0)
    img = img.copy()
    chk = best(True).scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=1, bounces=3, frame=0)
    chk.close()
    assert same_bits(img, cimg) and np.array_equal(m[1:5], cm[1:5])
    chk = best(False).scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=1, bounces=3, frame=0)
    chk.close()
    d = np.abs(img[..., :3].astype(np.float64) - cimg[..., :3])
    rel = (d / np.maximum(np.abs(cimg[..., :3]), 1e-3)).max(axis=2)
    rmse = float(np.sqrt((d ** 2).mean()))
    assert (rel > 1e-3).mean() < 0.02, (rel > 1e-3).mean()
    assert rmse <= 0.02 * float(cimg[..., :3].mean()), rmse
    assert abs(int(m[2]) - int(cm[2])) <= 1e-3 * int(cm[2])   # rays traced: path topology
    # fast f32 math mode (CUDA sinf/cosf/atan2f/powf): same tolerance against the reference
    sp.set_params(mathMode=1)
    img1, m1 = r.render_frame(frame=0)
    d = np.abs(img1[..., :3].astype(np.float64) - cimg[..., :3])
    rel = (d / np.maximum(np.abs(cimg[..., :3]), 1e-3)).max(axis=2)
    assert (rel > 1e-3).mean() < 0.02 and float(np.sqrt((d ** 2).mean())) <= 0.02 * float(cimg[..., :3].mean())
    r.close()


def test_multi_object_scene_vs_reference(params):
    """Instanced scene (13 objects, rotations, non-unit scales, textured plane): image, object
    and triangle ids against the reference."""
    sp = params
    wl = W.multi_object_workload(width=320, height=240, spp=2)
    r = sp.Renderer().load_workload(wl)
    chk = best(True).scene().load_workload(wl)
    sp.set_params(samplesPerPixel=2)
    img, m = r.render_frame(frame=1)
    cimg, cm = chk.render_seeded(spp=2, bounces=3, frame=1)
    assert same_bits(img, cimg) and np.array_equal(m[1:5], cm[1:5])
    g, e = r.primary_hits(), chk.primary_hits()
    assert np.array_equal(g["obj"], e["obj"]) and np.array_equal(g["tri"], e["tri"]) and same_bits(g["t"], e["t"])
    chk.close()
    r.close()


def test_five_bounces_vs_port(params):
    """BASELINE config 3 asks for 5 bounces; the reference is fixed at 3 (literal at
    simd_path_tracer.cpp:195), so the checker is the port, shown equal to the reference at 3."""
    sp = params
    wl = W.config3(256, 144, spp=4, bounces=5, env_size=(1024, 512))
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=4, bounceCount=5)
    img, m = r.render_frame(frame=2)
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=4, bounces=5, frame=2)
    assert same_bits(img, cimg) and np.array_equal(m[1:5], cm[1:5])
    chk.close()
    r.close()


def test_c3_full_size_properties(params):
    """BASELINE configs[2] at full size (3840x2160, 64 spp, 5 bounces) through size-independent
    properties: (a) rendering in strips equals rendering the whole frame, bit for bit (what the
    multi-GPU split relies on); (b) seeded rectangles of the frame equal the oracle port
    bit-for-bit; (c) every pixel finite, alpha 1; (d) rays = hits + misses, paths = W*H*spp."""
    sp = params
    wl = W.config3()
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=64, bounceCount=5)
    full, m = r.render_frame(frame=0)
    full = full.copy()
    assert np.isfinite(full).all() and (full[..., 3] == 1.0).all()
    assert m[sp.sp_Metric_PathsTraced] == 3840 * 2160 * 64
    # (e) EVERY pixel of the bench frame: the 2040 tile CRCs of the port's deterministic-math render
    # (tests/golden/g5_c3_full_frame.npz, tools/make_full_frame_golden.py).  A tile may differ only
    # if it holds one of the pixels where the port itself met an exact-t tie.
    assert_tiles_match(full, m, np.load(os.path.join(GOLD, "g5_c3_full_frame.npz")), "port_dm_5b", "ties_5b")
    assert m[sp.sp_Metric_RaysTraced] == m[sp.sp_Metric_RayHitCount] + m[sp.sp_Metric_RayMissCount]
    r.image[:] = 0
    total = np.zeros(12, np.uint64)
    for (b, e) in ((0, 576), (576, 1280), (1280, 2160)):
        mm, cost = r.render_rows(b, e, frame=0, want_cost=True)
        total += mm
        # per-row cost: nanoseconds of this call's kernels, spread over the rows by the work counted
        # in each (sky-kernel samples by the sky kernels' time, escaped rays and hits by the rest)
        kernel_ns = sp.last_stats().kernelMs * 1e6
        assert len(cost) == (e - 1) // 64 - b // 64 + 1 and np.all(cost > 0)
        assert abs(float(cost.sum()) - kernel_ns) <= 1e-4 * kernel_ns + len(cost)
    assert same_bits(r.image, full) and np.array_equal(total[1:5], m[1:5])
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg = np.zeros_like(full)
    rng = np.random.RandomState(8)
    rects = [(1800, 1000, 1816, 1016), (0, 0, 16, 8), (3824, 2152, 3840, 2160)]
    rects += [(x, y, x + 12, y + 12) for x, y in zip(rng.randint(1200, 2600, 5), rng.randint(400, 1700, 5))]
    for rect in rects:
        chk.render_seeded(spp=64, bounces=5, frame=0, rect=rect, image=cimg)
        x0, y0, x1, y1 = rect
        assert same_bits(full[y0:y1, x0:x1], cimg[y0:y1, x0:x1]), rect
    chk.close()
    r.close()


def test_c3_full_frame_vs_the_reference_at_its_three_bounces(gpu_sp):
    """The bench frame against the UNMODIFIED reference (fixed at 3 bounces,
    simd_path_tracer.cpp:195), every pixel: (a) deterministic-math mode vs libspref_dm: the 2040 tile
    CRCs of tests/golden/g5_c3_full_frame.npz, bit for bit (ties aside); (b) both math modes vs the
    plain reference (glibc libm) on the fixture's lattice of 32 400 pixels spread over the frame:
    stated tolerance -- at most 2 % of pixels off by more than 1e-3 relative, RMSE <= 2 % of the mean
    radiance (a last-ulp difference in a bounce direction can move a path to another texel of the
    noisy environment map); per-tile means of the whole frame within 1e-3 relative."""
    sp = gpu_sp
    gold = np.load(os.path.join(GOLD, "g5_c3_full_frame.npz"))
    wl = W.config3()
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=64, bounceCount=3, mathMode=0, renderMode=0)
    img, m = r.render_frame(frame=0)
    assert_tiles_match(img, m, gold, "ref_dm_3b", "ties_3b")
    for math_mode in (0, 1):
        sp.set_params(samplesPerPixel=64, bounceCount=3, mathMode=math_mode, renderMode=0)
        img, m = r.render_frame(frame=0)
        rep = relative_error_report(lattice(img), gold["lattice_ref_3b"])
        assert rep["fraction_above"]["0.001"] <= 0.02 and rep["rmse_over_mean"] <= 0.02, rep
        sums = img[:33 * 64, :, 0:3].astype(np.float64).reshape(33, 64, 60, 64, 3).sum(axis=(1, 3)).reshape(-1, 3)
        want = gold["tile_sum_ref_3b"][:33 * 60]
        assert np.all(np.abs(sums - want) <= 1e-3 * np.abs(want) + 1e-6), float(np.abs(sums / want - 1).max())
        assert abs(int(m[2]) - int(gold["metrics_ref_3b"][1])) <= 1e-3 * int(m[2])
    sp.set_params(samplesPerPixel=1, bounceCount=3, mathMode=0)
    r.close()


def test_c2_full_frame_fixture(gpu_sp):
    """BASELINE configs[1] against the committed fingerprint of the unmodified reference's closest
    hits (tests/golden/g6_c2_full_frame.npz: per-tile CRCs of triangle ids, distances, object ids):
    the comparison bench.py's `parity` block makes without a checker at hand."""
    sp = gpu_sp
    gold = np.load(os.path.join(GOLD, "g6_c2_full_frame.npz"))
    wl = W.config2()
    r = sp.Renderer().load_workload(wl)
    g = r.primary_hits(sample=0, frame=0)
    assert np.array_equal(tile_crcs(g["t"]), gold["tile_crc_t"]) and np.array_equal(tile_crcs(g["obj"]), gold["tile_crc_obj"])
    tiles_off = int((tile_crcs(g["tri"]) != gold["tile_crc_tri"]).sum())
    assert tiles_off <= 2 and int((g["tri"] >= 0).sum()) == int(gold["hit_pixels"])
    r.close()


def _slab_batch(sp, mins, maxs, origins, invs):
    mins = np.ascontiguousarray(mins, np.float32).reshape(-1, 12)
    maxs = np.ascontiguousarray(maxs, np.float32).reshape(-1, 12)
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    inv = np.ascontiguousarray(invs, np.float32).reshape(-1, 3)
    n = len(o)
    masks, tnear = np.zeros((n, 3), np.uint32), np.zeros((n, 4), np.float32)
    assert sp.lib.sp_b200_RayIntersectAabb4Batch(n, mins.ctypes.data, maxs.ctypes.data, o.ctypes.data, inv.ctypes.data,
                                                 masks.ctypes.data, tnear.ctypes.data) == 0
    return masks, tnear


def test_device_slab_kats(gpu_sp):
    """simd_RayIntersectAabb4 as the DEVICE evaluates it (sp_b200_RayIntersectAabb4Batch runs slab_exact,
    slab_fast and the resumable traversal's conservative slab_wide on caller data), against the reference's
    known answers -- unit_tests/test_simd_path_tracer.cpp:440-466 (masks 0xF, 0, 0 with reciprocals of
    axis-parallel directions: inf and NaN lanes), :469-485 (the recorded vector where the SSE form says hit
    and the scalar form says miss) -- and against the checker on 20 000 seeded queries including zero
    direction components, flat boxes and origins on box faces: the exact form must equal the reference's
    mask bit for bit, the hardware-min/max form must equal it whenever no reciprocal is infinite, and the
    conservative form must contain it."""
    sp = gpu_sp
    chk = best(False)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = lambda d: np.float32(1.0) / np.asarray(d, np.float32)  # noqa: E731
        mins = [(-0.5, -0.5, -0.5), (-0.5, -0.5, 1.0), (-0.5, -0.5, 2.5), (-0.5, -0.5, 4.0)]
        maxs = [(0.5, 0.5, 0.5), (0.5, 0.5, 2.0), (0.5, 0.5, 3.5), (0.5, 0.5, 5.0)]
        bmin = [(-0.375038534, 0.843911469, 0.264082730)] + [(0, 0, 0)] * 3
        bmax = [(-0.238676921, 0.916244209, 0.386187375)] + [(0, 0, 0)] * 3
        o_bug, d_bug = (-2.68516445, -1.71131170, -1.71610022), (0.576836348, 0.652352035, 0.491626590)
        masks, tnear = _slab_batch(sp, [mins, mins, mins, bmin], [maxs, maxs, maxs, bmax],
                                   [(0, 0, 10)] * 3 + [o_bug], [inv((0, 0, -1)), inv((1, 0, 0)), inv((0, 0, 1)), inv(d_bug)])
        assert masks[0, 0] == 0xF and masks[1, 0] == 0 and masks[2, 0] == 0
        assert tnear[0, 0] == np.float32(9.5)                       # :425-437
        assert masks[3, 0] & 1 == 1                                  # :469-485, the SSE answer
        # seeded sweep against the checker
        rng = np.random.RandomState(0x1A34C249 & 0x7FFFFFFF)
        n = 20000
        c = rng.uniform(-2, 2, (n, 4, 3)).astype(np.float32)
        h = (1.5 * rng.uniform(0, 1, (n, 4, 3)) ** 2).astype(np.float32)
        h[rng.rand(n, 4, 3) < 0.1] = 0.0                             # flat boxes
        lo, hi = (c - h).astype(np.float32), (c + h).astype(np.float32)
        o = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
        on_face = rng.rand(n) < 0.1
        o[on_face, 0] = lo[on_face, 0, 0]
        d = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        d[rng.rand(n, 3) < 0.05] = 0.0                               # axis-parallel components: 1/0
        d[(d == 0).all(axis=1)] = (0, 0, 1)
        d = (d / np.sqrt((d * d).sum(axis=1, dtype=np.float32))[:, None]).astype(np.float32)
        invd = inv(d)
        masks, _ = _slab_batch(sp, lo, hi, o, invd)
        want = np.array([chk.ray_aabb4(lo[i], hi[i], o[i], invd[i]) for i in range(n)], np.uint32)
    assert np.array_equal(masks[:, 0], want)
    finite = np.isfinite(invd).all(axis=1)
    assert finite.sum() > 15000 and np.array_equal(masks[finite, 1], want[finite])
    machine = masks[:, 2] != 0xFFFFFFFF
    assert machine.sum() > 15000 and not np.any(want[machine] & ~masks[machine, 2])
    assert (want != 0).sum() > 500


def test_metrics_mesh_test_counters(gpu_sp):
    """sp_Metrics slots 10 and 11 (sp_metrics.h:33-37).  TestsPerformed (sp_scene.cpp:286: one per object
    whose broadphase leaf the ray passes) does not depend on the tree: with distance culling off it must
    EQUAL the reference's count, per frame, in both schedulers.  MidphaseAabbTestCount (sp_scene.cpp:154)
    sums the children of the inner nodes the REFERENCE's tree visits (bvh.cpp:254) and so belongs to that
    tree; the library reports 4 x the 4-wide nodes it fetched (its own box tests), filled only when
    sp_b200_EnableStats is on: checked for consistency with sp_b200_Stats and between the schedulers."""
    sp = gpu_sp
    wl = W.multi_object_workload(width=160, height=120, spp=2, env_size=(256, 128))
    r = sp.Renderer().load_workload(wl)
    chk = best(True).scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=2, bounces=3, frame=3)
    sp.lib.sp_b200_EnableStats(1)
    # (every ray must really be traced for the count to be the reference's: the coverage pass settles
    # camera rays that pass an object's box but none of its triangles without entering the object)
    sp.lib.sp_b200_SetSkyCulling(0)
    seen = []
    for mode in (0, 1):
        sp.set_params(samplesPerPixel=2, bounceCount=3, cullByDistance=0, renderMode=mode)
        img, m = r.render_frame(frame=3)
        st = sp.last_stats()
        assert same_bits(img, cimg) and np.array_equal(m[1:5], cm[1:5])
        assert int(m[11]) == int(cm[11]) > 0                          # mesh tests: the reference's number
        assert int(m[10]) == 4 * int(st.nodeVisits) > 0 and int(st.objectTests) == int(m[11])
        seen.append((int(m[10]), int(m[11])))
    assert seen[0][1] == seen[1][1]
    sp.lib.sp_b200_SetSkyCulling(2)
    sp.lib.sp_b200_EnableStats(0)
    sp.set_params(samplesPerPixel=1, bounceCount=3, cullByDistance=1, renderMode=0)
    chk.close()
    r.close()


def test_watertight_option(gpu_sp):
    """sp_b200_Params::triangleTest = WATERTIGHT on the device (per-ray query kernel and both render
    schedulers).  (a) 200 000 rays aimed exactly at shared edges / vertices of a sheet: with the default
    test the GPU leaks exactly the rays the checker's Moller-Trumbore leaks (t bit for bit); with the
    watertight test it leaks none.  (b) an image of the sheet seen from above: no pixel of the sheet's
    interior may show the sky in watertight mode, whichever renderMode is set (the option renders with
    the per-pixel kernel)."""
    sp = gpu_sp
    mesh, o, d = W.crack_test_inputs()
    chk = best(False).scene()
    chk.add_mesh(mesh.vertices, mesh.indices, False)
    chk.add_object(0, W.MATERIAL_SURFACE)
    chk.build()
    want = chk.intersect_rays(o, d)
    chk.close()
    r = sp.Renderer()
    r.add_object(r.add_mesh(mesh.vertices, mesh.indices, False), W.MATERIAL_SURFACE)
    r.build()
    sp.set_params(triangleTest=sp.TRIANGLE_MOLLER_TRUMBORE)
    got = r.intersect_rays(o, d)
    assert same_bits(got["t"], want["t"]) and int((got["t"] < 0).sum()) > 0      # parity includes the cracks
    sp.set_params(triangleTest=sp.TRIANGLE_WATERTIGHT)
    wt = r.intersect_rays(o, d)
    assert int((wt["t"] < 0).sum()) == 0
    both = got["t"] > 0
    assert np.all(np.abs(wt["t"][both] - got["t"][both]) <= 1e-5 * np.abs(got["t"][both]))
    # (b) every pixel whose centre ray points into the sheet's interior must hit it
    for m_id, kw in ((W.MATERIAL_BACKGROUND, dict(emission=(0.0, 0.0, 0.0))), (W.MATERIAL_SURFACE, dict(albedo=(0, 0, 0), emission=(1.0, 1.0, 1.0), roughness=0.5))):
        r.register_material(m_id, **kw)
    r.set_background(W.MATERIAL_BACKGROUND)
    r.configure_camera((0.0, 0.0, 2.0), (0.0, 0.0, 0.0, 1.0), 0.8, 256, 256)
    images = []
    for mode in (0, 1):
        sp.set_params(samplesPerPixel=4, bounceCount=1, renderMode=mode, triangleTest=sp.TRIANGLE_WATERTIGHT)
        img, m = r.render_frame(frame=1)
        images.append(img.copy())
        inner = img[64:192, 64:192, 0]                 # well inside the sheet's projection
        assert np.all(inner == 1.0), int((inner != 1.0).sum())
    assert same_bits(images[0], images[1])
    sp.set_params(samplesPerPixel=1, bounceCount=3, renderMode=0, triangleTest=sp.TRIANGLE_MOLLER_TRUMBORE)
    r.close()


def test_progressive_accumulation(gpu_sp):
    """sp_b200_AccumulateFrame: the running mean of frames (the reference's own to-do, main.cpp:75).
    Six frames of 2 spp folded in on the device equal the same recurrence in numpy float32, bit for bit,
    and approach the 12-spp-per-pixel mean."""
    import torch
    sp = gpu_sp
    wl = W.config1(160, 120, env_size=(256, 128))
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=2, bounceCount=3)
    n = wl.width * wl.height
    accum = torch.zeros((wl.height, wl.width, 4), dtype=torch.float32, device="cuda:0")
    ref = None
    out = np.zeros((wl.height, wl.width, 4), np.float32)
    for f in range(6):
        img, _ = r.render_frame(frame=f)
        img = img.copy()
        assert sp.lib.sp_b200_AccumulateFrame(accum.data_ptr(), None, img.ctypes.data, n, f, out.ctypes.data) == 0
        if f == 0:
            ref = img.copy()
        else:
            ref[..., 0:3] = ref[..., 0:3] + (img[..., 0:3] - ref[..., 0:3]) / np.float32(f + 1)
            ref[..., 3] = img[..., 3]
        assert same_bits(out, ref), f
    assert same_bits(accum.cpu().numpy(), ref)
    sp.set_params(samplesPerPixel=1, bounceCount=3)
    r.close()


def test_pipelined_frames_begin_end(gpu_sp):
    """sp_b200_RenderRowsBegin / End: two frames in flight (begin k + 1 before ending k) give the bits, metrics
    and per-row cost totals of the same frames rendered one call at a time -- into pinned host images (rows
    streamed band by band), into caller-owned device images, with the strip moved between frames (the cached
    coverage no longer applies) and with the scene rebuilt in between; a third Begin is refused."""
    import torch
    sp = gpu_sp
    wl = W.config1(320, 200, env_size=(256, 128))
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=6, bounceCount=4, renderMode=0, tileHeight=8, tileWidth=64)
    H, Wd = wl.height, wl.width
    strips = [(0, H), (0, H), (40, 160), (40, 160), (0, 96), (0, H)]
    want = []
    for f, (b, e) in enumerate(strips):
        r.image[...] = 0
        m, cost = r.render_rows(b, e, frame=f, host=True, want_cost=True)
        want.append((r.image.copy(), m, cost))
    hosts = [torch.zeros((H, Wd, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    devs = [torch.zeros((H, Wd, 4), dtype=torch.float32, device="cuda:0") for _ in range(2)]
    for mode in ("host", "device", "host_rebuild"):
        got = [None] * len(strips)
        slots = {}
        def end(f):
            m, cost = r.render_rows_end(slots.pop(f))
            img = hosts[f % 2].numpy().copy() if mode != "device" else devs[f % 2].cpu().numpy()
            got[f] = (img, m, cost)
        for f, (b, e) in enumerate(strips):
            if mode == "device":
                devs[f % 2].zero_()
                torch.cuda.synchronize()
                slots[f] = r.render_rows_begin(b, e, frame=f, host=False, device_ptr=devs[f % 2].data_ptr(), want_cost=True)
            else:
                hosts[f % 2].zero_()
                if mode == "host_rebuild" and f in (2, 5):
                    r.build()
                slots[f] = r.render_rows_begin(b, e, frame=f, host_ptr=hosts[f % 2].data_ptr(), want_cost=True)
            if f == 1:
                with pytest.raises(RuntimeError):
                    r.render_rows_begin(b, e, frame=f, host_ptr=hosts[0].data_ptr())
            if f >= 1:
                end(f - 1)
        end(len(strips) - 1)
        for f, (img, m, cost) in enumerate(got):
            assert same_bits(img, want[f][0]), (mode, f)
            assert np.array_equal(m[1:5], want[f][1][1:5]), (mode, f)
            assert cost.shape == want[f][2].shape and (cost > 0).sum() == (want[f][2] > 0).sum(), (mode, f)
    # the one-call form still works with a frame in flight, and afterwards
    s0 = r.render_rows_begin(0, H, frame=1, host_ptr=hosts[0].data_ptr())
    r.image[...] = 0
    m, _ = r.render_rows(0, H, frame=0, host=True)
    assert same_bits(r.image, want[0][0])
    r.render_rows_end(s0)
    assert same_bits(hosts[0].numpy(), want[1][0])
    # an empty strip is a frame too: a slot to end, nothing rendered, nothing counted
    hosts[0].zero_()
    se = r.render_rows_begin(48, 48, frame=0, host_ptr=hosts[0].data_ptr())
    me, _ = r.render_rows_end(se)
    assert not me.any() and not hosts[0].numpy().any()
    # a frame begun on one thread and ended on another (what Begin leaves for End lives with the library)
    import threading
    hosts[1].zero_()
    s1 = r.render_rows_begin(0, H, frame=5, host_ptr=hosts[1].data_ptr())
    ended = []
    t = threading.Thread(target=lambda: ended.append(r.render_rows_end(s1)))
    t.start()
    t.join()
    assert len(ended) == 1 and np.array_equal(ended[0][0][1:5], want[5][1][1:5]) and same_bits(hosts[1].numpy(), want[5][0])
    sp.set_params(samplesPerPixel=1, bounceCount=3, tileHeight=64)
    r.close()


def test_reference_perf_test_queries(gpu_sp):
    """The queries of the reference's own performance tests (perf_tests/perf_tests.cpp:51-118 TestBvh,
    :212-305 TestMeshMidphase) with the reference's seeded inputs restated draw for draw
    (workloads.perf_bvh_inputs / perf_mesh_inputs), batched on the GPU: per ray the SET of intersected
    leaves (count, xor and sum of the leaf indices) must equal bvh_IntersectRay's, and sp_RayIntersectMesh's
    t must match bit for bit, over the reference compiled unmodified."""
    sp = gpu_sp
    chk = best(False)
    r = sp.Renderer()
    # the reference's input as it is (half its boxes have min > max: which of those leaves the reference
    # reports depends on its tree, see perf_bvh_inputs): the GPU's leaf set contains the reference's
    mn, mx, o, d = W.perf_bvh_inputs(sp.xorshift_bilateral_stream(0x1A34C249))
    assert mn.shape == (2048, 3) and o.shape == (8192, 3)
    want, build_s, query_s = chk.perf_bvh(mn, mx, o, d, 2048)
    boxes = W.boxes_as_triangles(mn, mx)
    got, ms = sp.mesh_leaves_batch(r.meshes[r.add_mesh(boxes.vertices, boxes.indices, False)], o, d)
    assert np.all(got[:, 0] >= want[:, 0]) and want[:, 0].max() > 20
    # the same draws with |radius|: well-defined leaf sets, which must be EQUAL
    mn, mx, o, d = W.perf_bvh_inputs(sp.xorshift_bilateral_stream(0x1A34C249), proper=True)
    want_p, _, _ = chk.perf_bvh(mn, mx, o, d, 2048)
    boxes = W.boxes_as_triangles(mn, mx)
    got_p, _ = sp.mesh_leaves_batch(r.meshes[r.add_mesh(boxes.vertices, boxes.indices, False)], o, d)
    assert np.array_equal(got_p, want_p) and want_p[:, 0].max() > 20
    mesh, o2, d2 = W.perf_mesh_inputs(sp.xorshift_bilateral_stream(0x1A34C249), 1 << 18)
    s = chk.scene()
    s.add_mesh(mesh.vertices, mesh.indices, False)
    t_want, _tri, secs = s.perf_mesh(0, o2, d2)
    s.close()
    m2 = r.add_mesh(mesh.vertices, mesh.indices, False)
    t_got, tri_got, ms2 = sp.mesh_intersect_batch(r.meshes[m2], o2, d2)
    assert same_bits(t_got, t_want) and (t_want >= 0).sum() > 10000
    timing = os.environ.get("SPB_TIMING_OUT")
    if timing:
        with open(timing, "a") as f:
            f.write(f"PERF_TESTS TestBvh 8192 rays x 2048 boxes: GPU kernel {ms * 1e3:.1f} us, {chk.name} query loop {query_s * 1e3:.2f} ms "
                    f"(tree build {build_s:.2f} s); TestMeshMidphase {len(o2)} rays: GPU kernel {ms2:.3f} ms, {chk.name} {secs * 1e3:.1f} ms\n")
    r.close()


def test_reference_unit_tests_against_the_library(gpu_sp):
    """INTEGRATION.md §1 exercised: tests/dropin/_build/dropin_unit_tests is ONE translation unit that
    includes the reference's own headers (all types), include/sp_b200.h with
    SP_B200_USE_REFERENCE_TYPES, and the reference's own Unity test functions -- extracted unmodified
    from unit_tests/test_simd_path_tracer.cpp:49-395 at build time -- and links libspb200.so where the
    reference's test links bvh.cpp / sp_scene.cpp / sp_material_system.cpp / simd_path_tracer.cpp.  Every
    test that passes against the reference's sources (tests/test_abi.py runs that binary) must pass here
    on the GPU; plus the reference's frame loop -- its inline WorkQueue, 16 WorkerThreads calling
    sp_PathTraceTile (main.cpp:728-759) -- whose image must equal the one-thread image while the
    concurrent callers share launches.  Built where the reference is mounted; the binary travels."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin", "_build", "dropin_unit_tests")
    if not os.path.exists(exe):
        pytest.skip("tests/dropin/_build not built (the reference checkout was never mounted here)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert "9 Tests 0 Failures 0 Ignored" in out, out[-3000:]
    for name in ("TestPathTraceSingleColor", "TestPathTraceTile", "TestConfigureCamera", "TestCalculateFilmP",
                 "TestRayIntersectScene", "TestMaterialAlbedoTexture", "TestMetrics", "TestRayIntersectMesh",
                 "TestWorkerThreadsThroughTheLibrary"):
        assert f"{name}:PASS" in out, name
    timing = os.environ.get("SPB_TIMING_OUT")
    if timing:
        with open(timing, "a") as f:
            f.write([l for l in out.splitlines() if l.startswith("WORKER_THREADS")][0] + "\n")


@pytest.mark.parametrize("mode", ["host", "device"])
def test_multi_device_frame(gpu_sp, mode):
    """Several devices behind the C ABI: sp_b200_InitDeviceList, then
    sp_b200_RenderFrame (every device copies its rows to the host image) or
    sp_b200_RenderFrameToDevice (strips gathered on the primary with peer copies).  The frame must be
    bit-identical to the one-device frame and the counters equal, for every frame while the strip
    boundaries are re-cut from the measured cost.  Uses every GPU of the box; on a one-GPU box the
    same device is listed three times, which exercises the same code (own worker, streams and buffers
    per entry).  Own process: the device set of a process is chosen once."""
    import json
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    devices = list(range(min(n, 4))) if n > 1 else [0, 0, 0]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_multi_device_worker.py")
    p = subprocess.run([sys.executable, worker, ",".join(map(str, devices)), mode], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert all(out["identical"]) and all(out["metrics_equal"]), out
    first, last = out["strips"][0], out["strips"][-1]
    assert first[0][0] == 0 and first[-1][1] == 368 and all(a[1] == b[0] for a, b in zip(first, first[1:]))
    assert last[0][0] == 0 and last[-1][1] == 368 and all(a[1] == b[0] for a, b in zip(last, last[1:]))
    assert first != last                       # the cut moved: this camera makes an even split uneven


def test_c5_instanced_scene(params):
    """BASELINE configs[4] shape (182 objects, 9.98 M instanced triangles; reduced resolution and
    sample count so the CPU checker finishes in seconds): objects beyond the reference's table of
    32 go through sp_b200_AddObjectToScene.  Checker: the port (the reference cannot hold the
    scene).  Bit-exact image, object ids, hit distances and triangle ids, except where the checker
    reports an exact-t tie between two candidates (bounded at 1e-4 of the pixels)."""
    sp = params
    wl = W.config5(480, 270, spp=2, bounces=5, env_size=(512, 256))
    r = sp.Renderer().load_workload(wl)
    assert sp.lib.sp_b200_SceneObjectCount(C.byref(r.scene)) == 182 and r.scene.objectCount == 32
    sp.set_params(samplesPerPixel=2, bounceCount=5)
    img, m = r.render_frame(frame=4)
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=2, bounces=5, frame=4)
    # tiny sphere triangles: a few rays per frame meet two edge-sharing triangles at bit-equal t
    ntie = assert_same_image_up_to_ties(img, cimg, chk, 2, 5, 4)
    assert m[1] == cm[1] and np.all(np.abs(m[2:5].astype(np.int64) - cm[2:5].astype(np.int64)) <= ntie * 2 * 5)
    g, e = r.primary_hits(), chk.primary_hits()
    # closest-hit ids: equal except equal-t ties (t itself must be bit-equal everywhere)
    assert np.array_equal(g["obj"], e["obj"]) and same_bits(g["t"], e["t"])
    assert (g["tri"] != e["tri"]).sum() <= 1e-4 * g["tri"].size
    assert len(np.unique(e["obj"])) > 100
    # straggler eviction in a multi-object scene: continuations parked in the TLAS and inside objects
    img = img.copy()
    for evict in ((12, 12), (32, 20)):
        sp.lib.sp_b200_SetStragglerEviction(*evict)
        img2, m2 = r.render_frame(frame=4)
        assert same_bits(img2, img) and np.array_equal(m2[1:5], m[1:5]), evict
    sp.lib.sp_b200_SetStragglerEviction(0, 0)
    chk.close()
    r.close()


def test_c5_instanced_scene_1080p(gpu_sp):
    """BASELINE configs[4] at 1920x1080, 4 spp, 5 bounces ("C5 parity runs only at 480x270x2"):
    8.3 M paths through 182 instances / 9.98 M instanced triangles, the production (wavefront) scheduler
    against the port in deterministic-math mode -- image bit for bit except at the port's own exact-t tie
    pixels (bounded at 1e-4), object ids and hit distances exact, triangle ids up to ties."""
    sp = gpu_sp
    wl = W.config5(1920, 1080, spp=4, bounces=5, env_size=(1024, 512))
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=4, bounceCount=5, renderMode=0)
    img, m = r.render_frame(frame=2)
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=4, bounces=5, frame=2)
    ntie = assert_same_image_up_to_ties(img, cimg, chk, 4, 5, 2)
    assert m[1] == cm[1] and np.all(np.abs(m[2:5].astype(np.int64) - cm[2:5].astype(np.int64)) <= max(1, ntie) * 4 * 5)
    g, e = r.primary_hits(), chk.primary_hits()
    assert np.array_equal(g["obj"], e["obj"]) and same_bits(g["t"], e["t"])
    assert (g["tri"] != e["tri"]).sum() <= 1e-4 * g["tri"].size
    assert len(np.unique(e["obj"])) > 120
    sp.set_params(samplesPerPixel=1, bounceCount=3)
    chk.close()
    r.close()


@pytest.mark.parametrize("case", ["close_up_inside_bounds", "looking_away", "off_centre_ragged_70spp",
                                  "behind_and_beside"])
def test_coverage_mask_and_ray_sorting_edge_cases(gpu_sp, case):
    """The coverage pass (which pixels can be hit at all), the sky kernel and the direction-sorted
    tiles must never change a bit.  Cameras that stress the projection: inside the mesh's bounds
    (triangles straddle the camera plane: the "everything" flag), looking away from the mesh (every
    pixel is sky), the mesh partly off screen on a ragged image size with a sample count that does
    not divide the 2048-item tiles, and the mesh behind / beside the camera.  Checker: the oracle
    (port or reference, deterministic math) on the whole image, bit for bit; counters equal."""
    sp = gpu_sp
    spp, bounces, size = 3, 4, (203, 149)
    wl = W.config1(size[0], size[1], env_size=(256, 128))
    px, py, pz = wl.camera_position
    if case == "close_up_inside_bounds":
        lo, hi = W.load_mesh("bunny").bounds()
        wl.camera_position = (float((lo[0] + hi[0]) / 2), float(hi[1] * 0.9 + lo[1] * 0.1), float(hi[2]))
    elif case == "looking_away":
        wl.camera_rotation = W.quat_axis_angle((0, 1, 0), np.pi)
    elif case == "off_centre_ragged_70spp":
        spp = 70
        wl.camera_position = (px + 0.08, py - 0.05, pz * 0.8)
    else:
        # the mesh at the edge of the view, partly outside it, a quarter of it beside the camera
        wl.camera_rotation = W.quat_axis_angle((0, 1, 0), np.pi * 0.15)
        wl.camera_position = (px + 0.05, py, pz * 0.35)
    wl.spp, wl.bounces = spp, bounces
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=spp, bounceCount=bounces, radianceClamp=10.0, envFilter=0, mathMode=0,
                  cullByDistance=1, tileWidth=64, tileHeight=64, renderMode=0, samplesPerPass=0)
    img, m = r.render_frame(frame=5)
    img = img.copy()
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=spp, bounces=bounces, frame=5)
    ntie = assert_same_image_up_to_ties(img, cimg, chk, spp, bounces, 5)
    assert m[1] == cm[1] and np.all(np.abs(m[2:5].astype(np.int64) - cm[2:5].astype(np.int64)) <= ntie * spp * bounces)
    if case == "looking_away":
        assert m[3] == 0 and m[4] == size[0] * size[1] * spp
    else:
        assert m[3] > 0
    # the same frame with every scheduling feature off, small passes, and through the per-pixel kernel
    sp.lib.sp_b200_SetSkyCulling(0)
    sp.lib.sp_b200_SetRaySorting(0)
    sp.lib.sp_b200_SetPrimaryCandidates(0)
    sp.lib.sp_b200_SetPathsPerPass(50000)
    img2, m2 = r.render_frame(frame=5)
    assert same_bits(img2, img) and np.array_equal(m2[1:5], m[1:5])
    sp.lib.sp_b200_SetSkyCulling(1)     # sky kernel with the per-sample loop for every pixel
    sp.lib.sp_b200_SetRaySorting(1)
    sp.lib.sp_b200_SetPrimaryCandidates(1)
    img2, m2 = r.render_frame(frame=5)
    assert same_bits(img2, img) and np.array_equal(m2[1:5], m[1:5])
    sp.lib.sp_b200_SetSkyCulling(2)     # default: one lookup where provably enough
    for thresholds in ((1, 1, 1), (12, 12, 12), (1, 0, 12)):
        sp.lib.sp_b200_SetRefillThresholds(*thresholds)
        img2, m2 = r.render_frame(frame=5)
        assert same_bits(img2, img) and np.array_equal(m2[1:5], m[1:5])
    # straggler eviction (continuation records + a second launch of the trace kernel): same walk, same bits
    for evict in ((12, 0), (0, 12), (16, 16), (32, 32), (3, 5)):
        sp.lib.sp_b200_SetStragglerEviction(*evict)
        img2, m2 = r.render_frame(frame=5)
        assert same_bits(img2, img) and np.array_equal(m2[1:5], m[1:5]), evict
    sp.lib.sp_b200_SetStragglerEviction(0, 0)
    sp.lib.sp_b200_SetPathsPerPass(0)
    sp.set_params(renderMode=1)
    img2, m2 = r.render_frame(frame=5)
    assert same_bits(img2, img) and np.array_equal(m2[1:5], m[1:5])
    sp.set_params(renderMode=0, samplesPerPixel=1, bounceCount=3)
    chk.close()
    r.close()


def test_miss_fusion_gives_the_same_bits(gpu_sp):
    """sp_b200_SetMissFusion: escaped rays shaded by the trace kernel where it retires them (mode 1) give the
    image, the path / ray / hit / miss counters and the per-row cost units of the miss queue + k_shade_miss
    (mode 0): a single-object scene (packet mode, straggler eviction, candidate lists) and a multi-object one
    (four-class vote), nearest and bilinear environment filter, sky culling on and off, and the reference's
    fixture itself with the fusion on."""
    sp = gpu_sp
    try:
        for wl, spp, bounces in ((W.config1(200, 150, env_size=(256, 128)), 7, 5),
                                 (W.multi_object_workload(width=160, height=120, spp=4, env_size=(256, 128)), 4, 4)):
            r = sp.Renderer().load_workload(wl)
            for env_filter in (0, 1):
                for sky in (2, 0):
                    sp.lib.sp_b200_SetSkyCulling(sky)
                    sp.set_params(samplesPerPixel=spp, bounceCount=bounces, renderMode=0, mathMode=0, envFilter=env_filter,
                                  cullByDistance=1, radianceClamp=10.0, tileHeight=8, tileWidth=64)
                    want = None
                    for mode in (0, 1):
                        sp.lib.sp_b200_SetMissFusion(mode)
                        r.image[...] = 0
                        m, cost = r.render_rows(0, wl.height, frame=2, host=True, want_cost=True)
                        got = (r.image.copy(), m[1:5].copy(), cost > 0)
                        if want is None:
                            want = got
                        assert same_bits(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), (env_filter, sky, mode)
            r.close()
        sp.lib.sp_b200_SetSkyCulling(2)
        sp.lib.sp_b200_SetMissFusion(1)
        g = np.load(os.path.join(GOLD, "g1_bunny_96x64.npz"))
        sp.set_params(samplesPerPixel=2, bounceCount=3, mathMode=0, envFilter=0, renderMode=0, cullByDistance=1, tileHeight=64)
        r = sp.Renderer().load_workload(W.config1(96, 64, env_size=(512, 256)))
        img, m = r.render_frame(frame=1)
        assert same_bits(img, g["image_ref_dm"]) and np.array_equal(m[1:5], g["metrics_ref_dm"])
        r.close()
    finally:
        sp.lib.sp_b200_SetMissFusion(-1)
        sp.lib.sp_b200_SetSkyCulling(2)
        sp.set_params(samplesPerPixel=1, bounceCount=3, envFilter=0, tileHeight=64)


def test_wavefront_pass_split_and_stats(gpu_sp):
    """Wavefront scheduling details: any samples-per-pass split gives the same bits (the sample
    order of the accumulation is kept, simd_path_tracer.cpp:321); the stats launch counts the same
    rays; per-tile-row cost sums to the rays traced."""
    sp = gpu_sp
    wl = W.config1(200, 150, env_size=(256, 128))
    r = sp.Renderer().load_workload(wl)
    ref_img = None
    for spass in (0, 1, 3, 7):
        sp.set_params(samplesPerPixel=7, bounceCount=4, renderMode=0, samplesPerPass=spass, mathMode=0,
                      envFilter=0, cullByDistance=1, radianceClamp=10.0)
        img, m = r.render_frame(frame=3)
        if ref_img is None:
            ref_img, ref_m = img.copy(), m.copy()
        assert same_bits(img, ref_img) and np.array_equal(m[1:5], ref_m[1:5])
    # bands of rows: 200 x 150 at 7 spp in passes of at most 4096 / 20000 paths (several bands,
    # the last one ragged; with 4096 the samples of a pixel are split over passes as well)
    for paths in (4096, 20000):
        sp.lib.sp_b200_SetPathsPerPass(paths)
        sp.set_params(samplesPerPass=0)
        img, m = r.render_frame(frame=3)
        assert same_bits(img, ref_img) and np.array_equal(m[1:5], ref_m[1:5])
    sp.lib.sp_b200_SetPathsPerPass(0)
    # sky culling off: every pixel through the queues -- same bits, same counters, same row costs
    sp.lib.sp_b200_SetSkyCulling(0)
    img, m = r.render_frame(frame=3)
    assert same_bits(img, ref_img) and np.array_equal(m[1:5], ref_m[1:5])
    _, cost_all = r.render_rows(0, 150, frame=3, want_cost=True)
    # row costs are nanoseconds of the call's kernels (sky-kernel work by the sky kernels' time, queue
    # work by the rest): they add up to the kernel time whichever kernels did the work
    ns = sp.last_stats().kernelMs * 1e6
    assert len(cost_all) == 3 and np.all(cost_all > 0) and abs(float(cost_all.sum()) - ns) <= 1e-4 * ns + 3
    sp.lib.sp_b200_SetSkyCulling(2)
    _, cost_sky = r.render_rows(0, 150, frame=3, want_cost=True)
    ns = sp.last_stats().kernelMs * 1e6
    assert len(cost_sky) == 3 and np.all(cost_sky > 0) and abs(float(cost_sky.sum()) - ns) <= 1e-4 * ns + 3
    sp.set_params(renderMode=1)
    img, m = r.render_frame(frame=3)
    assert same_bits(img, ref_img) and np.array_equal(m[1:5], ref_m[1:5])
    sp.set_params(renderMode=0, samplesPerPass=2)
    sp.lib.sp_b200_EnableStats(1)
    mm, cost = r.render_rows(0, 150, frame=3, want_cost=True)
    st = sp.last_stats()
    sp.lib.sp_b200_EnableStats(0)
    assert int(st.rays) == int(ref_m[2]) and st.nodeVisits > st.rays and st.triangleTests > 0
    assert 0 < int(cost.sum()) and len(cost) == 3
    assert same_bits(r.image, ref_img)
    sp.set_params(samplesPerPixel=1, bounceCount=3, samplesPerPass=0)
    r.close()


def test_tone_map_matches_oracle(gpu_sp):
    """sp_b200_ToneMap (output stage, post_processing.frag.glsl:19-26) against the port's restatement:
    identical bytes on a rendered frame, on a radiance ramp and on special values, in both math
    modes' own pairing (mathMode 0 <-> deterministic-math port)."""
    sp = gpu_sp
    wl = W.config1(160, 120, env_size=(256, 128))
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=2, bounceCount=3, mathMode=0, renderMode=0)
    img, _ = r.render_frame(frame=1)
    rng = np.random.RandomState(3)
    ramp = np.concatenate([np.linspace(0, 20, 20000), rng.lognormal(0, 2, 20000), [0, -1, np.nan, np.inf, 1e30]]).astype(np.float32)
    synth = np.stack([ramp, ramp[::-1], ramp * 0.25, np.ones_like(ramp)], axis=-1)
    port = ora.load_port_dm()
    for exposure in (1.0, 0.37):
        assert np.array_equal(sp.tone_map(img, exposure), port.tone_map(img, exposure))
        assert np.array_equal(sp.tone_map(synth, exposure), port.tone_map(synth, exposure))
    r.close()
    sp.set_params(samplesPerPixel=1, bounceCount=3)


def test_assets_to_pixels_end_to_end(gpu_sp, tmp_path):
    """The rows either side of the path plugged together through the C ABI only: a mesh read by
    sp_b200_LoadObj and an environment map read by LoadExrImage (fixture files) are rendered, tone
    mapped and written as PPM.  The oracle port gets the same arrays: the image must be bit-identical
    (deterministic math) and the RGBA8 bytes equal."""
    sp = gpu_sp
    here = os.path.dirname(os.path.abspath(__file__))
    vertices, indices = sp.load_obj(os.path.join(here, "golden", "quad_mix.obj"))
    env = sp.load_exr(os.path.join(here, "golden", "exr", "zip_f32_rgb_70x45.exr"))
    assert vertices is not None and env is not None and env.shape == (45, 70, 4)
    mesh = W.Mesh(vertices, indices, smooth=False)
    lo, hi = mesh.bounds()
    wl = W.Workload(
        name="assets_end_to_end", meshes=[mesh],
        objects=[W.SceneObject(0, W.MATERIAL_SURFACE, rotation=W.quat_axis_angle((0.2, 1.0, 0.0), 0.6))],
        materials=[W.Material(W.MATERIAL_BACKGROUND, emission_texture=W.IMAGE_ENV),
                   W.Material(W.MATERIAL_SURFACE, albedo=(0.6, 0.5, 0.4), roughness=0.5)],
        textures={W.IMAGE_ENV: env}, background=W.MATERIAL_BACKGROUND,
        camera_position=W.frame_camera(lo, hi, 1.6), camera_rotation=(0.0, 0.0, 0.0, 1.0), film_distance=0.8,
        width=96, height=72, spp=4, bounces=4)
    r = sp.Renderer().load_workload(wl)
    sp.set_params(samplesPerPixel=4, bounceCount=4, radianceClamp=10.0, envFilter=0, mathMode=0, cullByDistance=1,
                  tileWidth=64, tileHeight=64, renderMode=0, samplesPerPass=0)
    img, m = r.render_frame(frame=9)
    chk = ora.load_port_dm().scene().load_workload(wl)
    cimg, cm = chk.render_seeded(spp=4, bounces=4, frame=9)
    ntie = assert_same_image_up_to_ties(img, cimg, chk, 4, 4, 9)
    assert m[3] > 0 and m[4] > 0 and (ntie > 0 or np.array_equal(m[1:5], cm[1:5]))
    rgba8 = sp.tone_map(img)
    if ntie == 0:
        assert np.array_equal(rgba8, ora.load_port_dm().tone_map(cimg))
    out = tmp_path / "frame.ppm"
    sp.write_ppm(str(out), rgba8)
    assert out.stat().st_size == len(b"P6\n96 72\n255\n") + 96 * 72 * 3
    chk.close()
    r.close()
    sp.set_params(samplesPerPixel=1, bounceCount=3)


# ---- environment pre-processing (src/cubemap.cpp) --------------------------

def test_cube_map_and_irradiance_match_oracle(gpu_sp):
    """sp_b200_CreateCubeMap / sp_b200_CreateIrradianceCubeMap (k_cube_map, k_irradiance) against the
    reference's CreateCubeMap / CreateIrradianceCubeMap (cubemap.cpp:108-291, compiled unmodified in
    oracle/_ref; both sampling branches) in deterministic-math mode: every float of every face bit
    for bit.  Ragged and 1x1 faces, sample counts that are not powers of two (the random branch's
    serial XorShift32 stream is reached per texel by a GF(2) jump on the device), the committed
    fixture made from the reference, and other sampleDelta values against the port."""
    sp = gpu_sp
    sp.set_params(mathMode=0)
    chk = best(dm=True)
    env = W.make_env_map(256, 128)
    for (w, h, spp) in ((7, 9, 13), (1, 1, 1), (16, 16, 32), (33, 5, 100)):
        assert same_bits(sp.create_cube_map(env, 3 * w, 2 * h), chk.create_cube_map(env, 3 * w, 2 * h)), (w, h)
        for sampling in (sp.IRRADIANCE_UNIFORM, sp.IRRADIANCE_RANDOM):
            got = sp.create_irradiance_cube_map(env, w, h, spp=spp, sampling=sampling)
            want = chk.create_irradiance_cube_map(env, w, h, spp=spp, sampling=sampling)
            assert same_bits(got, want), (w, h, spp, sampling)
    # more than one shared-memory chunk of terms per texel (SPB_IRR_CHUNK = 1024): 2500 random
    # samples, and a finer uniform grid (sampleDelta 0.05: 126 x 32 = 4032 terms) against the port
    port = ora.load_port_dm()
    assert same_bits(sp.create_irradiance_cube_map(env, 3, 2, spp=2500, sampling=sp.IRRADIANCE_RANDOM),
                     chk.create_irradiance_cube_map(env, 3, 2, spp=2500, sampling=1))
    for delta in (0.05, 0.25, 3.0):
        assert same_bits(sp.create_irradiance_cube_map(env, 3, 2, sample_delta=delta),
                         port.create_irradiance_cube_map(env, 3, 2, sample_delta=delta)), delta
    g = np.load(os.path.join(GOLD, "g4_cubemap.npz"))
    genv = W.make_env_map(int(g["env_params"][0]), int(g["env_params"][1]), str(g["env_variant"]))
    cw, ch = (int(v) for v in g["cube_size"])
    iw, ih = (int(v) for v in g["irradiance_size"])
    assert same_bits(sp.create_cube_map(genv, cw, ch), g["cube_dm"])
    assert same_bits(sp.create_irradiance_cube_map(genv, iw, ih), g["irradiance_uniform_dm"])
    assert same_bits(sp.create_irradiance_cube_map(genv, iw, ih, spp=int(g["spp"]), sampling=sp.IRRADIANCE_RANDOM),
                     g["irradiance_random_dm"])


def _time_bakers(sp, env):
    """With SPB_TIMING_OUT set: wall time of the two bakers at the reference's sizes with the map
    already resident on the device and the faces left there (launch + kernel + synchronize; best
    of 5), appended to that file for profiles/.  Not a test."""
    path = os.environ.get("SPB_TIMING_OUT")
    if not path:
        return
    import time
    import torch
    img = np.ascontiguousarray(env, dtype=np.float32)
    hdr = sp.HdrImage(img.ctypes.data_as(C.POINTER(C.c_float)), img.shape[1], img.shape[0])
    faces = torch.empty(6 * 1024 * 1024 * 4, dtype=torch.float32, device="cuda")
    lines = []
    for name, call in (
            ("cube_map 4096x2048 -> 6x1024x1024", lambda: sp.lib.sp_b200_CreateCubeMap(C.byref(hdr), 1024, 1024, None, faces.data_ptr())),
            ("irradiance uniform 6x32x32 (1008 terms/texel)", lambda: sp.lib.sp_b200_CreateIrradianceCubeMap(C.byref(hdr), 32, 32, 32, 0, 0.1, None, faces.data_ptr())),
            ("irradiance random 6x32x32 spp 32", lambda: sp.lib.sp_b200_CreateIrradianceCubeMap(C.byref(hdr), 32, 32, 32, 1, 0.1, None, faces.data_ptr()))):
        call()  # uploads the map on first use
        best_s = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            call()
            best_s = min(best_s, time.perf_counter() - t0)
        lines.append("%s: %.3f ms" % (name, best_s * 1e3))
    sp.lib.sp_b200_FlushTextureCache()
    with open(path, "a") as f:
        f.write("\n".join(lines) + "\n")


def test_cube_map_reference_sizes(gpu_sp):
    """The sizes the reference bakes at start-up (main.cpp:1307-1315): a 4096x2048 map to 6 x 1024^2
    faces and to a 6 x 32^2 irradiance map on the uniform grid (1008 terms per texel) -- bit-exact in
    deterministic-math mode; with CUDA's float libm (mathMode 1) against the plain reference within
    a stated tolerance: cube map texels differ by more than 1e-3 relative on at most 2 % of the
    texels (a 1-ulp change of u or v moves the bilinear weights by 4096 ulp), irradiance texels by
    at most 1e-3 relative."""
    sp = gpu_sp
    env = W.make_env_map(4096, 2048)
    sp.set_params(mathMode=0)
    chk = best(dm=True)
    want_cube = chk.create_cube_map(env, 1024, 1024)
    got_cube = sp.create_cube_map(env, 1024, 1024)
    assert same_bits(got_cube, want_cube)
    want_irr = chk.create_irradiance_cube_map(env, 32, 32)
    assert same_bits(sp.create_irradiance_cube_map(env, 32, 32), want_irr)
    _time_bakers(sp, env)
    sp.set_params(mathMode=1)
    try:
        plain = best(dm=False)
        fast_cube = sp.create_cube_map(env, 256, 256)
        ref_cube = plain.create_cube_map(env, 256, 256)
        rel = np.abs(fast_cube - ref_cube) / np.maximum(np.abs(ref_cube), 1e-3)
        assert (rel.max(axis=-1) > 1e-3).mean() <= 0.02
        fast_irr = sp.create_irradiance_cube_map(env, 32, 32)
        ref_irr = plain.create_irradiance_cube_map(env, 32, 32)
        assert np.allclose(fast_irr, ref_irr, rtol=1e-3, atol=0)
    finally:
        sp.set_params(mathMode=0)


def test_work_queue_driver_matches_reference_scheduling(gpu_sp):
    """the reference's own frame loop -- AddRayTracingWorkQueue pushes one
    sp_Task per 64x64 tile, the workers pop, reseed 0xF51C0E49 and call sp_PathTraceTile
    (main.cpp:728-759, 819-844) -- through sp_b200_AddRayTracingWorkQueue /
    sp_b200_DrainRayTracingWorkQueue, against the checker's restatement of that pool over the
    reference's sp_PathTraceTile (ora_render_tiles): image bit-exact (deterministic math), summed
    counters equal, one sp_Metrics per tile, ragged right/bottom tiles, 2 spp; a queue too small for
    the frame takes the first `capacity` tiles only (the reference would trip its Assert at
    work_queue.h:27 instead)."""
    sp = gpu_sp
    wl = W.config1(200, 150, env_size=(256, 128))      # 4 x 3 tiles, the last column / row ragged
    r = sp.Renderer().load_workload(wl)
    chk = best(True).scene().load_workload(wl)
    for spp in (1, 2):
        sp.set_params(samplesPerPixel=spp, bounceCount=3, mathMode=0, tileWidth=64, tileHeight=64)
        r.image[:] = 0
        tiles, per_tile = r.render_work_queue()
        cimg, cm, _ = chk.render_tiles(64, 64, spp=spp, bounces=3, threads=4)
        assert tiles == 12 and per_tile.shape == (12, 12)
        assert same_bits(r.image, cimg)
        assert np.array_equal(per_tile[:, 1:5].sum(axis=0), cm[1:5])
        # per tile: paths = pixels x spp, in ComputeTiles (row-major) order
        assert per_tile[0, 1] == 64 * 64 * spp and per_tile[3, 1] == 8 * 64 * spp and per_tile[11, 1] == 8 * 22 * spp
        assert np.all(per_tile[:, 0] > 0)
    # the tile equals what one sp_PathTraceTile call gives
    one = r.image.copy()
    r.image[:] = 0
    r.path_trace_tile((64, 64, 128, 128), 0xF51C0E49)
    assert same_bits(r.image[64:128, 64:128], one[64:128, 64:128]) and not r.image[:64].any()
    r.image[:] = 0
    tiles, per_tile = r.render_work_queue(capacity=5)
    assert tiles == 5 and not r.image[64:, 64:].any() and same_bits(r.image[:64], one[:64])
    sp.set_params(samplesPerPixel=1)
    chk.close()
    r.close()


def test_device_lbvh_builder(gpu_sp):
    """sp_BuildMeshMidphase with SP_B200_BUILDER_DEVICE_LBVH (k_lbvh_keys, radix
    sort, k_lbvh_nodes, k_lbvh_fit with the collapse's dynamic programme, k_lbvh_emit level by level: the
    finished 4-wide tree comes back and the host only checks it).  (1) The tree has the shape the
    host emulation of the same per-element functions gives (node count, depth, stack need) -- the
    device passes computed the same binary tree; (2) results do not depend on the builder: the
    fixtures made by the unmodified reference (image, hit ids, ray queries, serial tile stream) are
    reproduced bit for bit on device-built trees, wavefront and per-pixel kernels, with the
    candidate lists and the coverage pass on; (3) a 2-triangle mesh falls back to the host builder
    and says so."""
    sp = gpu_sp
    hs = ora.load_hostsim()
    sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_DEVICE_LBVH)
    try:
        for mesh in (W.load_mesh("bunny"), W.load_mesh("monkey"), W.icosphere_mesh(6)):
            r = sp.Renderer()
            r.add_mesh(mesh.vertices, mesh.indices)
            info = sp.last_build_info()
            want = hs.build_info(mesh.vertices, mesh.indices, 1)
            assert info.builder == sp.BUILDER_DEVICE_LBVH and info.fellBack == 0 and info.deviceMs > 0
            assert info.triangleCount == len(mesh.indices) // 3
            assert (info.nodeCount, info.maxDepth, info.stackNeed) == (want["nodeCount"], want["maxDepth"], want["stackNeed"])
            r.close()
        g = np.load(os.path.join(GOLD, "g1_bunny_96x64.npz"))
        for mode in (0, 1):
            sp.set_params(samplesPerPixel=2, bounceCount=3, mathMode=0, renderMode=mode, cullByDistance=1)
            r = sp.Renderer().load_workload(W.config1(96, 64, env_size=(512, 256)))
            assert sp.last_build_info().fellBack == 0
            img, m = r.render_frame(frame=1)
            assert same_bits(img, g["image_ref_dm"]) and np.array_equal(m[1:5], g["metrics_ref_dm"])
            ph = r.primary_hits(sample=0, frame=1)
            assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
            r.image[:] = 0
            state, tm = r.path_trace_tile((16, 8, 48, 40), 0xF51C0E49)
            assert state == int(g["tile_state_ref_dm"]) and same_bits(r.image, g["tile_image_ref_dm"])
            r.close()
        sp.set_params(samplesPerPixel=1, renderMode=0)
        g = np.load(os.path.join(GOLD, "g2_monkey_160x90_primary.npz"))
        r = sp.Renderer().load_workload(W.config2(160, 90, env_size=(64, 32)))
        ph = r.primary_hits()
        assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
        r.close()
        g = np.load(os.path.join(GOLD, "g3_multi_80x60.npz"))
        r = sp.Renderer().load_workload(W.multi_object_workload(width=80, height=60, spp=2, env_size=(256, 128)))
        q = r.intersect_rays(g["origins"], g["dirs"])
        assert same_bits(q["t"], g["rays_t"]) and np.array_equal(q["tri"], g["rays_tri"])
        assert np.array_equal(q["obj"], g["rays_obj"]) and same_bits(q["normal"], g["rays_normal"])
        r.close()
        tiny = W.plane_mesh()
        r = sp.Renderer()
        r.add_mesh(tiny.vertices, tiny.indices)
        assert sp.last_build_info().fellBack == 1 and sp.last_build_info().deviceMs == 0
        r.close()
    finally:
        sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_HOST_SAH)
        sp.set_params(samplesPerPixel=1, renderMode=0)
    if os.environ.get("SPB_TIMING_OUT"):
        sphere = W.icosphere_mesh(6)
        lines = []
        for builder in (sp.BUILDER_HOST_SAH, sp.BUILDER_DEVICE_LBVH):
            sp.lib.sp_b200_SetMeshBuilder(builder)
            for name, mesh in (("bunny 4968", W.load_mesh("bunny")), ("monkey 15744", W.load_mesh("monkey")), ("icosphere 81920", sphere)):
                best_wall, best_dev = 1e9, 1e9
                for _ in range(3):
                    r = sp.Renderer()
                    r.add_mesh(mesh.vertices, mesh.indices)
                    info = sp.last_build_info()
                    best_wall, best_dev = min(best_wall, info.wallMs), min(best_dev, info.deviceMs)
                    r.close()
                lines.append("builder %d %s triangles: sp_BuildMeshMidphase %.2f ms wall, device passes %.3f ms, %d nodes"
                             % (builder, name, best_wall, best_dev, info.nodeCount))
        sp.lib.sp_b200_SetMeshBuilder(sp.BUILDER_HOST_SAH)
        with open(os.environ["SPB_TIMING_OUT"], "a") as f:
            f.write("\n".join(lines) + "\n")
