"""The oracle port (and the hostsim build of the device arithmetic) against the committed golden
fixtures, which tools/make_golden.py produced from the UNMODIFIED reference (CPU only)."""
import os

import ctypes as C

import numpy as np
import pytest

from vk_cinematic_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, dtype=np.float32).view(np.uint32),
                          np.ascontiguousarray(b, dtype=np.float32).view(np.uint32))


def test_g1_port_image_and_hits(port, port_dm):
    g = gold("g1_bunny_96x64.npz")
    wl = W.config1(96, 64, env_size=(512, 256))
    for L, key in ((port, "ref"), (port_dm, "ref_dm")):
        s = L.scene().load_workload(wl)
        img, m = s.render_seeded(spp=2, bounces=3, frame=1)
        assert same_bits(img, g["image_" + key]) and np.array_equal(m[1:5], g["metrics_" + key])
        tile = np.zeros_like(img)
        state, tm = s.path_trace_tile(tile, (16, 8, 48, 40), 2, 3, 0xF51C0E49)
        assert state == int(g["tile_state_" + key]) and same_bits(tile, g["tile_image_" + key])
        assert np.array_equal(tm[1:5], g["tile_metrics_" + key])
        if key == "ref":
            ph = s.primary_hits(sample=0, frame=1)
            assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
        s.close()


def test_g2_port_monkey_primary(port):
    g = gold("g2_monkey_160x90_primary.npz")
    s = port.scene().load_workload(W.config2(160, 90, env_size=(64, 32)))
    ph = s.primary_hits()
    assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
    s.close()


def test_g3_port_rays_and_image(port, port_dm):
    g = gold("g3_multi_80x60.npz")
    wl = W.multi_object_workload(width=80, height=60, spp=2, env_size=(256, 128))
    s = port.scene().load_workload(wl)
    r = s.intersect_rays(g["origins"], g["dirs"])
    assert same_bits(r["t"], g["rays_t"]) and np.array_equal(r["tri"], g["rays_tri"])
    assert np.array_equal(r["obj"], g["rays_obj"]) and np.array_equal(r["material"], g["rays_material"])
    assert same_bits(r["normal"], g["rays_normal"]) and same_bits(r["uv"], g["rays_uv"])
    img, m = s.render_seeded(spp=2, bounces=3, frame=5)
    assert same_bits(img, g["image_ref"])
    s.close()
    s = port_dm.scene().load_workload(wl)
    img, m = s.render_seeded(spp=2, bounces=3, frame=5)
    assert same_bits(img, g["image_ref_dm"]) and np.array_equal(m[1:5], g["metrics_ref_dm"])
    s.close()


# hostsim = vk_cinematic_b200/csrc/spb_core.cuh + the host builder compiled for the host: the
# LOGIC of the CUDA path (not the GPU) against the reference's outputs.  No GPU parity claim.

@pytest.mark.parametrize("stepped", [0, 1, 2])
@pytest.mark.parametrize("cull", [1, 0])
def test_hostsim_logic_against_golden(hostsim, cull, stepped):
    """stepped = 1 traverses with the resumable state machine of the production trace kernel
    (spb_core.cuh trav_*, single-object entry at ray start included), 2 with the second machine
    (conservative inner tests, -DSPB_TRAV2 builds), 0 with the non-resumable walk of the per-pixel
    kernels."""
    hostsim.lib.hostsim_set_cull(cull)
    hostsim.lib.hostsim_set_stepped(stepped)
    g = gold("g1_bunny_96x64.npz")
    s = hostsim.scene().load_workload(W.config1(96, 64, env_size=(512, 256)))
    img, m = s.render_seeded(spp=2, bounces=3, frame=1)
    assert same_bits(img, g["image_ref_dm"]) and np.array_equal(m[1:5], g["metrics_ref_dm"])
    ph = s.primary_hits(sample=0, frame=1)
    assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
    tile = np.zeros_like(img)
    state, tm = s.path_trace_tile(tile, (16, 8, 48, 40), 2, 3, 0xF51C0E49)
    assert state == int(g["tile_state_ref_dm"]) and same_bits(tile, g["tile_image_ref_dm"])
    s.close()

    g = gold("g3_multi_80x60.npz")
    s = hostsim.scene().load_workload(
        W.multi_object_workload(width=80, height=60, spp=2, env_size=(256, 128)))
    r = s.intersect_rays(g["origins"], g["dirs"])
    assert same_bits(r["t"], g["rays_t"]) and np.array_equal(r["tri"], g["rays_tri"])
    assert np.array_equal(r["obj"], g["rays_obj"]) and same_bits(r["normal"], g["rays_normal"])
    img, m = s.render_seeded(spp=2, bounces=3, frame=5)
    assert same_bits(img, g["image_ref_dm"])
    s.close()
    hostsim.lib.hostsim_set_cull(1)
    hostsim.lib.hostsim_set_stepped(0)


def test_hostsim_five_bounces_against_port(hostsim, port_dm):
    wl = W.config1(64, 48, env_size=(256, 128))
    a = port_dm.scene().load_workload(wl)
    b = hostsim.scene().load_workload(wl)
    ia, ma = a.render_seeded(spp=2, bounces=5, frame=2)
    ib, mb = b.render_seeded(spp=2, bounces=5, frame=2)
    assert same_bits(ia, ib) and np.array_equal(ma[1:5], mb[1:5])
    a.close()
    b.close()


def test_c5_instanced_scene_hostsim_vs_port(hostsim, port_dm):
    """BASELINE configs[4] shape: 182 objects (64 bunnies + 118 level-6 icospheres, 9.98 M
    instanced triangles), far past the reference's 32-object table, so the checker is the port
    (median-split trees above 24 000 leaves; same leaf predicate).  The device arithmetic built
    for the host must agree bit for bit -- image, object ids, triangle ids."""
    wl = W.config5(160, 90, spp=1, bounces=5, env_size=(256, 128))
    assert wl.notes["objects"] == 182 and wl.notes["instanced_triangles"] > 9_900_000
    hostsim.lib.hostsim_set_stepped(1)
    a = port_dm.scene().load_workload(wl)
    b = hostsim.scene().load_workload(wl)
    ia, ma = a.render_seeded(spp=1, bounces=5, frame=4)
    ib, mb = b.render_seeded(spp=1, bounces=5, frame=4)
    assert same_bits(ia, ib) and np.array_equal(ma[1:5], mb[1:5])
    pa, pb = a.primary_hits(), b.primary_hits()
    assert np.array_equal(pa["obj"], pb["obj"]) and np.array_equal(pa["tri"], pb["tri"])
    assert len(np.unique(pa["obj"])) > 60
    hostsim.lib.hostsim_set_stepped(0)
    a.close()
    b.close()


@pytest.mark.parametrize("case", ["framed", "close_up", "off_centre", "big_env"])
def test_hostsim_shortcuts_equal_plain_evaluation(hostsim, case):
    """The two shortcuts of the wavefront renderer that rest on the size of the reference's jitter
    (spb_core.cuh): per-pixel candidate triangles for camera rays, and sky pixels settled by one
    environment lookup.  Every camera ray's shortcut result must equal its walk's (t bits, triangle,
    object; equal-t ties aside) and every one-lookup pixel must equal the sample loop's bits.
    Host build of the device arithmetic: a logic check, no GPU claim."""
    spp = 8
    size, env = (224, 160), (512, 256)
    if case == "big_env":
        # the bench's environment size; a large image, because the jitter in DIRECTION shrinks with
        # the square of the resolution and a 4096-texel map needs it small
        size, env, spp = (1600, 400), (4096, 2048), 4
    wl = W.config1(size[0], size[1], env_size=env)
    px, py, pz = wl.camera_position
    if case == "close_up":
        wl.camera_position = (px, py + 0.02, pz * 0.45)
    elif case == "off_centre":
        wl.camera_position = (px + 0.07, py - 0.04, pz * 0.8)
        wl.camera_rotation = W.quat_axis_angle((0.3, 1.0, 0.1), 0.35)
    s = hostsim.scene().load_workload(wl)
    out = np.zeros(9, np.uint64)
    hostsim.lib.hostsim_check_shortcuts(s.h, spp, 3, os.cpu_count() or 1, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    pixels, rays, fallback, mismatch, ties, sky, settled, sky_bad, sum_k = (int(v) for v in out)
    assert pixels == size[0] * size[1] and rays == pixels * spp
    assert mismatch == ties and ties <= 1e-4 * rays, (mismatch, ties)   # only equal-t ties may differ
    assert fallback <= 0.02 * pixels                                    # the list almost never overflows
    assert sky_bad == 0
    assert sky > 0.2 * pixels and settled >= 0.8 * sky                  # the shortcut actually applies
    assert sum_k > 0
    s.close()


@pytest.mark.parametrize("case", ["bunny", "bunny_edge", "multi_object", "instanced"])
def test_hostsim_coverage_is_conservative(hostsim, case):
    """Coverage pass (spb_core.cuh cover_triangle, what k_cover runs per instanced triangle): no camera
    ray of any pixel in an unmarked 8x4 block may hit anything.  The multi-object scene has a ground
    plane that passes under the camera (the "everything" flag); the instanced one is the C5 layout at
    reduced size.  Host build of the device arithmetic: a logic check, no GPU claim."""
    spp = 4
    if case == "bunny":
        wl = W.config1(320, 200, env_size=(256, 128))
    elif case == "bunny_edge":
        wl = W.config1(333, 187, env_size=(256, 128))
        px, py, pz = wl.camera_position
        wl.camera_rotation = W.quat_axis_angle((0, 1, 0), np.pi * 0.15)
        wl.camera_position = (px + 0.05, py, pz * 0.35)
    elif case == "multi_object":
        wl = W.multi_object_workload(width=200, height=150, spp=1, env_size=(256, 128))
    else:
        wl = W.config5(240, 136, spp=1, bounces=2, env_size=(256, 128), sphere_level=3)
    s = hostsim.scene().load_workload(wl)
    out = np.zeros(5, np.uint64)
    hostsim.lib.hostsim_check_coverage(s.h, spp, 2, os.cpu_count() or 1, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    blocks, unmarked, everything, rays, hits = (int(v) for v in out)
    assert hits == 0
    if case == "multi_object":
        assert everything == 1
    else:
        # (the instanced field and the close-up at the view's edge fill most of the view; the framed
        # single mesh leaves most of it empty)
        floor = 0.3 if case == "bunny" else 0.02
        assert everything == 0 and floor * blocks < unmarked < blocks and rays > 0
    s.close()


def test_hostsim_wide_slab_is_conservative(hostsim):
    """The resumable machine opens subtrees with a padded four-FFMA box test (spb_core.cuh
    slab_wide) and applies the reference's predicate (simd.h:198-271, here slab_fast) only to the
    leaves that come off the stack.  That is only sound if the padded test passes WHENEVER the
    reference's does.  4 M seeded adversarial (box, ray) pairs -- flat and point boxes, origins on
    faces and 1000 extents away, direction components down to 1e-28, rays aimed at corners and
    edges and nudged by one ulp -- must show no pair that the exact test passes and the wide one
    rejects.  (Every ray here is aimed AT its box, so the pairs the exact test rejects are grazing
    misses by a few ulps: the wide test is expected to accept a good part of those, not all.)"""
    total = np.zeros(4, np.uint64)
    for seed in (0x1A34C249, 0x45BA12F3, 77, 123456789):
        out = np.zeros(4, np.uint64)
        hostsim.lib.hostsim_check_wide_slab(seed, 1_000_000, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        total += out
    pairs, exact, wide, violations = (int(x) for x in total)
    assert pairs > 3_500_000 and exact > pairs // 10
    assert violations == 0
    assert exact <= wide < pairs - pairs // 4, (pairs, exact, wide)


def test_nested_objects_tlas_stays_inside_its_stack_share(hostsim, port_dm):
    """intersect_scene() gives the TLAS walk a third of the traversal stack and
    used to drop children silently beyond it, while nothing limited the TLAS depth once
    sp_b200_AddObjectToScene lifted the 32-object cap.  48 nested quads, each 2.25x the area of the
    one before, are an input a surface-area split peels one at a time: the SAH tree of these boxes
    needs 37 stack entries (12 four-wide levels) against a share of 30; rebuilt with median splits it
    needs 14.  The builder
    must keep the TLAS inside SPB_TLAS_STACK_LIMIT (median-split rebuild), and all three walks --
    the exact non-resumable one, the resumable machine, the port's -- must then see every object."""
    wl = W.nested_objects_workload(48)
    a = port_dm.scene().load_workload(wl)
    ia, ma = a.render_seeded(spp=1, bounces=3, frame=1)
    pa = a.primary_hits()
    assert len(np.unique(pa["obj"])) > 8            # rings of ever larger quads, seen through each other
    for stepped in (0, 1, 2):
        hostsim.lib.hostsim_set_stepped(stepped)
        b = hostsim.scene().load_workload(wl)
        info = np.zeros(4, np.uint32)
        hostsim.lib.hostsim_flat_info(b.h, info.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert info[3] == 48 and info[0] <= 30 + 2    # TLAS share (30) + the two-triangle mesh tree
        ib, mb = b.render_seeded(spp=1, bounces=3, frame=1)
        pb = b.primary_hits()
        assert same_bits(ia, ib) and np.array_equal(ma[1:5], mb[1:5])
        assert np.array_equal(pa["obj"], pb["obj"]) and same_bits(pa["t"], pb["t"])
        b.close()
    hostsim.lib.hostsim_set_stepped(0)
    a.close()


def _crack_scene(lib_scene, mesh):
    s = lib_scene
    s.add_mesh(mesh.vertices, mesh.indices, False)
    s.add_object(0, W.MATERIAL_SURFACE)
    s.build()
    return s


def test_hostsim_watertight_option_closes_the_cracks(hostsim, port):
    """sp_b200_Params::triangleTest = WATERTIGHT (north star: "watertight ray-triangle intersection"; an
    option, never the parity default -- SURVEY.md §0).  200 000 rays aimed at points lying exactly on
    shared edges and vertices of a sheet of triangles: the reference's Moller-Trumbore
    (ray_intersection.cpp:156-190), which tests the two neighbours independently, lets some of them
    through (and the device restatement of it must leak exactly the rays the port leaks: parity includes
    the cracks); the watertight test must let none through, and away from those rays it must agree with
    Moller-Trumbore on what is hit and, to a few ulps, where."""
    mesh, o, d = W.crack_test_inputs()
    a = _crack_scene(port.scene(), mesh)
    ra = a.intersect_rays(o, d)
    a.close()
    hostsim.lib.hostsim_set_stepped(1)
    b = _crack_scene(hostsim.scene(), mesh)
    rb = b.intersect_rays(o, d)
    b.close()
    hostsim.lib.hostsim_set_stepped(0)                   # the watertight option lives in the non-resumable walk
    hostsim.lib.hostsim_set_triangle_test(1)
    c = _crack_scene(hostsim.scene(), mesh)
    rc = c.intersect_rays(o, d)
    c.close()
    hostsim.lib.hostsim_set_triangle_test(0)
    assert same_bits(ra["t"], rb["t"])                    # parity, cracks included
    differ = ra["tri"] != rb["tri"]                      # rays ON an edge meet both neighbours at bit-equal t:
    assert np.all(ra["tri"][differ] >= 0) and np.all(rb["tri"][differ] >= 0)   # the winner is the order's
    leaks_mt, leaks_wt = int((rb["t"] < 0).sum()), int((rc["t"] < 0).sum())
    assert leaks_wt == 0, leaks_wt
    assert leaks_mt > 0, "the sheet was meant to show the reference's cracks"
    both = (rb["t"] > 0) & (rc["t"] > 0)
    assert both.sum() > 150000
    assert np.all(np.abs(rb["t"][both] - rc["t"][both]) <= 1e-5 * np.abs(rb["t"][both]))
