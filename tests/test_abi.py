"""C-ABI boundary checks that need no GPU: libspb200.so loads, exports every symbol
include/sp_b200.h declares, its PODs have the reference's sizes, the host-only entry points
(tiles, work queue, camera, tables) agree with the oracle, and a compute call without a CUDA
device fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(np.float32).eps


def test_exports_every_declared_symbol(sp):
    names = sp.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(sp.lib, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", sp.LIB_PATH], text=True)
    exported = set(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert set(names) <= exported


def test_header_compiles_as_c_and_sizes_match(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "sp_b200.h"\nint main(void){return sizeof(sp_Scene)==7104?0:1;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_compute_tiles_and_work_queue(sp, port):
    """tile.h:11-42, work_queue.h:13-44 (unit_tests.cpp:40-122)"""
    for (w, h, tw, th, cap) in [(10, 10, 2, 2, 64), (9, 9, 2, 2, 64), (10, 10, 2, 2, 10),
                                (3840, 2160, 64, 64, 4096), (1024, 768, 64, 64, 4096), (0, 0, 64, 64, 8)]:
        tiles = (sp.Tile * cap)()
        n = sp.lib.ComputeTiles(w, h, tw, th, tiles, cap)
        en, et = port.compute_tiles(w, h, tw, th, cap)
        assert n == en
        got = np.frombuffer(tiles, dtype=np.uint32).reshape(cap, 4)[:n]
        assert np.array_equal(got, et[:n])
    buf = (C.c_uint8 * 256)()
    arena = sp.MemoryArena(C.addressof(buf), 0, 256)
    q = sp.lib.CreateWorkQueue(C.byref(arena), 4, 4)
    assert arena.size == 16 and q.maxObjects == 4
    for v in (1, 2):
        val = C.c_uint32(v)
        assert sp.lib.WorkQueuePush(C.byref(q), C.byref(val), 4)
    assert q.tail == 2
    first = C.cast(sp.lib.WorkQueuePop(C.byref(q), 4), C.POINTER(C.c_uint32))[0]
    second = C.cast(sp.lib.WorkQueuePop(C.byref(q), 4), C.POINTER(C.c_uint32))[0]
    assert (first, second) == (1, 2) and q.head == 2


def test_camera_and_film_positions(sp, port):
    """simd_path_tracer.cpp:1-63 (test_simd_path_tracer.cpp:137-200), bit-equal to the oracle"""
    plane = sp.ImagePlane(None, 4, 2)
    cam = sp.sp_Camera()
    sp.lib.sp_ConfigureCamera(C.byref(cam), C.byref(plane), sp.V3((0, 2, 0)), sp.Q((0, 0, 0, 1)), 0.1)
    assert cam.basis.right.tuple() == (1, 0, 0) and cam.basis.up.tuple() == (0, 1, 0)
    assert cam.basis.forward.tuple() == (0, 0, -1)
    assert (cam.halfPixelWidth, cam.halfPixelHeight) == (0.125, 0.25)
    assert (cam.halfFilmWidth, cam.halfFilmHeight) == (0.5, 0.25)
    rng = np.random.RandomState(7)
    for _ in range(20):
        pos = rng.uniform(-3, 3, 3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        fd = float(rng.uniform(0.1, 2.0))
        w, h = int(rng.randint(1, 4000)), int(rng.randint(1, 3000))
        plane = sp.ImagePlane(None, w, h)
        sp.lib.sp_ConfigureCamera(C.byref(cam), C.byref(plane), sp.V3(pos), sp.Q(q), fd)
        e = port.camera_fields(pos, q, fd, w, h)
        got = np.float32(cam.basis.right.tuple() + cam.basis.up.tuple() + cam.basis.forward.tuple() +
                         cam.position.tuple() + cam.filmCenter.tuple() +
                         (cam.halfPixelWidth, cam.halfPixelHeight, cam.halfFilmWidth, cam.halfFilmHeight))
        exp = np.concatenate([e["right"], e["up"], e["forward"], e["position"], e["filmCenter"],
                              [e["halfPixelWidth"], e["halfPixelHeight"], e["halfFilmWidth"],
                               e["halfFilmHeight"]]]).astype(np.float32)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
        s = port.scene()
        s.configure_camera(pos, q, fd, w, h)
        pix = rng.uniform(0, [w, h], (5, 2)).astype(np.float32)
        film = (sp.vec3 * 5)()
        n = sp.lib.sp_CalculateFilmPositions(C.byref(cam), film, pix.ctypes.data_as(C.POINTER(sp.vec2)), 5)
        assert n == 5
        got = np.frombuffer(film, dtype=np.float32).reshape(5, 3)
        assert np.array_equal(got.view(np.uint32), s.film_positions(pix).view(np.uint32))
        s.close()


def test_transform_aabb(sp, port):
    """aabb.h:29-58 (test_simd_path_tracer.cpp:202-214)"""
    r = sp.lib.TransformAabb(sp.V3((-0.5,) * 3), sp.V3((0.5,) * 3), sp.V3((5, 0, 0)), sp.Q((0, 0, 0, 1)),
                             sp.V3((1, 1, 1)))
    assert np.abs(np.float32(r.min.tuple()) - (4.5, -0.5, -0.5)).max() <= EPS
    assert np.abs(np.float32(r.max.tuple()) - (5.5, 0.5, 0.5)).max() <= EPS
    rng = np.random.RandomState(3)
    for _ in range(50):
        lo = rng.uniform(-2, 0, 3)
        hi = lo + rng.uniform(0.1, 3, 3)
        pos = rng.uniform(-5, 5, 3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        sc = rng.uniform(0.2, 3, 3)
        r = sp.lib.TransformAabb(sp.V3(lo), sp.V3(hi), sp.V3(pos), sp.Q(q), sp.V3(sc))
        emn, emx = port.transform_aabb(lo, hi, pos, q, sc)
        assert np.array_equal(np.float32(r.min.tuple()).view(np.uint32), emn.view(np.uint32))
        assert np.array_equal(np.float32(r.max.tuple()).view(np.uint32), emx.view(np.uint32))


def test_material_tables(sp):
    """sp_material_system.cpp:1-57: first match wins, NULL when absent, false when full"""
    ms = sp.sp_MaterialSystem()
    for i in range(sp.SP_MAX_MATERIALS):
        m = sp.sp_Material(sp.V3((i, 0, 0)), sp.U32_MAX, sp.V3((0, 0, 0)), sp.U32_MAX, 0.5)
        assert sp.lib.sp_RegisterMaterial(C.byref(ms), m, 100 + (i % 20))
    assert not sp.lib.sp_RegisterMaterial(C.byref(ms), m, 999)
    assert sp.lib.sp_FindMaterialById(C.byref(ms), 105).contents.albedo.x == 5.0  # not 25
    assert not sp.lib.sp_FindMaterialById(C.byref(ms), 7)
    px = np.zeros(4, np.float32)
    for i in range(sp.SP_MAX_IMAGES):
        img = sp.HdrImage(px.ctypes.data_as(C.POINTER(C.c_float)), 1 + i, 1)
        assert sp.lib.sp_RegisterTexture(C.byref(ms), img, i % 4)
    assert not sp.lib.sp_RegisterTexture(C.byref(ms), img, 99)
    assert sp.lib.sp_FindTexture(C.byref(ms), 2).contents.width == 3
    assert not sp.lib.sp_FindTexture(C.byref(ms), 50)


def test_seed_matches_oracle(sp, port):
    rng = np.random.RandomState(11)
    for _ in range(200):
        p, s, f = (int(v) for v in rng.randint(0, 2 ** 31, 3))
        assert sp.lib.sp_b200_Seed(p, s, f) == port.seed(p, s, f) != 0


def test_midphase_tree_structure_host(sp):
    """test_bvh.cpp:99-179 on this library's builder (host side, no device needed)"""
    from vk_cinematic_b200 import workloads as W
    for mesh in (W.icosphere_mesh(2), W.load_mesh("bunny"), W.triangle_mesh()):
        v = np.ascontiguousarray(mesh.vertices)
        i = np.ascontiguousarray(mesh.indices)
        m = sp.lib.sp_CreateMesh(v.ctypes.data_as(C.POINTER(sp.VertexPNT)), len(v),
                                 i.ctypes.data_as(C.POINTER(sp.u32)), len(i), 0)
        sp.lib.sp_BuildMeshMidphase(C.byref(m), None, None)
        info = sp.sp_b200_TreeInfo()
        sp.lib.sp_b200_MeshTreeInfo(m, C.byref(info))
        assert info.leafCount == len(i) // 3 and info.allLeavesReachable
        assert info.parentsContainChildren and info.maxDepth < 24
        lo, hi = mesh.bounds()
        assert np.array_equal(np.float32(info.rootMin.tuple()), lo)
        assert np.array_equal(np.float32(info.rootMax.tuple()), hi)
        sp.lib.sp_b200_ReleaseMesh(C.byref(m))


def test_reference_unit_tests_against_the_reference_itself():
    """The bar for tests/dropin: the reference's own test file, unmodified, against the reference's own
    sources (tests/dropin/_build/reference_unit_tests, built where the checkout is mounted).  15 of its 16
    tests pass; TestEvaluateLightPath fails THERE too (GGX at roughness 0 is 0/0), which is why
    the drop-in run leaves it out."""
    exe = os.path.join(ROOT, "tests", "dropin", "_build", "reference_unit_tests")
    if not os.path.exists(exe):
        import pytest
        pytest.skip("tests/dropin/_build not built (the reference checkout was never mounted here)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "16 Tests 1 Failures 0 Ignored" in p.stdout
    assert "TestEvaluateLightPath:FAIL" in p.stdout


def test_product_never_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/ (or hostsim)."""
    pkg = os.path.join(ROOT, "vk_cinematic_b200")
    bad = re.compile(r"oracle|libsporacle|libspref|hostsim|import\s+ora\b")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                for ln, line in enumerate(text.splitlines(), 1):
                    stripped = line.strip()
                    if stripped.startswith(("//", "#", "*", '"""')) or "DESIGN.md" in line:
                        continue
                    code = line.split("//")[0].split("#")[0]
                    assert not bad.search(code), (f, ln, line)
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libspb200.so")], text=True)
    assert "oracle" not in out and "spref" not in out


def test_compute_without_gpu_fails_loudly():
    """No CPU fallback: with no visible CUDA device a compute entry point aborts with a message."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from vk_cinematic_b200 import sp\n"
            "r = sp.Renderer()\n" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert p.returncode != 0
    assert "no CPU fallback" in p.stderr or "no usable CUDA device" in p.stderr
