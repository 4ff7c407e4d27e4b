"""Device BVH builder, CPU part: the per-element functions the kernels run
(vk_cinematic_b200/csrc/spb_lbvh.cuh: Morton keys, Karras' binary radix tree) executed in a loop on
the host, then the host finalisation the device path shares (bvh4_from_binary: validation,
parent-first renumbering, 4-wide collapse).  The trees differ from the SAH builder's; results must
not (every triangle alone in a child slot with its own AABB): images, hit ids and ray queries are
compared bit for bit with the fixtures made by the unmodified reference.
The GPU part (the four device passes against this host emulation) is
tests/test_gpu_parity.py::test_device_lbvh_builder.
"""
import os

import numpy as np
import pytest

from vk_cinematic_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                          np.ascontiguousarray(b, np.float32).view(np.uint32))


@pytest.fixture(params=[1, 2], ids=["host_collapse", "device_collapse_emulation"])
def lbvh(hostsim, request):
    hostsim.set_builder(request.param)
    yield hostsim
    hostsim.set_builder(0)
    hostsim.lib.hostsim_set_stepped(0)


@pytest.mark.parametrize("stepped", [0, 1, 2])
def test_lbvh_trees_give_the_reference_results(lbvh, stepped):
    lbvh.lib.hostsim_set_stepped(stepped)
    g = np.load(os.path.join(GOLD, "g1_bunny_96x64.npz"))
    s = lbvh.scene().load_workload(W.config1(96, 64, env_size=(512, 256)))
    img, m = s.render_seeded(spp=2, bounces=3, frame=1)
    assert same_bits(img, g["image_ref_dm"]) and np.array_equal(m[1:5], g["metrics_ref_dm"])
    ph = s.primary_hits(sample=0, frame=1)
    assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
    s.close()
    g = np.load(os.path.join(GOLD, "g2_monkey_160x90_primary.npz"))
    s = lbvh.scene().load_workload(W.config2(160, 90, env_size=(64, 32)))
    ph = s.primary_hits()
    assert np.array_equal(ph["tri"], g["tri"]) and same_bits(ph["t"], g["t"])
    s.close()
    g = np.load(os.path.join(GOLD, "g3_multi_80x60.npz"))
    s = lbvh.scene().load_workload(W.multi_object_workload(width=80, height=60, spp=2, env_size=(256, 128)))
    r = s.intersect_rays(g["origins"], g["dirs"])
    assert same_bits(r["t"], g["rays_t"]) and np.array_equal(r["tri"], g["rays_tri"])
    assert np.array_equal(r["obj"], g["rays_obj"]) and same_bits(r["normal"], g["rays_normal"])
    s.close()


def test_lbvh_tree_shape(hostsim):
    """Every triangle is a leaf exactly once; depth and stack need stay within the traversal
    stack's share; the LBVH tree costs more area than the SAH tree (it is the faster build, not the
    better tree) but stays within 1.5x on the reference's meshes."""
    for name in ("bunny", "monkey"):
        mesh = W.load_mesh(name)
        sah = hostsim.build_info(mesh.vertices, mesh.indices, 0)
        lb = hostsim.build_info(mesh.vertices, mesh.indices, 1)
        assert lb["leafCount"] == sah["leafCount"] == len(mesh.indices) // 3
        assert lb["fellBack"] == 0 and lb["stackNeed"] + 4 <= 64
        assert sah["area"] <= lb["area"] <= 1.5 * sah["area"]
    sphere = W.icosphere_mesh(5)
    lb = hostsim.build_info(sphere.vertices, sphere.indices, 1)
    assert lb["leafCount"] == 20480 and lb["fellBack"] == 0


def test_device_collapse_emulation_matches_host_collapse(hostsim):
    """The device form of the 4-wide collapse (spb_lbvh.cuh lbvh_dp / lbvh_gather / lbvh_emit run level by level
    in a loop, then bvh4_adopt_device_tree) makes the tree the host collapse makes from the same binary tree:
    same node count, depth, stack need, leaves and summed area -- the dynamic programme is the same arithmetic
    in the same order; only the numbering inside a level may differ."""
    rng = np.random.RandomState(11)
    soups = [rng.rand(n, 3, 3).astype(np.float32) * np.float32(s) for n, s in ((9, 1.0), (64, 3.0), (1000, 0.25), (4097, 10.0))]
    meshes = [(W.load_mesh(n).vertices, W.load_mesh(n).indices) for n in ("bunny", "monkey")]
    meshes += [(W.icosphere_mesh(4).vertices, W.icosphere_mesh(4).indices)]
    meshes += [_mesh_from_triangles(t) for t in soups]
    one = [[0, 0, 0], [1, 0, 0], [0, 1, 0]]
    meshes.append(_mesh_from_triangles([one] * 300))
    for v, i in meshes:
        a = hostsim.build_info(v, i, 1)
        b = hostsim.build_info(v, i, 2)
        assert a["fellBack"] == 0 and b["fellBack"] == 0
        assert a == b, (a, b)


def test_damaged_device_trees_are_refused(hostsim):
    """bvh4_adopt_device_tree is the gate between what the GPU returns and what the traversal kernels walk (their
    pushes have no bound check): a tree with a leaf or node referenced twice, a cycle, a hidden child, a wrong
    depth, a repeated primitive, a child sticking out of its parent's box, a reference past the array or an orphan
    node is refused (the caller then builds on the host); the undamaged tree is accepted."""
    import ctypes as C
    f = C.POINTER(C.c_float)
    hostsim.lib.hostsim_adopt_damaged.argtypes = [f, f, C.c_uint32, C.c_int]
    hostsim.lib.hostsim_adopt_damaged.restype = C.c_int
    rng = np.random.RandomState(5)
    for n in (40, 700, 5000):
        c = rng.rand(n, 3).astype(np.float32) * 10
        mn = np.ascontiguousarray(c - rng.rand(n, 3).astype(np.float32) * 0.2)
        mx = np.ascontiguousarray(c + rng.rand(n, 3).astype(np.float32) * 0.2)
        got = [hostsim.lib.hostsim_adopt_damaged(mn.ctypes.data_as(f), mx.ctypes.data_as(f), n, k) for k in range(10)]
        assert got == [1] + [0] * 9, (n, got)


def _mesh_from_triangles(tris):
    tris = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    v = np.zeros((len(tris) * 3, 8), np.float32)
    v[:, :3] = tris.reshape(-1, 3)
    return v, np.arange(len(tris) * 3, dtype=np.uint32)


def test_lbvh_degenerate_inputs(hostsim):
    """Identical triangles (equal Morton keys: the sorted position breaks the tie), fewer triangles
    than a 4-wide node holds, a flat mesh (no extent on one axis), non-finite vertices, and
    geometrically spaced triangles whose radix tree is a spine as long as an axis has key bits (the
    rest share cell 0 and are split by position): whatever comes out must fit the traversal stack's
    share -- bvh4_from_binary refuses deeper trees and the SAH builder takes over."""
    one = [[0, 0, 0], [1, 0, 0], [0, 1, 0]]
    v, i = _mesh_from_triangles([one] * 300)
    info = hostsim.build_info(v, i, 1)
    assert info["leafCount"] == 300 and info["fellBack"] == 0 and info["stackNeed"] + 4 <= 64
    v, i = _mesh_from_triangles([one, one])
    assert hostsim.build_info(v, i, 1)["leafCount"] == 2
    rng = np.random.RandomState(4)
    flat = rng.rand(500, 3, 3).astype(np.float32)
    flat[..., 1] = 2.5
    v, i = _mesh_from_triangles(flat)
    assert hostsim.build_info(v, i, 1)["leafCount"] == 500
    bad = rng.rand(200, 3, 3).astype(np.float32)
    bad[7, 1, 0] = np.nan
    bad[9, 2, 2] = np.inf
    v, i = _mesh_from_triangles(bad)
    assert hostsim.build_info(v, i, 1)["leafCount"] == 200
    # spine: triangle k sits at 2^-k along x -- each Morton bit splits one triangle off
    k = np.arange(40, dtype=np.float64)
    base = (2.0 ** -k)[:, None, None] * np.ones((40, 3, 3))
    base[:, 1, 1] += 1e-9
    base[:, 2, 2] += 1e-9
    v, i = _mesh_from_triangles(np.concatenate([base, [[[0, 0, 0], [0, 1e-9, 0], [0, 0, 1e-9]]]]))
    info = hostsim.build_info(v, i, 1)
    assert info["leafCount"] == 41 and info["stackNeed"] + 4 <= 64


def test_lbvh_render_of_degenerate_scene_matches_port(lbvh, port_dm):
    """A scene made of the awkward meshes above renders the same image as the port."""
    wl = W.multi_object_workload(width=64, height=48, spp=1, env_size=(128, 64), count=6, seed=77)
    a = port_dm.scene().load_workload(wl)
    b = lbvh.scene().load_workload(wl)
    ia, ma = a.render_seeded(spp=1, bounces=3, frame=2)
    ib, mb = b.render_seeded(spp=1, bounces=3, frame=2)
    assert same_bits(ia, ib) and np.array_equal(ma[1:5], mb[1:5])
    a.close()
    b.close()
