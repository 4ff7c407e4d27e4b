"""Environment pre-processing (src/cubemap.cpp): CreateCubeMap and
CreateIrradianceCubeMap.

CPU part (this file, `-m "not gpu"`): the port's restatement (oracle/sp_oracle.cpp) and the host
build of the device arithmetic (tests/hostsim over vk_cinematic_b200/csrc/spb_cubemap.cuh, with
the same per-texel RNG jump and term-then-fold decomposition the kernels use) against
  * the committed fixture tests/golden/g4_cubemap.npz, made from the UNMODIFIED reference
    (tools/make_cubemap_golden.py), and
  * the reference itself (oracle/_ref) where it is built,
bit for bit; plus known answers that do not depend on any implementation.
The GPU part is in tests/test_gpu_parity.py (test_cube_map_*, test_irradiance_*).
"""
import os

import numpy as np
import pytest

import ora
from vk_cinematic_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g4_cubemap.npz")


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                          np.ascontiguousarray(b, np.float32).view(np.uint32))


def golden():
    g = np.load(GOLD)
    env = W.make_env_map(int(g["env_params"][0]), int(g["env_params"][1]), str(g["env_variant"]))
    assert int(env.view(np.uint32).sum(dtype=np.uint64)) == int(g["env_checksum"]), "env generator changed"
    return g, env


@pytest.mark.parametrize("which", ["port", "port_dm", "hostsim"])
def test_golden_fixture(which, request):
    """Fixture made by the unmodified reference: the port equals it in both libm modes, the host
    build of the device code in deterministic-math mode."""
    lib = request.getfixturevalue(which)
    g, env = golden()
    tag = "" if which == "port" else "_dm"
    cw, ch = (int(v) for v in g["cube_size"])
    iw, ih = (int(v) for v in g["irradiance_size"])
    spp = int(g["spp"])
    assert same_bits(lib.create_cube_map(env, cw, ch), g["cube" + tag])
    assert same_bits(lib.create_irradiance_cube_map(env, iw, ih, spp=spp, sampling=0), g["irradiance_uniform" + tag])
    assert same_bits(lib.create_irradiance_cube_map(env, iw, ih, spp=spp, sampling=1), g["irradiance_random" + tag])


def test_port_and_hostsim_equal_reference(ref, ref_dm, port, port_dm, hostsim):
    """Side by side with the reference's own code on other sizes: ragged faces, 1x1 faces, a sample
    count that is not a power of two (the RNG jump strides change with it)."""
    env = W.make_env_map(160, 80)
    for (w, h, spp) in ((7, 9, 13), (1, 1, 1), (16, 16, 32)):
        for r, p in ((ref, port), (ref_dm, port_dm), (ref_dm, hostsim)):
            assert same_bits(r.create_cube_map(env, 2 * w, 2 * h), p.create_cube_map(env, 2 * w, 2 * h))
            for sampling in (0, 1):
                a = r.create_irradiance_cube_map(env, w, h, spp=spp, sampling=sampling)
                b = p.create_irradiance_cube_map(env, w, h, spp=spp, sampling=sampling)
                assert same_bits(a, b), (r.name, p.name, w, h, spp, sampling)


def test_hostsim_equals_port_for_other_sample_deltas(port_dm, hostsim):
    """sampleDelta is a literal in the reference (cubemap.cpp:160); the port and the device code
    take it as an argument and must agree on the float-accumulated loop values for any step."""
    env = W.make_env_map(64, 32)
    for delta in (0.25, 0.05, 1.0, 3.0):
        a = port_dm.create_irradiance_cube_map(env, 3, 3, sampling=0, sample_delta=delta)
        b = hostsim.create_irradiance_cube_map(env, 3, 3, sampling=0, sample_delta=delta)
        assert same_bits(a, b), delta


def test_known_answers(port):
    """Implementation-free checks.  A constant map bakes to the same constant (alpha included) in
    the cube map; its irradiance on the uniform grid is PI * L * mean(cos*sin) over the 63 x 16 grid
    (cubemap.cpp:162-198), i.e. L * 0.9981 for sampleDelta 0.1, with alpha 1; radiance above
    RADIANCE_CLAMP (10) is clamped before integration.  Face centres look along the face axes:
    with a map that is red for x > 0 (u in the first and last quarter), face +X is red."""
    const = np.tile(np.array([0.5, 2.0, 4.0, 0.25], np.float32), (16, 32, 1))
    cube = port.create_cube_map(const, 4, 4)
    assert np.array_equal(cube.reshape(-1, 4), np.tile(const[0, 0], (96, 1)))
    irr = port.create_irradiance_cube_map(const, 2, 2, sampling=0)
    phis = np.arange(63) * 0.1
    thetas = np.arange(16) * 0.1
    weight = np.pi * np.mean(np.cos(thetas) * np.sin(thetas)) * len(phis) / len(phis)
    assert np.allclose(irr[..., :3], const[0, 0, :3] * weight, rtol=2e-5)
    assert np.all(irr[..., 3] == 1.0)
    bright = const.copy()
    bright[..., :3] = 50.0
    assert np.allclose(port.create_irradiance_cube_map(bright, 1, 1, sampling=0)[..., :3], 10.0 * weight, rtol=2e-5)
    # random branch, constant map: sum of L * cosine / spp with cosine in [0, 1], clamped at 10
    rnd = port.create_irradiance_cube_map(const, 2, 2, spp=64, sampling=1)
    assert np.all(rnd[..., :3] <= const[0, 0, :3]) and np.all(rnd[..., :3] > 0.3 * const[0, 0, :3])
    # orientation: azimuth = atan2(z, x), u = az / 2pi; x > 0 <=> u < 1/4 or u > 3/4
    half = np.zeros((32, 64, 4), np.float32)
    half[:, :16, 0] = 1.0
    half[:, 48:, 0] = 1.0
    faces = port.create_cube_map(half, 8, 8)
    assert faces[0, 2:6, 2:6, 0].min() == 1.0 and faces[1, 2:6, 2:6, 0].max() == 0.0
    # +Y is the top of the map: v = cos(inc)/2 + 1/2 flipped -> row 0 is straight up
    top = np.zeros((32, 64, 4), np.float32)
    top[:8, :, 1] = 1.0
    faces = port.create_cube_map(top, 8, 8)
    assert faces[2, 3:5, 3:5, 1].min() == 1.0 and faces[3, :, :, 1].max() == 0.0


def test_serial_stream_is_shared_across_texels(port):
    """The random branch never reseeds (cubemap.cpp:122-123): texel i starts where texel i-1
    stopped, so baking a 2x1 face differs from baking its texels separately -- the property the
    device code's per-texel jump has to reproduce (checked bit for bit by the tests above)."""
    env = W.make_env_map(64, 32)
    both = port.create_irradiance_cube_map(env, 2, 1, spp=8, sampling=1)
    again = port.create_irradiance_cube_map(env, 2, 1, spp=8, sampling=1)
    assert same_bits(both, again)
    assert not same_bits(both[0, 0, 0], both[0, 0, 1])


def _direction_image(width, height):
    """An equirectangular map whose texel holds the direction it represents (the construction of
    the reference's TestCreateCubeMap, unit_tests.cpp:283-346, at texel centres): az = 2 pi u,
    cos(inc) = 1 - 2 v (MapToEquirectangular, math_lib.h:873-884, with the v flip of the bakers)."""
    u = (np.arange(width) + 0.5) / width
    v = (np.arange(height) + 0.5) / height
    az = 2 * np.pi * u[None, :]
    cos_inc = 1 - 2 * v[:, None]
    sin_inc = np.sqrt(np.maximum(0, 1 - cos_inc ** 2))
    img = np.ones((height, width, 4), np.float32)
    img[..., 0] = sin_inc * np.cos(az)
    img[..., 1] = cos_inc * np.ones_like(az)
    img[..., 2] = sin_inc * np.sin(az)
    return img


@pytest.mark.parametrize("which", ["port", "hostsim", "ref"])
def test_reference_cube_map_unit_tests(which, request):
    """unit_tests.cpp:480-538 (TestMapCubeMapFaceToVector, TestMapCubeMapFaceToBasisVectors) and
    :283-346 (TestCreateCubeMap, the test the reference keeps disabled: "our terrible sampling is
    struggling at such a low res" -- restated at a resolution where it holds), through the baker's
    output: on a map of directions, texel (1, 1) of a 2x2 face looks exactly along the face's forward
    axis (fx = fy = 0), texel (1, 0) along forward + up, texel (0, 1) along forward - right."""
    lib = request.getfixturevalue(which)
    faces = lib.create_cube_map(_direction_image(512, 256), 2, 2)[..., :3].astype(np.float64)
    forward = np.array([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)], np.float64)
    up = np.array([(0, 1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1), (0, 1, 0), (0, 1, 0)], np.float64)
    right = np.array([(0, 0, -1), (0, 0, 1), (1, 0, 0), (1, 0, 0), (1, 0, 0), (-1, 0, 0)], np.float64)
    for layer in range(6):
        # bilinear blend of unit vectors 0.7 degrees apart; straight up / down the lookup lands on the
        # centre of the first / last row (v clamps to the edge), sin(inc) = 0.088 away from the pole
        tol = 0.1 if layer in (2, 3) else 2e-2
        assert np.allclose(faces[layer, 1, 1], forward[layer], atol=tol), layer
        assert np.allclose(faces[layer, 0, 1] * np.sqrt(2), forward[layer] + up[layer], atol=2 * tol), layer
        assert np.allclose(faces[layer, 1, 0] * np.sqrt(2), forward[layer] - right[layer], atol=2 * tol), layer


def test_sphere_coords_sample(port):
    """unit_tests.cpp:540-558 (TestSphereCoordsSample): (phi, theta) = (0, 0) is the tangent-space
    pole (0, 1, 0), which the uniform branch maps onto the texel's own direction -- so the first
    term of every texel's sum is a lookup straight along the normal."""
    assert np.allclose(port.spherical_to_cartesian((0.0, 0.0)), (0, 1, 0), atol=1.2e-7)
    env = _direction_image(256, 128)
    # with a step that leaves one sample (phi = theta = 0 only): irradiance = PI * L(normal) * cos0 * sin0 = 0
    one = port.create_irradiance_cube_map(env, 2, 2, sampling=0, sample_delta=7.0)
    assert np.all(one[..., :3] == 0) and np.all(one[..., 3] == 1)
