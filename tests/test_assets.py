"""Asset input: sp_b200_LoadObj against a pure-Python restatement of its
loader rules (the rules of tools/make_mesh_fixtures.py, which made assets/*.npz) and, where the
reference checkout is mounted, against those fixtures on the reference's own OBJ files.  Host
code only: no GPU needed."""
import os

import numpy as np
import pytest

from vk_cinematic_b200 import sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_obj_python(path):
    """One vertex per distinct (v, vt, vn) corner in first-seen order; polygons fan-triangulated in
    file order; numbers through float() -> float32 (reference src/mesh.cpp:5-62 delegates all of
    this to assimp, which is absent: see vk_cinematic_b200/csrc/spb_assets.cpp)."""
    pos, tex, nrm, verts, index_of, indices = [], [], [], [], {}, []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            pos.append([float(x) for x in t[1:4]])
        elif t[0] == "vt":
            tex.append([float(t[1]), float(t[2]) if len(t) > 2 else 0.0])
        elif t[0] == "vn":
            nrm.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            corners = []
            for tok in t[1:]:
                parts = (tok.split("/") + ["", ""])[:3]
                key = []
                for k, n in zip(parts, (len(pos), len(tex), len(nrm))):
                    if k == "":
                        key.append(-1)
                    else:
                        i = int(k)
                        key.append(i - 1 if i > 0 else n + i)
                key = tuple(key)
                if key not in index_of:
                    index_of[key] = len(verts)
                    verts.append(key)
                corners.append(index_of[key])
            for k in range(1, len(corners) - 1):
                indices.extend([corners[0], corners[k], corners[k + 1]])
    out = np.zeros((len(verts), 8), np.float32)
    for i, (vi, ti, ni) in enumerate(verts):
        out[i, 0:3] = np.asarray(pos[vi], np.float32)
        if ni >= 0:
            out[i, 3:6] = np.asarray(nrm[ni], np.float32)
        if ti >= 0:
            out[i, 6:8] = np.asarray(tex[ti], np.float32)
    return out, np.asarray(indices, np.uint32)


def test_load_obj_fixture():
    path = os.path.join(HERE, "golden", "quad_mix.obj")
    v, i = sp.load_obj(path)
    ev, ei = load_obj_python(path)
    assert v.shape == ev.shape and np.array_equal(v.view(np.uint32), ev.view(np.uint32))
    assert np.array_equal(i, ei)
    # a quad, three triangles, a pentagon and a triangle: 2 + 3 + 3 + 1 = 9 triangles
    assert len(i) == 27 and i.max() == len(v) - 1
    # corner 1/1/1 is shared by the quad and the pentagon; "1", "1//2", "1/1" are distinct vertices
    assert len(v) == 4 + 3 + 3 + 3 + 2 + 3
    assert np.array_equal(v[0], np.array([0, 0, 0, 0, 0, 1, 0, 0], np.float32))
    # first corner of the pentagon: "-1" = the last v record so far, no vt, no vn
    assert np.array_equal(v[13], np.array([-1e-3, 25.0, -0.333333343, 0, 0, 0, 0, 0], np.float32))
    # last corner of the file: "4//-2" = v 4 with the first vn
    assert np.array_equal(v[17], np.array([0, 1, 0, 0, 0, 1, 0, 0], np.float32))


def test_load_obj_errors(tmp_path):
    assert sp.load_obj(os.path.join(HERE, "golden", "does_not_exist.obj")) is None
    for name, text in (("zero_index", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n"),
                       ("out_of_range", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n"),
                       ("two_corners", "v 0 0 0\nv 1 0 0\nf 1 2\n"),
                       ("no_faces", "v 0 0 0\nv 1 0 0\nv 0 1 0\n"),
                       ("bad_number", "v 0 zero 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")):
        p = tmp_path / (name + ".obj")
        p.write_text(text)
        assert sp.load_obj(str(p)) is None, name


@pytest.mark.parametrize("name", ["bunny", "monkey"])
def test_load_obj_reference_assets(name):
    """The reference's own meshes (BASELINE.json configs) load to exactly the committed fixtures
    every parity test renders."""
    src = os.path.join("/root/reference/assets", name + ".obj")
    if not os.path.exists(src):
        pytest.skip("reference checkout not mounted")
    v, i = sp.load_obj(src)
    d = np.load(os.path.join(ROOT, "assets", name + ".npz"))
    assert np.array_equal(v.view(np.uint32), d["vertices"].view(np.uint32)) and np.array_equal(i, d["indices"])


# ---------------------------------------------------------------------------------------------
# OpenEXR input: the library's LoadExrImage (same C ABI as the reference's asset_loader.h:11-22)
# against (a) the arrays the fixtures were written from (tools/make_exr_fixtures.py, OpenCV's
# encoder) and (b) the reference's own loader -- its asset_loader.cpp over the vendored tinyexr,
# compiled unmodified by `make -C oracle exrref` -- where that library is present.

EXR_DIR = os.path.join(HERE, "golden", "exr")
EXR_CASES = ["none_f32_rgb_13x9", "zip_f32_rgb_70x45", "zip_half_rgba_33x20", "zips_half_rgb_21x7",
             "rle_f32_rgb_40x11", "zip_f32_grey_17x5", "piz_half_rgb_16x8", "piz_half_rgb_70x45",
             "piz_f32_rgba_37x33"]


def _reference_exr_loader():
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libexrref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.LoadExrImage.argtypes = [C.POINTER(sp.HdrImage), C.c_char_p]
    lib.LoadExrImage.restype = C.c_int

    def load(p):
        img = sp.HdrImage()
        if lib.LoadExrImage(C.byref(img), os.fsencode(p)) != 0:
            return None
        out = np.ctypeslib.as_array(img.pixels, shape=(img.height, img.width, 4)).copy()
        sp._libc_free(img.pixels)
        return out
    return load


@pytest.mark.parametrize("name", EXR_CASES)
def test_load_exr_fixture(name):
    got = sp.load_exr(os.path.join(EXR_DIR, name + ".exr"))
    want = np.load(os.path.join(EXR_DIR, name + ".npy"))
    assert got is not None and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ref = _reference_exr_loader()
    if ref is not None:
        r = ref(os.path.join(EXR_DIR, name + ".exr"))
        assert r is not None and np.array_equal(got.view(np.uint32), r.view(np.uint32))


def test_load_exr_refuses_what_it_cannot_read(tmp_path):
    """Return value 1 (the reference's failure code), image untouched: unsupported compression
    (PXR24: this reader says no rather than guess), missing file, truncated files, garbage."""
    assert sp.load_exr(os.path.join(EXR_DIR, "pxr24_f32_rgb_16x8.exr")) is None
    assert sp.load_exr(os.path.join(EXR_DIR, "nope.exr")) is None
    data = open(os.path.join(EXR_DIR, "zip_f32_rgb_70x45.exr"), "rb").read()
    for cut in (3, 40, 400, len(data) // 2, len(data) - 7):
        p = tmp_path / f"cut{cut}.exr"
        p.write_bytes(data[:cut])
        assert sp.load_exr(str(p)) is None
    piz = open(os.path.join(EXR_DIR, "piz_half_rgb_70x45.exr"), "rb").read()
    for cut in (len(piz) // 3, len(piz) - 5):
        p = tmp_path / f"pizcut{cut}.exr"
        p.write_bytes(piz[:cut])
        assert sp.load_exr(str(p)) is None
    for k in range(0, len(piz), 97):          # corrupted PIZ streams must not crash
        broken = bytearray(piz)
        broken[k] ^= 0xA5
        p = tmp_path / "pizflip.exr"
        p.write_bytes(bytes(broken))
        sp.load_exr(str(p))
    p = tmp_path / "garbage.exr"
    p.write_bytes(bytes(range(256)) * 8)
    assert sp.load_exr(str(p)) is None
    # a flipped byte inside a compressed chunk must not crash (it may or may not be detected)
    broken = bytearray(data)
    broken[len(broken) // 2] ^= 0x5A
    p = tmp_path / "flipped.exr"
    p.write_bytes(bytes(broken))
    sp.load_exr(str(p))


# ---------------------------------------------------------------------------------------------
# Image output: sp_b200_SaveExrImage / sp_b200_SavePpm.  Three independent
# readers check every file: the library's own LoadExrImage, the reference's loader (tinyexr,
# compiled unmodified) where present, and OpenCV's OpenEXR decoder.

def _test_image(h, w, seed):
    rng = np.random.RandomState(seed)
    img = (rng.rand(h, w, 4) ** 3 * 60).astype(np.float32)
    img[h // 3: h // 2] = np.linspace(0, 3, w, dtype=np.float32)[None, :, None]  # smooth rows: matches to find
    img[0, 0] = (0.0, -0.0, 1e-8, 65519.0)       # zero, negative zero, a half subnormal's neighbourhood, just below overflow
    img[0, 1 % w] = (6.1e-5, 5.9e-8, 2.98e-8, 1.0)
    return img


@pytest.mark.parametrize("pixel_type", ["half", "float"])
@pytest.mark.parametrize("compression", ["none", "zips", "zip"])
@pytest.mark.parametrize("shape", [(45, 70), (1, 1), (16, 3), (33, 129)])
def test_save_exr_round_trip(tmp_path, pixel_type, compression, shape):
    img = _test_image(shape[0], shape[1], seed=shape[0] * 131 + shape[1])
    want = img.astype(np.float16).astype(np.float32) if pixel_type == "half" else img
    path = str(tmp_path / "out.exr")
    assert sp.save_exr(path, img, {"half": sp.EXR_HALF, "float": sp.EXR_FLOAT}[pixel_type],
                       {"none": sp.EXR_NONE, "zips": sp.EXR_ZIPS, "zip": sp.EXR_ZIP}[compression])
    got = sp.load_exr(path)
    assert got is not None and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ref = _reference_exr_loader()
    if ref is not None:
        r = ref(path)
        assert r is not None and np.array_equal(r.view(np.uint32), want.view(np.uint32))
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    bgra = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert bgra is not None and bgra.shape == want.shape
    assert np.array_equal(np.ascontiguousarray(bgra[..., [2, 1, 0, 3]]).view(np.uint32), want.view(np.uint32))


def test_save_exr_compresses_and_refuses(tmp_path):
    """ZIP must actually shrink a smooth image (the match search and the fixed Huffman code work),
    noise falls back to stored chunks without growing by more than the chunk headers, and bad
    arguments return 1 without writing."""
    smooth = np.zeros((64, 256, 4), np.float32)
    smooth[..., :3] = np.linspace(0, 1, 256, dtype=np.float32)[None, :, None]
    smooth[..., 3] = 1.0
    p = str(tmp_path / "smooth.exr")
    assert sp.save_exr(p, smooth, sp.EXR_FLOAT, sp.EXR_ZIP)
    assert os.path.getsize(p) < smooth.nbytes // 8
    assert np.array_equal(sp.load_exr(p), smooth)
    noise = np.random.RandomState(1).rand(32, 64, 4).astype(np.float32).view(np.uint32)
    noise = (noise ^ np.random.RandomState(2).randint(0, 2 ** 23, noise.shape).astype(np.uint32)).view(np.float32)
    p = str(tmp_path / "noise.exr")
    assert sp.save_exr(p, noise, sp.EXR_FLOAT, sp.EXR_ZIPS)
    assert os.path.getsize(p) <= noise.nbytes + 32 * 16 + 512
    assert np.array_equal(sp.load_exr(p).view(np.uint32), noise.view(np.uint32))
    assert not sp.save_exr(str(tmp_path / "no" / "dir.exr"), smooth)
    assert not sp.save_exr(p, smooth, 0, sp.EXR_ZIP)
    assert not sp.save_exr(p, smooth, sp.EXR_FLOAT, 4)          # PIZ: not written


def test_save_ppm(tmp_path):
    rgba8 = (np.arange(7 * 5, dtype=np.uint32).reshape(5, 7) * 0x01030507 + 0xFF000000) & 0xFFFFFFFF
    p = str(tmp_path / "out.ppm")
    sp.write_ppm(p, rgba8)
    data = open(p, "rb").read()
    assert data.startswith(b"P6\n7 5\n255\n")
    body = np.frombuffer(data[len(b"P6\n7 5\n255\n"):], np.uint8).reshape(5, 7, 3)
    for ch in range(3):
        assert np.array_equal(body[..., ch], (rgba8 >> (8 * ch)) & 0xFF)
    with pytest.raises(OSError):
        sp.write_ppm(str(tmp_path / "no" / "dir.ppm"), rgba8)


@pytest.mark.parametrize("pixel_type", ["half", "float"])
@pytest.mark.parametrize("compression", ["none", "zips", "zip"])
@pytest.mark.parametrize("shape,tile", [((45, 70), (32, 32)), ((45, 70), (16, 8)), ((1, 1), (64, 64)),
                                        ((33, 129), (128, 16)), ((64, 64), (64, 64))])
def test_tiled_exr_round_trip(tmp_path, pixel_type, compression, shape, tile):
    """Single-part tiled files (what LoadEXR also reads, tinyexr.h:5905-5990): written by
    sp_b200_SaveExrImageTiled, read back by the library's LoadExrImage, by the reference's loader and
    by OpenCV; ragged right / bottom tiles, tiles larger than the image."""
    img = _test_image(shape[0], shape[1], seed=shape[0] * 17 + tile[0])
    want = img.astype(np.float16).astype(np.float32) if pixel_type == "half" else img
    path = str(tmp_path / "tiled.exr")
    assert sp.save_exr(path, img, {"half": sp.EXR_HALF, "float": sp.EXR_FLOAT}[pixel_type],
                       {"none": sp.EXR_NONE, "zips": sp.EXR_ZIPS, "zip": sp.EXR_ZIP}[compression], tile=tile)
    got = sp.load_exr(path)
    assert got is not None and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ref = _reference_exr_loader()
    if ref is not None:
        r = ref(path)
        assert r is not None and np.array_equal(r.view(np.uint32), want.view(np.uint32))
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    bgra = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert bgra is not None and bgra.shape == want.shape
    assert np.array_equal(np.ascontiguousarray(bgra[..., [2, 1, 0, 3]]).view(np.uint32), want.view(np.uint32))


def test_tiled_exr_damaged_files(tmp_path):
    """Truncated and corrupted tiled files return 1 or decode without a fault; a tile header that
    points outside the image or at another level is refused."""
    img = _test_image(40, 50, seed=5)
    path = str(tmp_path / "t.exr")
    assert sp.save_exr(path, img, sp.EXR_FLOAT, sp.EXR_ZIP, tile=(16, 16))
    data = open(path, "rb").read()
    for cut in (10, 200, len(data) // 2, len(data) - 3):
        p = tmp_path / "cut.exr"
        p.write_bytes(data[:cut])
        assert sp.load_exr(str(p)) is None
    for k in range(0, len(data), 53):
        broken = bytearray(data)
        broken[k] ^= 0xFF
        p = tmp_path / "flip.exr"
        p.write_bytes(bytes(broken))
        sp.load_exr(str(p))
    # first chunk's tile x -> 1000, then its level -> 1
    first = int.from_bytes(data[data.index(b"tiledesc\0") + 9 + 4 + 9 + 1:][:8], "little")
    for field, value in ((0, 1000), (8, 1)):
        broken = bytearray(data)
        broken[first + field:first + field + 4] = value.to_bytes(4, "little")
        p = tmp_path / "hdr.exr"
        p.write_bytes(bytes(broken))
        assert sp.load_exr(str(p)) is None
    assert not sp.save_exr(path, img, sp.EXR_FLOAT, sp.EXR_ZIP, tile=(0, 16))
