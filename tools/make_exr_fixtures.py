#!/usr/bin/env python
"""Small OpenEXR files for tests/test_assets.py, written with OpenCV's OpenEXR encoder (an
implementation independent of both the reference's tinyexr and the library's reader):

    OPENCV_IO_ENABLE_OPENEXR=1 python tools/make_exr_fixtures.py     # -> tests/golden/exr/

Each file comes with the array it was written from (.npy, RGBA order, float32; HALF files hold the
values after rounding to half, i.e. what any correct reader must return)."""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "exr")


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(7)
    T, C = cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_COMPRESSION
    F, H = cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_TYPE_HALF
    cases = [
        ("none_f32_rgb_13x9", (9, 13, 3), F, cv2.IMWRITE_EXR_COMPRESSION_NO),
        ("zip_f32_rgb_70x45", (45, 70, 3), F, cv2.IMWRITE_EXR_COMPRESSION_ZIP),     # 3 blocks of 16 lines, the last ragged
        ("zip_half_rgba_33x20", (20, 33, 4), H, cv2.IMWRITE_EXR_COMPRESSION_ZIP),
        ("zips_half_rgb_21x7", (7, 21, 3), H, cv2.IMWRITE_EXR_COMPRESSION_ZIPS),
        ("rle_f32_rgb_40x11", (11, 40, 3), F, cv2.IMWRITE_EXR_COMPRESSION_RLE),
        ("zip_f32_grey_17x5", (5, 17, 1), F, cv2.IMWRITE_EXR_COMPRESSION_ZIP),
        ("piz_half_rgb_16x8", (8, 16, 3), H, cv2.IMWRITE_EXR_COMPRESSION_PIZ),
        ("piz_half_rgb_70x45", (45, 70, 3), H, cv2.IMWRITE_EXR_COMPRESSION_PIZ),    # 2 blocks of 32 lines, the last ragged
        ("piz_f32_rgba_37x33", (33, 37, 4), F, cv2.IMWRITE_EXR_COMPRESSION_PIZ),
        ("pxr24_f32_rgb_16x8", (8, 16, 3), F, cv2.IMWRITE_EXR_COMPRESSION_PXR24),   # refused by the library
    ]
    for name, shape, typ, comp in cases:
        img = (rng.rand(*shape) ** 3 * 50).astype(np.float32)
        if name.startswith("rle"):
            img[:, 10:30] = 0.25                     # runs, so that RLE actually compresses
        if name.startswith("zip_f32_rgb") or name.startswith("piz_half_rgb_70"):
            img[10:30] = np.linspace(0, 3, shape[1], dtype=np.float32)[None, :, None]
        if typ == H:
            img = img.astype(np.float16).astype(np.float32)
        bgr = img[..., ::-1] if shape[2] == 3 else (img[..., [2, 1, 0, 3]] if shape[2] == 4 else img)
        path = os.path.join(OUT, name + ".exr")
        assert cv2.imwrite(path, np.ascontiguousarray(bgr), [T, typ, C, comp])
        rgba = np.ones(shape[:2] + (4,), np.float32)
        if shape[2] == 1:
            rgba[...] = img
        else:
            rgba[..., :shape[2]] = img
        np.save(os.path.join(OUT, name + ".npy"), rgba)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
