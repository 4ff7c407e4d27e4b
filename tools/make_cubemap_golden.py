#!/usr/bin/env python
"""Generate tests/golden/g4_cubemap.npz from the UNMODIFIED reference's environment bakers
(src/cubemap.cpp compiled into oracle/_ref/libspref*.so; see oracle/ref_driver.cpp).

Run HERE (where /root/reference is mounted and `make -C oracle ref` works):

    python tools/make_cubemap_golden.py

The fixture holds the results only; tests rebuild the input map from
vk_cinematic_b200.workloads.make_env_map (seeded) with the recorded parameters.  `*_dm` arrays
come from the deterministic-math build (libm float calls evaluated in double and rounded once),
the others from glibc's float functions.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ora  # noqa: E402
from vk_cinematic_b200 import workloads as W  # noqa: E402

ENV = (96, 48, "kiara")
CUBE = (12, 10)      # width, height (not square on purpose)
IRRADIANCE = (5, 4)
SPP = 32


def main():
    assert ora.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    env = W.make_env_map(*ENV)
    out = {"env_params": np.array([ENV[0], ENV[1]], np.uint32), "env_variant": ENV[2],
           "cube_size": np.array(CUBE, np.uint32), "irradiance_size": np.array(IRRADIANCE, np.uint32),
           "spp": np.uint32(SPP), "env_checksum": np.uint64(env.view(np.uint32).sum(dtype=np.uint64))}
    for tag, lib in (("", ora.load_ref()), ("_dm", ora.load_ref_dm())):
        out["cube" + tag] = lib.create_cube_map(env, *CUBE)
        out["irradiance_uniform" + tag] = lib.create_irradiance_cube_map(env, *IRRADIANCE, spp=SPP, sampling=0)
        out["irradiance_random" + tag] = lib.create_irradiance_cube_map(env, *IRRADIANCE, spp=SPP, sampling=1)
    path = os.path.join(ROOT, "tests", "golden", "g4_cubemap.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
