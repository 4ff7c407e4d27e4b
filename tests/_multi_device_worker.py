"""Worker of tests/test_gpu_parity.py::test_multi_device_frame: one process, several devices through
libspb200's own multi-device path (sp_b200_InitDevices / sp_b200_InitDeviceList).

    python tests/_multi_device_worker.py <device list, comma separated> <mode: host|device>

Renders the same frames on ONE device first (the same library, before sp_b200_InitDeviceList) and then
over all listed devices; prints one JSON line with whether the images are bit-identical, the counters,
the strips of every frame and the per-device kernel times."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from vk_cinematic_b200 import sp, workloads as W  # noqa: E402


def main():
    devices = [int(x) for x in sys.argv[1].split(",")]
    mode = sys.argv[2] if len(sys.argv) > 2 else "host"
    width, height, spp, bounces = 640, 368, 8, 4
    wl = W.config1(width, height, env_size=(512, 256))
    # off-centre so that an even split is unbalanced and the re-cut has something to do
    px, py, pz = wl.camera_position
    wl.camera_position = (px + 0.02, py + 0.06, pz)
    frames = 4
    assert sp.lib.sp_b200_Init(devices[0]) == 0
    host = torch.zeros((height, width, 4), dtype=torch.float32).pin_memory()
    r = sp.Renderer(devices[0]).load_workload(wl, pixels=host.numpy())
    sp.set_params(samplesPerPixel=spp, bounceCount=bounces, tileWidth=64, tileHeight=16)
    single, single_m = [], []
    for f in range(frames):
        img, m = r.render_frame(frame=f)
        single.append(img.copy())
        single_m.append(m.copy())
    r.close()
    sp.lib.sp_b200_Shutdown()

    dl = (C.c_int32 * len(devices))(*devices)
    n = sp.lib.sp_b200_InitDeviceList(dl, len(devices))
    assert n == len(devices) and sp.lib.sp_b200_DeviceCount() == n
    r = sp.Renderer(devices[0]).load_workload(wl, pixels=host.numpy())   # built AFTER InitDevices
    sp.set_params(samplesPerPixel=spp, bounceCount=bounces, tileWidth=64, tileHeight=16)
    out = {"devices": devices, "mode": mode, "identical": [], "metrics_equal": [], "strips": [], "kernel_ms": []}
    dev_image = torch.zeros((height, width, 4), dtype=torch.float32, device=f"cuda:{devices[0]}")
    for f in range(frames):
        host.zero_()
        if mode == "host":
            img, m = r.render_frame(frame=f)
        else:
            dev_image.zero_()
            mm = sp.sp_Metrics()
            assert sp.lib.sp_b200_RenderFrameToDevice(C.byref(r.ctx), f, dev_image.data_ptr(), C.byref(mm)) == 0
            torch.cuda.synchronize()
            img, m = dev_image.cpu().numpy(), np.array(list(mm.values), dtype=np.uint64)
        out["identical"].append(bool(np.array_equal(img.view(np.uint32), single[f].view(np.uint32))))
        out["metrics_equal"].append(bool(np.array_equal(m[1:5], single_m[f][1:5])))
        strips, kms = [], []
        for i in range(n):
            st, b, e = sp.sp_b200_Stats(), C.c_uint32(), C.c_uint32()
            assert sp.lib.sp_b200_GetDeviceStats(i, C.byref(st), C.byref(b), C.byref(e)) == 0
            strips.append([b.value, e.value])
            kms.append(st.kernelMs)
        out["strips"].append(strips)
        out["kernel_ms"].append(kms)
    r.close()
    sp.lib.sp_b200_Shutdown()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
