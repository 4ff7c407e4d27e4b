"""__graft_entry__.smoke(): one small render through libspb200.so on cuda:0, checked against the
oracle (TEST INFRASTRUCTURE: the oracle is the checker here, never the thing measured)."""
import numpy as np

import ora
from vk_cinematic_b200 import sp, workloads as W


def run_smoke():
    import torch
    assert torch.cuda.is_available(), "smoke() needs a CUDA device; libspb200 has no CPU path"
    wl = W.config1(192, 128, env_size=(512, 256))
    r = sp.Renderer(0).load_workload(wl)
    sp.set_params(samplesPerPixel=2, bounceCount=3, cullByDistance=1, mathMode=0, envFilter=0,
                  radianceClamp=10.0)
    img, m = r.render_frame(frame=1)
    hits = r.primary_hits(sample=0, frame=1)
    stats = sp.last_stats()
    checker = ora.load_ref_dm() if ora.have_ref() else ora.load_port_dm()
    s = checker.scene().load_workload(wl)
    cimg, cm = s.render_seeded(spp=2, bounces=3, frame=1)
    chits = s.primary_hits(sample=0, frame=1)
    s.close()
    r.close()
    bit_diff = int((img.view(np.uint32) != cimg.view(np.uint32)).any(axis=2).sum())
    tri_diff = int((hits["tri"] != chits["tri"]).sum())
    print(f"smoke: {wl.width}x{wl.height} 2spp vs {checker.name}: differing pixels {bit_diff}, "
          f"triangle-id mismatches {tri_diff}, rays {int(m[2])} (oracle {int(cm[2])}), "
          f"kernel {stats.kernelMs:.3f} ms")
    assert bit_diff == 0 and tri_diff == 0 and int(m[2]) == int(cm[2])
