#!/usr/bin/env python
"""DRAM traffic of the kernels captured in an .ncu-rep (ncu --set full): per launch and summed.
Writes profiles/trace_traffic.json, which bench.py reports as roofline.traffic.
usage: python tools/ncu_traffic.py rep.ncu-rep workload spp width [out.json] [traced rays per step]
(the traced-ray count, printed by `bench.py --quick` as traced_rays_per_step, lets bench.py scale the
per-launch DRAM bytes to a strip of the frame)"""
import csv
import json
import subprocess
import sys


def main():
    rep, workload, spp, width = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    out = sys.argv[5] if len(sys.argv) > 5 else "profiles/trace_traffic.json"
    traced = float(sys.argv[6]) if len(sys.argv) > 6 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    launches = []
    for r in rows[2:]:
        def val(name):
            i = col[name]
            return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
        launches.append({"kernel": r[col["Kernel Name"]], "ms_under_ncu": t,
                         "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum")})
    total = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches)
    doc = {"workload": workload, "spp": spp, "width": width, "launches": len(launches),
           "dram_bytes_per_step": total, "dram_bytes_per_launch": total / max(1, len(launches)),
           "traced_rays_per_step": traced,
           "source": rep.split("/")[-1] + " (ncu --set full --clock-control none, every k_trace launch of one frame)",
           "per_launch": launches}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in doc.items() if k != "per_launch"}))


if __name__ == "__main__":
    main()
