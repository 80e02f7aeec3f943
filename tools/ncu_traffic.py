#!/usr/bin/env python
"""DRAM bytes per launch of every stage of one step, from an `ncu --set full` capture of
`python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline` (bench.py reads the result as `roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep > profiles/ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

STAGE_OF = [("project_kernel", "pre_bin"), ("bin_walk_kernel", "pre_bin"), ("bin_scan_kernel", "pre_bin"), ("scan_kernel", "pre_bin"),
            ("emit_kernel", "pre_bin"), ("bucket_sort_kernel", "bin_sort"), ("lsd_sort_kernel", "bin_sort"), ("onesweep_pass_kernel", "sort_passes"),
            ("tile_ranges_kernel", "ranges"), ("composite_fwd_kernel", "comp_fwd"), ("composite_bwd_kernel", "comp_bwd"),
            ("preprocess_bwd_kernel", "pre_bwd")]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    out, kernels = {}, {}
    for r in rows[2:]:
        name = r[ik]
        stage = next((s for k, s in STAGE_OF if k in name), None)
        if stage is None:
            continue
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        out[stage] = out.get(stage, 0.0) + b
        kernels.setdefault(stage, []).append(name.split("(")[0])
    # the capture holds ONE step: a stage's kernels are summed (pre_bin = projection + count + scan + scatter; bin_sort = its size classes)
    res = {s: {"dram_bytes_per_launch": round(v), "kernels": kernels[s]} for s, v in out.items()}
    res["_source"] = ("ncu --set full --clock-control none, python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline (C2T, 1 GPU), "
                      "dram__bytes_read.sum + dram__bytes_write.sum, summed over the kernels of each stage of one step")
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
