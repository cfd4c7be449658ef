#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report: per kernel, dram__bytes_read.sum + dram__bytes_write.sum of ONE
launch (what bench.py copies into roofline.traffic) together with the issue statistics of the same launch.
usage: python profiles/make_traffic_json.py gpurun_out/<report>.ncu-rep <tile e.g. 8x16> <config c2|c4> [summary file name]"""
import csv
import io
import json
import os
import subprocess
import sys

STAGE = {"raster_forward_kernel": "lgs_rasterize_forward_packed", "raster_backward_v2_kernel": "lgs_rasterize_backward",
         "raster_backward_kernel": "lgs_rasterize_backward", "project_forward_kernel": "lgs_project_forward",
         "project_backward_kernel": "lgs_project_backward", "emit_pairs_rec_kernel": "lgs_emit_pairs_u16"}


def main():
    rep, tile, config = sys.argv[1], sys.argv[2], sys.argv[3]
    src = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H = rows[0]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    out["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` captures on B200; "
                       "bench.py copies the entry of its dominant stage into roofline.traffic when tile and config match. "
                       "Regenerate with profiles/make_traffic_json.py after a kernel changes.")

    def col(r, name):
        return float(r[H.index(name)].replace(",", "")) if name in H else None
    for r in rows[2:]:
        name = r[H.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").strip()
        stage = STAGE.get(name)
        if stage is None:
            continue
        rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
        unit_r = rows[1][H.index("dram__bytes_read.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_r, 1)
        out[stage] = {"tile": tile, "config": config, "kernel": name, "bytes": int((rd + wr) * scale),
                      "warp_instructions": int(col(r, "smsp__inst_executed.sum")),
                      "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "duration_us_under_ncu": col(r, "gpu__time_duration.sum"), "source": f"profiles/{src}"}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
