#!/bin/bash
# Sweep of raster launch parameters on the C2 workload; prints per-stage ms/view. Usage: gpurun -- 'bash profiles/sweep.sh'
mkdir -p gpurun_out
for tile in 16x16 8x16 8x8; do for wpb in 4 2 1; do for st in bulk cpasync; do
  LGS_WPB=$wpb timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --tile $tile --staging $st > gpurun_out/sw.json 2>gpurun_out/sw.err || { echo "FAIL $tile $wpb $st"; tail -3 gpurun_out/sw.err; continue; }
  python - "$tile" "$wpb" "$st" <<'PY'
import json,sys
d=json.load(open('gpurun_out/sw.json'))
s=d['stages']
g=lambda k: s.get(k,{}).get('ms_per_view',0)
print(f"tile {sys.argv[1]:6s} wpb {sys.argv[2]} {sys.argv[3]:8s} views/s {d['value']:7.1f} ms/view {d['path_roofline']['ms_per_view']:.3f} fwd {g('lgs_rasterize_forward_packed'):.3f} bwd {g('lgs_rasterize_backward'):.3f} tsort {g('lgs_sort_pairs_u32(tile)'):.3f} dsort {g('lgs_sort_pairs_u32(depth)'):.3f} emit {g('lgs_emit_pairs'):.3f} pf {g('lgs_project_forward'):.3f} pb {g('lgs_project_backward'):.3f} D {d['workload_stats']['D']}")
PY
done; done; done
