#!/bin/bash
# A/B of the round-2 backward kernel and tile order on the C2 bench workload (run on a B200 through gpurun).
# usage: bash profiles/ab_r2.sh  -> gpurun_out/ab_r2.txt
out=gpurun_out/ab_r2.txt
: > $out
for cfg in "LGS_BWD=v1 LGS_TILE_ORDER=0" "LGS_BWD=v1 LGS_TILE_ORDER=1" "LGS_BWD=v2 LGS_TILE_ORDER=0" "LGS_BWD=v2 LGS_TILE_ORDER=1"; do
  echo "== $cfg" >> $out
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
st=d['stages']
print('value %.1f views/s  ms/view %.4f (serial %.4f)' % (d['value'], d['path_roofline']['ms_per_view'], d['path_roofline']['ms_per_view_single_stream']))
for k,v in sorted(st.items(), key=lambda kv:-kv[1]['ms_per_view']): print('   %-34s %.4f ms' % (k, v['ms_per_view']))
" >> $out
done
cat $out
