#!/bin/bash
# A/B on the C2 bench workload: forward packed pairs, tile shapes.  usage: bash profiles/ab_r2b.sh -> gpurun_out/ab_r2b.txt
out=gpurun_out/ab_r2b.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-e2e $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
st=d['stages']
print('value %.1f views/s  ms/view %.4f (serial %.4f)  D %d' % (d['value'], d['path_roofline']['ms_per_view'], d['path_roofline']['ms_per_view_single_stream'], d['workload_stats']['D']))
print('   ' + '  '.join('%s=%.4f' % (k.replace('lgs_','').replace('rasterize_','r').replace('_packed',''), v['ms_per_view']) for k,v in sorted(st.items(), key=lambda kv:-kv[1]['ms_per_view'])[:6]))
" >> $out
}
EXTRA="" run LGS_FWD_PAIRS=0
EXTRA="" run LGS_FWD_PAIRS=1
EXTRA="--tile 12x16" run LGS_FWD_PAIRS=0
EXTRA="--tile 12x16" run LGS_FWD_PAIRS=1
EXTRA="--tile 16x16" run LGS_FWD_PAIRS=1
cat $out
