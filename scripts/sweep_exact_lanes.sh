#!/bin/bash
# exact kernel: lanes per walk as a function of the undecided lists' lengths (TM_EXACT_LANES=far_warp,far_wide,near_wide)
for v in "4096,100000,131072" "4096,400000,131072" "4096,100000,500000" "4096,400000,500000" "0,0,0" "16384,100000,131072"; do
  echo "TM_EXACT_LANES=$v"
  TM_EXACT_LANES=$v TM_DIRECT=0 python scripts/bench_floor.py --sizes ${1:-1000000,2500000,10000000} --out gpurun_out/tmp_floor.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(' ', r['points'], round(r['ms'],4), r['phases_ms']['evaluate'])
"
done
