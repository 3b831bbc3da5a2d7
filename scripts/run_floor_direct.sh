#!/bin/bash
# direct (TM_DIRECT=1) against sorted (TM_DIRECT=0) path, device time per call as the cloud grows: where is the crossover?
sizes=${1:-10000,100000,300000,450000,600000,1000000,1250000,2500000}
for d in 1 0; do echo TM_DIRECT=$d; TM_DIRECT=$d python scripts/bench_floor.py --sizes $sizes --out gpurun_out/floor_d$d.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['points'], round(r['ms'],4), r['phases_ms']['evaluate'], r['phases_ms']['total'])
"; done
