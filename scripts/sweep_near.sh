#!/bin/bash
# step / phase times against the tile parameters (D_near factor, D_max factor, voxel edge); experiments only
for cfg in "0.25 0.8 2.0" "0.25 0.65 2.0" "0.25 0.5 2.0" "0.2 0.8 2.0" "0.2 0.65 2.5" "0.2 0.8 2.5" "0.22 0.7 2.3"; do
  set -- $cfg
  TM_NEAR_FACTOR=$2 TM_REACH_FACTOR=$3 timeout 300 python bench.py --skip-cpu --skip-brute --skip-e2e --cell $1 --steps 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); p=d['phases_ms']; s=d['stats']
print('cell $1 near $2 reach $3', 'step %.3f'%d['ms_per_step'], 'evaluate %.3f'%p['evaluate'], 'tree %.3f'%p['tree'], 'scan %.3f'%p['scan'], 'bounds/pt %.1f'%(s['bound_tests']/1e7), 'slow', s['points_slow'], 'far', s['points_far'], 'ring', s['points_ring'])"
done
