#!/bin/bash
# evaluate-phase time against the tile parameters (D_near factor, voxel edge); experiments only
for cell in 0.25 0.2 0.3; do for nf in 0.5 0.65 0.8 1.0; do
  TM_NEAR_FACTOR=$nf timeout 300 python bench.py --skip-cpu --skip-brute --skip-e2e --cell $cell --steps 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); p=d['phases_ms']; s=d['stats']
print('cell $cell near $nf', 'step %.3f'%d['ms_per_step'], 'evaluate %.3f'%p['evaluate'], 'tree %.3f'%p['tree'], 'bounds/pt %.1f'%(s['bound_tests']/1e7), 'slow', s['points_slow'], 'far', s['points_far'], 'ring', s['points_ring'])"
done; done
