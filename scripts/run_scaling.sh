#!/bin/bash
# strong scaling of the default bench (and, with "sweep" as first argument, the configs[4] sweep) on the GPUs of one box;
# results under gpurun_out/
set -u
if [ "${1:-}" = "sweep" ]; then
  for w in 8 4 2; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29600 + w)) scripts/bench_sweep_multi.py --out gpurun_out/r02_sweep_w$w.json 2> gpurun_out/r02_sweep_w$w.err | tail -8
  done
  python scripts/bench_sweep_multi.py --out gpurun_out/r02_sweep_w1.json 2> gpurun_out/r02_sweep_w1.err | tail -8
fi
for w in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29700 + w)) bench.py --gpus $w --steps 10 --warmup 3 > gpurun_out/r02_scale_n$w.json 2> gpurun_out/r02_scale_n$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n$w.json')); print($w, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_pinned']['value'], d['weak']['value'], d['sharded_parity']['sharded_vs_single_gpu_bitwise'], d['table_broadcast_ms'])"
done
python bench.py --skip-cpu > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n1.json')); print(1, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_pinned']['value'])"
