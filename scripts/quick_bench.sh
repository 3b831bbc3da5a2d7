timeout 300 python bench.py --skip-cpu --skip-brute --skip-e2e > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -3 gpurun_out/q_bench.err; python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print(d['value'], d['ms_per_step']); print(d['phases_ms']); print({k:d['stats'][k] for k in ('pairs_evaluated','cull_tests','points_slow','points_far','points_ring')})"
