#!/bin/bash
# Everything the profiles/ index of a round is refreshed from, in one gpurun call (one B200, about 15 minutes):
#   tools/gpurun_retry.sh --timeout 2400 -- 'bash tools/round_end_gpu.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py > gpurun_out/bench_config_b.json 2> gpurun_out/bench_config_b.err; echo "bench b rc=$?"
timeout 600 python bench.py --config c --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_config_c.json 2> gpurun_out/bench_config_c.err; echo "bench c rc=$?"
timeout 600 python bench.py --config b640 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_config_b_res640.json 2> gpurun_out/bench_config_b_res640.err; echo "bench b640 rc=$?"
timeout 600 python bench.py --vae light --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_vae_light.json 2> gpurun_out/bench_vae_light.err; echo "bench light rc=$?"
OPS_CSV=gpurun_out/ops_profile.csv timeout 600 python tools/gpu_profile_step.py > gpurun_out/ops_profile.log 2>&1; echo "ops rc=$?"
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-torch-baseline --no-batch2"
timeout 300 python bench.py $B > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv --log-file gpurun_out/ncu_launches_bench.csv python bench.py $B > gpurun_out/ncu_launches_bench.log 2>&1; echo "ncu rc=$?"
gzip -f gpurun_out/ncu_launches_bench.csv
for f in b c b_res640; do python -c "
import json
d=json.loads(open('gpurun_out/bench_config_$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], (d.get('e2e') or {}).get('sec_per_frame'), d['roofline']['frac'], d['clocks'])"; done
