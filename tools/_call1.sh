set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py -m gpu -x -q > gpurun_out/c1_cli.log 2>&1; echo "cli rc=$?" >> gpurun_out/c1_cli.log
B="--steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-torch-baseline --no-batch2"
for k in 2 3 4 6 2; do
  MDC_GN_BPS=$k timeout 300 python bench.py $B > gpurun_out/c1_bps_$k.json 2> gpurun_out/c1_bps_$k.err; echo "bps $k rc=$?"
  python -c "import json,sys; d=json.loads(open('gpurun_out/c1_bps_$k.json').read().strip().splitlines()[-1]); print('BPS',$k,d['ms_per_step'])"
done
timeout 900 python bench.py > gpurun_out/c1_bench_b.json 2> gpurun_out/c1_bench_b.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c1_cli.log
