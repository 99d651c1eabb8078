mkdir -p gpurun_out
B="--steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-torch-baseline --no-batch2"
for k in dense sparse dense sparse; do
  if [ $k = dense ]; then export MDC_NO_SPARSEHEAD=1; else unset MDC_NO_SPARSEHEAD; fi
  timeout 300 python bench.py $B > gpurun_out/c5_$k.json 2> gpurun_out/c5_$k.err; echo "$k rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/c5_$k.json').read().strip().splitlines()[-1]); print('BENCH','$k',d['ms_per_step'],d['roofline']['frac'],d['clocks'])"
done
unset MDC_NO_SPARSEHEAD
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 8 gpurun_out/c5_pytest_gpu.log
