mkdir -p gpurun_out
for k in dense sparse; do
  if [ $k = dense ]; then export MDC_NO_SPARSEHEAD=1; else unset MDC_NO_SPARSEHEAD; fi
  timeout 900 python -m pytest tests/test_gpu_fullwidth.py -m gpu -q -s -k later_steps > gpurun_out/c6_$k.log 2>&1; echo "$k rc=$?"
  grep -E "loss\[|full width\]|passed|failed" gpurun_out/c6_$k.log | head -20
done
