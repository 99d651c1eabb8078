set -x
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-torch-baseline --no-batch2"
MDC_NO_SPARSEHEAD=1 timeout 300 python bench.py $B > gpurun_out/c3_dense.json 2> gpurun_out/c3_dense.err; echo "dense rc=$?"
timeout 300 python bench.py $B > gpurun_out/c3_sparse.json 2> gpurun_out/c3_sparse.err; echo "sparse rc=$?"
python - <<'PY'
import json
for k in ("dense","sparse"):
    try:
        d=json.loads(open(f"gpurun_out/c3_{k}.json").read().strip().splitlines()[-1]); print("BENCH",k,d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"])
    except Exception as e: print("BENCH",k,"failed",e)
PY
timeout 1200 python -m pytest tests/test_gpu_switches.py -k sparse_output_head -x -q > gpurun_out/c3_head.log 2>&1; echo "head rc=$?"
tail -n 30 gpurun_out/c3_head.log
