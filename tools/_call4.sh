mkdir -p gpurun_out
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-torch-baseline --no-batch2"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'head_points|_sp_kernel|gn_stats_s_kernel|gn_apply_s_kernel|gn_bwd_stats_s_kernel|gn_bwd_apply_s_kernel|dec_grad|loss_points_kernel' -c 120 --csv --log-file gpurun_out/c4_ncu.csv python bench.py $B > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/c4_ncu.log
