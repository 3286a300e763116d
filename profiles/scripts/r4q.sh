cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
B="python bench.py --batch 64 --det-mb 64 --par-mb 64 --steps 1 --warmup 1 --no-secondary --cpu-sample 0"
FCP_LOG_CONV=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 160 -c 330 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/r2b_ncu_bench.log 2> gpurun_out/r2b_ncu_shapes.log
tail -1 gpurun_out/r2b_ncu_bench.log | cut -c1-160
grep -c conv_tc gpurun_out/r2b_ncu_shapes.log
du -sh gpurun_out
