set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
B="python bench.py --batch 64 --det-mb 64 --par-mb 64 --steps 1 --warmup 1 --no-secondary --cpu-sample 0"
FCP_LOG_CONV=1 timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 165 -c 340 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_bench.log 2> gpurun_out/r2_ncu_shapes.log
tail -1 gpurun_out/r2_ncu_bench.log | cut -c1-200
FCP_LOG_CONV=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 158 -c 12 -o gpurun_out/r2_conv_tc $B > gpurun_out/r2_ncu_full.log 2> gpurun_out/r2_ncu_full_shapes.log
tail -1 gpurun_out/r2_ncu_full.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none -k regex:'stem_rows|maxpool3s2|global_avgpool|det_decode|det_nms|det_gather|warp_kernel|solve_kernel|parse_tail|parse_prep|unpack_faces|channel_affine|upsample2x|fc_kernel' -s 30 -c 30 -o gpurun_out/r2_other $B > gpurun_out/r2_ncu_other.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
FCP_TRACE=1 timeout 900 python bench.py > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_trace.log
tail -c 3000 gpurun_out/r2_final_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.log 2>&1
tail -c 600 gpurun_out/r2_final_ref.log
