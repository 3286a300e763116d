cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
FCP_TRACE=1 timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/r4b_c4_trace.txt 2>&1
grep "fcp trace" gpurun_out/r4b_c4_trace.txt | head -30
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary --cpu-sample 0 2>&1 | tail -1 | tee gpurun_out/r4b_c3.txt
