set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv2d" 2>&1 | tail -15 > gpurun_out/r2a_conv.log
cat gpurun_out/r2a_conv.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
FCP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 --conv-impl 2 > gpurun_out/r2a_bench2.log 2> gpurun_out/r2a_trace2.log
tail -1 gpurun_out/r2a_bench2.log
FCP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 --conv-impl 1 > gpurun_out/r2a_bench1.log 2> gpurun_out/r2a_trace1.log
tail -1 gpurun_out/r2a_bench1.log
