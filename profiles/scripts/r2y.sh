set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2y_pytest.log 2>&1; tail -4 gpurun_out/r2y_pytest.log
python profiles/diag_process_dir.py 2>&1 | tail -4
FCP_TRACE=1 timeout 900 python bench.py > gpurun_out/r2y_bench.log 2> gpurun_out/r2y_trace.log
tail -c 5000 gpurun_out/r2y_bench.log
