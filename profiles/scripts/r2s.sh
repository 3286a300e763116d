set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2s_pytest.log 2>&1
grep -E "as_batch \(float|stem routes|passed|failed|FAILED|Error" gpurun_out/r2s_pytest.log | head
timeout 1200 python profiles/process_dir_bench.py --images 512 > gpurun_out/r2s_process_dir.jsonl 2> gpurun_out/r2s_process_dir.err
cat gpurun_out/r2s_process_dir.jsonl; tail -3 gpurun_out/r2s_process_dir.err
