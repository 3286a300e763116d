cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r4l_pytest.log 2>&1; tail -3 gpurun_out/r4l_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_ref.jsonl 2> gpurun_out/r2b_ref.err; tail -1 gpurun_out/r2b_ref.jsonl | cut -c1-300
