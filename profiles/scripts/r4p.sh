cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r4p_pytest.log 2>&1; tail -3 gpurun_out/r4p_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
