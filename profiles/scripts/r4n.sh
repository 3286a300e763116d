cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_determinism.py -x -q -m gpu -s -k "enhance or pipeline_mixed or determin or reproduc" 2>&1 | grep -E "err|vs |passed|failed|Error" | tail -10
timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
