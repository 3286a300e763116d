cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "enhance" 2>&1 | grep -E "err|vs |passed|failed|Error|diff" | tail -8
timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 classes', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
FCP_UPCONV_MATERIALIZE=1 timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 materialized', d['value'], d['ms_per_step'], d['roofline']['achieved'])"
