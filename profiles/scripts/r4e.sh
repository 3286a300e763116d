cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export FCP_TC_PAIR=2
FCP_LOG_CONV=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv2d" 2>&1 | grep -v "^\[conv_tc [0-9]" | tail -15
timeout 300 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -5
export FCP_TC_PAIR=1
for i in 1; do
timeout 600 python bench.py --steps 4 --warmup 3 --no-secondary --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 pair', d['value'], d['roofline']['achieved'])"
timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 pair', d['value'], d['roofline']['achieved'])"
done
