cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for ab in 0 128; do
echo "== ablate $ab"
FCP_TC_ABLATE=$ab timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['value'], d['roofline']['achieved'])"
FCP_TC_ABLATE=$ab timeout 600 python bench.py --steps 4 --warmup 3 --no-secondary --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', d['value'], d['roofline']['achieved'])"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_determinism.py -x -q -m gpu 2>&1 | tail -3
