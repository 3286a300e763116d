cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_determinism.py -x -q -m gpu -s -k "detect or pipeline or determin or smoke" 2>&1 | grep -E "witness|err|passed|failed|Error|px" | tail -10
for f in 0 1; do
[ $f = 1 ] && export FCP_NO_FUSE_SHORTCUT=1
timeout 600 python bench.py --steps 4 --warmup 3 --no-secondary --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 nofuse=$f', d['value'], d['roofline']['achieved'], d['stages_ms']['detect_net'])"
done
