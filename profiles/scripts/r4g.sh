cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export FCP_TC_PAIR=2
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv2d" 2>&1 | tail -15
echo "rc=$?"
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
