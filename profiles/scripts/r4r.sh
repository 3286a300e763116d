cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "folded or stem_direct" 2>&1 | grep -E "folded|stem routes|passed|failed|Error" | tail -10
