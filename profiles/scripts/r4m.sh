cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "enhance or pipeline_mixed" 2>&1 | tail -8
timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 direct conv_last', d['value'], d['ms_per_step'])"
FCP_CONV_LAST_GEMM=1 timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 gemm conv_last', d['value'], d['ms_per_step'])"
