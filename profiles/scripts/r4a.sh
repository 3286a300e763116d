cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu -k "conv2d or enhance or pipeline_mixed" -s 2>&1 | tail -40 > gpurun_out/r4a_tests.txt
cat gpurun_out/r4a_tests.txt
timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -3 | tee gpurun_out/r4a_c4.txt
FCP_RRDB_LAYERWISE=1 timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 2>&1 | tail -3 | tee gpurun_out/r4a_c4_layerwise.txt
