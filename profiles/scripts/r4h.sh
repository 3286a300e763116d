cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
FCP_TC_PAIR=2 timeout 600 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for pr in 1 0; do
echo "== pair $pr"
FCP_TC_PAIR=$pr FCP_LOG_CONV=1 FCP_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-secondary --cpu-sample 0 > gpurun_out/r4h_$pr.txt 2>&1
grep "pairs resident" gpurun_out/r4h_$pr.txt | head -1
tail -1 gpurun_out/r4h_$pr.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', d['value'], d['roofline']['achieved'])"
grep "fcp trace" gpurun_out/r4h_$pr.txt | grep -E "cin256  cout256|cin128  cout128|cin256  cout1024|cin512  cout512|cin128  cout512|cin1024 cout256|cin64   cout256" | head -14
done
