cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for pr in 1 0; do
echo "== pair $pr"
FCP_TC_PAIR=$pr FCP_LOG_CONV=1 FCP_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-secondary --cpu-sample 0 > gpurun_out/r4f_$pr.txt 2>&1
grep "pairs resident" gpurun_out/r4f_$pr.txt | head -2
grep "fcp trace" gpurun_out/r4f_$pr.txt | grep -E "cin256  cout256|cin128  cout128|cin256  cout1024|cin512  cout512|cin128  cout512|cin1024 cout256" | head -12
done
