set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "passed|failed|error|relative max|max \|err\|" > gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_pytest.log
for ab in 0 1 4 5; do
FCP_TC_ABLATE=$ab FCP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 --conv-impl 2 > gpurun_out/r2b_bench2_ab$ab.log 2> gpurun_out/r2b_trace2_ab$ab.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2b_bench2_ab$ab.log").read().strip().splitlines()[-1])
print("ablate $ab", d["value"], d["roofline"]["achieved"], d["clocks"])
PY
done
FCP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 --conv-impl 1 > gpurun_out/r2b_bench1.log 2> gpurun_out/r2b_trace1.log
tail -c 900 gpurun_out/r2b_bench1.log
