set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2l_pytest.log 2>&1
grep -E "detect 1024|passed|failed|FAILED|Error" gpurun_out/r2l_pytest.log | head
for ab in 0 32 8; do
FCP_TC_ABLATE=$ab timeout 600 python bench.py --steps 4 --warmup 2 --cpu-sample 0 --no-secondary > gpurun_out/r2l_bench_ab$ab.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r2l_bench_ab$ab.log").read().strip().splitlines()[-1])
print("ablate $ab", round(d["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"])
PY
done
