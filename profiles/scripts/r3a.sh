set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r3a_pytest.log 2>&1; tail -3 gpurun_out/r3a_pytest.log
timeout 900 python bench.py --cpu-sample 0 --no-secondary > gpurun_out/r3a_bench.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r3a_bench.log").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"], d["config"]["micro_batch"])
PY
