set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2q_pytest.log 2>&1; tail -5 gpurun_out/r2q_pytest.log
FCP_TRACE=1 timeout 600 python bench.py --steps 4 --warmup 2 --cpu-sample 0 --no-secondary > gpurun_out/r2q_bench.log 2> gpurun_out/r2q_trace.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2q_bench.log").read().strip().splitlines()[-1])
print("bench direct stem", round(d["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"], d["stages_ms"])
PY
grep "k7" gpurun_out/r2q_trace.log
FCP_STEM_ROWS=1 FCP_TRACE=1 timeout 600 python bench.py --steps 4 --warmup 2 --cpu-sample 0 --no-secondary > gpurun_out/r2q_bench_rows.log 2> gpurun_out/r2q_trace_rows.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2q_bench_rows.log").read().strip().splitlines()[-1])
print("bench rows stem", round(d["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"], d["stages_ms"])
PY
grep "k7" gpurun_out/r2q_trace_rows.log
