set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -40 > gpurun_out/r2c_pytest.log
cat gpurun_out/r2c_pytest.log
FCP_TRACE=1 timeout 900 python bench.py > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_trace.log
tail -c 6000 gpurun_out/r2c_bench.log
for ab in 1 4 5; do
FCP_TC_ABLATE=$ab FCP_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 0 --no-secondary > gpurun_out/r2c_bench_ab$ab.log 2> gpurun_out/r2c_trace_ab$ab.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c_bench_ab$ab.log").read().strip().splitlines()[-1])
print("ablate $ab", d["value"], d["roofline"]["achieved"], d["clocks"])
PY
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_ref.log 2>&1
tail -c 1500 gpurun_out/r2c_ref.log
