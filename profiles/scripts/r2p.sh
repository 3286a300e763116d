set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2p_pytest.log 2>&1; tail -3 gpurun_out/r2p_pytest.log
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_target.py $( [ $tool = memcheck ] && echo --pipeline ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported" gpurun_out/r2_sanitizer_$tool.log | head -10
done
timeout 600 python bench.py --steps 4 --warmup 2 --cpu-sample 0 --no-secondary > gpurun_out/r2p_bench.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r2p_bench.log").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"])
PY
