cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3k_pytest.log 2>&1; tail -3 gpurun_out/r3k_pytest.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.jsonl 2> gpurun_out/r2_final_ref.err
FCP_TRACE=1 timeout 900 python bench.py > gpurun_out/r2_final_bench.jsonl 2> gpurun_out/r2_final_trace.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_final_bench.jsonl").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"])
print("secondary", {k: round(v["value"],1) for k, v in d["secondary"].items()})
print("cpu", d.get("cpu_baseline", {}).get("value"))
PY
python __graft_entry__.py --smoke 2>&1 | tail -1
