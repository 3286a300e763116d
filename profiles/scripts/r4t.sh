cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 400 python bench.py --cpu-sample 4 > gpurun_out/r2b_final_bench.jsonl 2> gpurun_out/r2b_final_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2b_final_bench.jsonl").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"])
print("secondary", {k: round(v["value"],1) for k, v in d["secondary"].items()})
print("cpu", d.get("cpu_baseline", {}).get("value"), "traffic", d["roofline"]["traffic"])
PY
