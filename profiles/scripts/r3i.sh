cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for rep in 1 2; do
for v in "" "FCP_NO_HALO=1"; do
env $v timeout 600 python bench.py --cpu-sample 0 --no-secondary --steps 4 > gpurun_out/ab.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
print("c3 [$v] rep $rep:", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"])
PY
env $v timeout 600 python bench.py --config c4 --steps 2 --warmup 1 > gpurun_out/ab4.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/ab4.log").read().strip().splitlines()[-1])
print("c4 [$v] rep $rep:", round(d["value"],1), "conv", round(d["roofline"]["achieved"],1))
PY
done; done
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv
