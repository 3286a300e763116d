cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for mb in "8 32" "16 32" "32 64" "64 128" "32 256" "128 256"; do
set -- $mb
timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 0 --no-secondary --det-mb $1 --par-mb $2 > gpurun_out/r2z_mb_$1_$2.log 2>&1
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2z_mb_$1_$2.log").read().strip().splitlines()[-1])
    print("det_mb $1 par_mb $2:", round(d["value"],1), "img/s, e2e", round(d["e2e"]["value"],1), "conv TFLOP/s", round(d["roofline"]["achieved"],1), "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("det_mb $1 par_mb $2: failed", e)
PY
done
