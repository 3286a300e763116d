cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for rep in 1 2; do
for v in "" "FCP_B200_LIB=$PWD/face_crop_plus_b200/libfcpb200_prehalo.so"; do
env $v timeout 600 python bench.py --cpu-sample 0 --no-secondary --steps 4 > gpurun_out/ab.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
print("c3 [${v:0:20}] rep $rep:", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"])
PY
env $v timeout 600 python bench.py --config c4 --steps 2 --warmup 1 > gpurun_out/ab4.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/ab4.log").read().strip().splitlines()[-1])
print("c4 [${v:0:20}] rep $rep:", round(d["value"],1), "conv", round(d["roofline"]["achieved"],1))
PY
done; done
