set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "conv2d" > gpurun_out/r3c_conv.log 2>&1; tail -3 gpurun_out/r3c_conv.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r3c_pytest.log 2>&1; tail -3 gpurun_out/r3c_pytest.log
FCP_TRACE=1 timeout 900 python bench.py --cpu-sample 0 --no-secondary > gpurun_out/r3c_bench.log 2> gpurun_out/r3c_trace.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r3c_bench.log").read().strip().splitlines()[-1])
print("bench halo", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"])
PY
grep -E " k3 " gpurun_out/r3c_trace.log | sort -t= -k3 -n -r | head -10
FCP_NO_HALO=1 FCP_TRACE=1 timeout 900 python bench.py --cpu-sample 0 --no-secondary --steps 3 > gpurun_out/r3c_bench_nohalo.log 2> gpurun_out/r3c_trace_nohalo.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r3c_bench_nohalo.log").read().strip().splitlines()[-1])
print("bench no halo", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"])
PY
timeout 600 python bench.py --config c4 --steps 2 --warmup 1 > gpurun_out/r3c_c4.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r3c_c4.log").read().strip().splitlines()[-1])
print("c4", round(d["value"],1), "img/s", round(d["roofline"]["achieved"],1), "TFLOP/s")
PY
