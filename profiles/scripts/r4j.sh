cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r4j_pytest.log 2>&1; tail -3 gpurun_out/r4j_pytest.log
FCP_TRACE=1 timeout 900 python bench.py > gpurun_out/r2b_final_bench.jsonl 2> gpurun_out/r2b_final_trace.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2b_final_bench.jsonl").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"])
print("secondary", {k: round(v["value"],1) for k, v in d["secondary"].items()})
print("cpu", d.get("cpu_baseline", {}).get("value"))
PY
FCP_TRACE=1 timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/r2b_c4_bench.jsonl 2> gpurun_out/r2b_c4_trace.log
grep "fcp trace" gpurun_out/r2b_c4_trace.log | head -12
FCP_LOG_CONV=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:conv_tc_kernel -s 520 -c 210 --csv --log-file gpurun_out/r2b_c4_launches.csv python bench.py --config c4 --batch 8 --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r2b_c4_ncu.log 2> gpurun_out/r2b_c4_ncu_shapes.log
tail -2 gpurun_out/r2b_c4_launches.csv | cut -c1-200
python __graft_entry__.py --smoke 2>&1 | tail -1
du -sh gpurun_out
