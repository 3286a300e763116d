cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_pair" 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_target.py $( [ $tool = memcheck ] && echo --pipeline ) > gpurun_out/r2b_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|conv |rrdbnet|pipeline faces|Error|hazard" gpurun_out/r2b_sanitizer_$tool.log | head -24
done
