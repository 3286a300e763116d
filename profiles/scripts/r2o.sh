set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for tool in memcheck synccheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_target.py $( [ $tool = memcheck ] && echo --pipeline ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|conv \(|pipeline faces|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -20
done
