cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r4d_tests.txt
timeout 900 python bench.py 2>gpurun_out/r4d_bench.err | tail -1 > gpurun_out/r4d_bench.jsonl
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r4d_bench.jsonl').read())
print(d['value'], d['e2e']['value'], d['roofline']['achieved'])
print(json.dumps(d.get('secondary'), indent=0)[:2500])
PY
