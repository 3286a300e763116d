set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2d_pytest.log 2>&1
grep -E "RRDBNet.forward|1024x1024 enhance|parse of 8|pipeline vs oracle|C5 \(|relative max|max \|err\||passed|failed|FAILED|Error" gpurun_out/r2d_pytest.log | head -40
timeout 300 profiles/bin/mma_peak 100000 > gpurun_out/r2d_mma_peak.jsonl 2>&1
cat gpurun_out/r2d_mma_peak.jsonl
timeout 900 python profiles/ref_on_b200.py --batch 16 --steps 3 > gpurun_out/r2d_ref_on_b200.jsonl 2> gpurun_out/r2d_ref_on_b200.err
cat gpurun_out/r2d_ref_on_b200.jsonl; tail -3 gpurun_out/r2d_ref_on_b200.err
timeout 600 python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2d_c5.log 2>&1
tail -c 1500 gpurun_out/r2d_c5.log
timeout 600 python bench.py --config c4 --steps 2 --warmup 1 > gpurun_out/r2d_c4.log 2>&1
tail -c 1200 gpurun_out/r2d_c4.log
