cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --cpu-sample 0 > gpurun_out/r2b_bench_n$N.jsonl 2> gpurun_out/r2b_bench_n$N.err
tail -1 gpurun_out/r2b_bench_n$N.jsonl | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', d['ms_per_step'], 'ranks', d.get('ranks_ms_per_step'))
print('secondary', {k: round(v['value'],1) for k, v in (d.get('secondary') or {}).items()})
print('gather', d.get('gather'))
"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
