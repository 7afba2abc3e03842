run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e ${@:2} 2>/dev/null | tail -1; }
: > gpurun_out/sweep_n8.jsonl
run 29601 --robot iiwa14 --op rnea_grad --batch 131072 >> gpurun_out/sweep_n8.jsonl
run 29602 --robot iiwa14 --op rnea_grad --batch 2097152 >> gpurun_out/sweep_n8.jsonl
run 29603 --robot atlas --op rnea_grad --batch 32768 >> gpurun_out/sweep_n8.jsonl
run 29604 --robot atlas --op rnea_grad --batch 262144 >> gpurun_out/sweep_n8.jsonl
run 29605 --robot atlas --op minv --batch 262144 >> gpurun_out/sweep_n8.jsonl
run 29606 --robot iiwa14 --op minv --batch 1048576 >> gpurun_out/sweep_n8.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/sweep_n8.jsonl'):
    d=json.loads(l); c=d['config']
    print(c['robot'], c['op'], 'per-GPU', c['batch_per_gpu'], 'total', c['batch_per_gpu']*d['n_gpus'], '%.3e evals/s %.3f ms'%(d['value'], d['ms_per_step']))
PY
