N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 300 gpurun_out/r2_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload mappo --steps 2 --warmup 1 > gpurun_out/r2_mappo_${N}gpu.json 2> gpurun_out/r2_mappo_${N}gpu.err
tail -c 300 gpurun_out/r2_mappo_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu'), d['on_device_pipeline']['value'], d['learner_allreduce'])
m=json.loads(open('gpurun_out/r2_mappo_${N}gpu.json').read().strip().splitlines()[-1])
print(m['value'], m['ms_per_step'], m['config'])
"
