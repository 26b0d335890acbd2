#!/bin/bash
# weak-scaling line at N GPUs (run under `gpurun --gpus N`): bash scripts/gpu_scaling.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 1000 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 300 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][0])
print(d['n_gpus'], 'value=%.0f ms=%.4f e2e=%.0f ok=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['synctest_consistent']))
PY
