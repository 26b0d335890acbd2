#!/bin/bash
# 2-GPU session: weak scaling (headline workload per GPU) and BASELINE C5 as specified (10M total, d=32) strong scaling
mkdir -p gpurun_out
N=${1:-2}
run() { # name args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r02_multi_${name}_n$N.json 2> gpurun_out/r02_multi_${name}_n$N.err
  echo "$name rc=$?"; grep -c "nranks" gpurun_out/r02_multi_${name}_n$N.err; grep -m2 "nranks" gpurun_out/r02_multi_${name}_n$N.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_multi_${name}_n$N.json"))
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","synctest_consistent")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("sharded_parity"), d["timeline"].get("synchronous"))
except Exception as e:
    print("no json", e)
PY
  tail -c 400 gpurun_out/r02_multi_${name}_n$N.err
}
run weak --steps 300 --warmup 5
run c5strong --workload stress_10m_d32 --scaling strong --steps 40 --warmup 3
