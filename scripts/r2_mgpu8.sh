#!/bin/bash
# 8-GPU call: bench lines at N = 8, 4, 2 (and the hashes to compare with N = 1)
# LESSON of round 2: a multi-GPU call is charged N x its wall time -- keep every command under a SHORT timeout (a hung 8-GPU bench with timeout 600 x 3 cost 108 GPU-minutes)
set -u
mkdir -p gpurun_out
tag=${1:-mg8}
for n in ${2:-8 4}; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/${tag}_bench$n.json 2> gpurun_out/${tag}_bench$n.err; echo "bench$n rc=$?"
  python - $n $tag <<'PY'
import json, sys
n, tag = sys.argv[1], sys.argv[2]
try:
    l = json.loads(open("gpurun_out/%s_bench%s.json" % (tag, n)).read().strip().splitlines()[-1])
    print("N=%s value %.3e ms/step %.4f e2e %.3e parity %s dt_hash %s state_hash %s launches %d" % (n, l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("parity_vs_1gpu"), l.get("dt_hash"), l.get("state_hash"), l["gpu_launches"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/%s_bench%s.err" % (tag, n)).read()[-1500:])
PY
done
