#!/bin/bash
# 2-GPU call: slab == 1-GPU parity tests, then the bench line at N=2 (with the parity check and hashes inside)
set -u
mkdir -p gpurun_out
sel=${1:-slab}
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "$sel" > gpurun_out/mg2_parity.log 2>&1; echo "slab parity rc=$?"; tail -6 gpurun_out/mg2_parity.log
if [ "${2:-}" = "ext" ]; then timeout 900 python -m pytest tests/test_gpu_extended.py -x -q -m gpu -p no:cacheprovider -k "slab" > gpurun_out/mg2_ext.log 2>&1; echo "slab ext rc=$?"; tail -6 gpurun_out/mg2_ext.log; fi
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/mg2_bench2.json 2> gpurun_out/mg2_bench2.err; echo "bench2 rc=$?"; grep -v Warning gpurun_out/mg2_bench2.err | tail -5
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/mg2_bench2.json").read().strip().splitlines()[-1])
    print("N=2 value %.3e ms/step %.3f e2e %.3e parity %s dt_hash %s state_hash %s launches %d" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("parity_vs_1gpu"), l.get("dt_hash"), l.get("state_hash"), l["gpu_launches"]))
except Exception as e:
    print("FAILED", e)
PY
