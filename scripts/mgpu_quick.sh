#!/bin/bash
# quick 2-rank spot checks (subset of scripts/mgpu_modules.sh)
for cfg in "${@:-128 96 3 rk2 periodic p2p pv,tc}"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py $cfg 2>&1 | grep -E "mgpu_check|MISMATCH|subcycles|SpruceError" | head -8
done
