#!/bin/bash
# multi-GPU correctness sweep: slab run == single-GPU run, bit for bit, for N ranks and both transports
N=${1:-2}
rc=0
for cfg in "300 210 6 rk2 periodic p2p" "131 96 5 rk4 periodic p2p" "120 80 6 rk2 fixed p2p" "90 70 4 euler reflect p2p" "300 210 6 rk2 periodic nccl" "1024 512 20 rk2 periodic p2p" "257 130 5 rk4 periodic p2p" "1000 300 12 rk4 periodic p2p"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_check.py $cfg 2>&1 | grep -E "mgpu_check|MISMATCH|Error|error" | head -8
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
