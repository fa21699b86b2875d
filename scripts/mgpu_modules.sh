#!/bin/bash
# slab decomposition with physics modules == single GPU, bit for bit (incl. sub-cycle counts)
N=${1:-2}
rc=0
for cfg in "128 96 4 rk2 periodic p2p tc,rl,ah" "128 96 3 rk2 periodic p2p tcsat" "96 80 3 euler reflect p2p tc,rl" "200 64 3 rk4 periodic p2p ah,tcsat,rl" "128 96 3 rk2 periodic p2p pv,tc" "120 90 3 rk4 periodic p2p av" "96 80 3 rk2 reflect p2p av,tc" "130 97 4 rk2 ucnp p2p 2feic" "131 96 3 rk4 periodic p2p 2f"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py $cfg 2>&1 | grep -E "mgpu_check|MISMATCH|subcycles|Error|error" | head -12
  [ ${PIPESTATUS[0]} -ne 0 ] && rc=1
done
exit $rc
