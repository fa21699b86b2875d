#!/bin/bash
# per-step timeline of rank 0 (CUDA events inside the library) at N ranks: bench.py's timed run with SPRUCE_TIMELINE
set -u
n=${1:-8}
SPRUCE_TIMELINE=3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/tl_bench$n.json 2> gpurun_out/tl_bench$n.err
grep "timeline rank 0\]" gpurun_out/tl_bench$n.err | tail -48 > gpurun_out/tl_rank0_n$n.txt
grep "timeline rank 3\]" gpurun_out/tl_bench$n.err | tail -48 > gpurun_out/tl_rank3_n$n.txt
cat gpurun_out/tl_rank0_n$n.txt
