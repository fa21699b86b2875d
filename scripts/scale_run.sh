#!/bin/bash
# strong-scaling sweep on one box: bench.py at N = 1, 2, 4, 8 (N>1 through torchrun), one JSON line each into gpurun_out/scale_<tag>.jsonl
TAG=${1:-r1}; SIZE=${2:-4096}; STEPS=${3:-50}; NS=${4:-"1 2 4 8"}
mkdir -p gpurun_out
OUT=gpurun_out/scale_${TAG}_${SIZE}.jsonl
: > $OUT
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 5 --size $SIZE --no-cpu-baseline >> $OUT 2> gpurun_out/scale_${TAG}_${SIZE}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps $STEPS --warmup 5 --size $SIZE >> $OUT 2> gpurun_out/scale_${TAG}_${SIZE}_n$N.err
  fi
  echo "N=$N rc=$?"
done
python - <<PY
import json
for ln in open("$OUT"):
    ln=ln.strip()
    if ln.startswith("{"):
        d=json.loads(ln); print(d["n_gpus"], "%.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"])
PY
