#!/bin/bash
# quick check of a stage-kernel change: smoke, the stage-kernel parity tests, bench lines (exact / relaxed, 2-D / full instance)
set -u
mkdir -p gpurun_out
tag=${1:-q}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "golden or orszag or stratified or edge or 2d_instance or batched" > gpurun_out/${tag}_parity.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/${tag}_parity.log
timeout 900 python -m pytest tests/test_gpu_extended.py -x -q -m gpu -p no:cacheprovider -k "stage_kernel or relaxed" > gpurun_out/${tag}_ext.log 2>&1; echo "ext rc=$?"; tail -4 gpurun_out/${tag}_ext.log
run() { name=$1; shift
    timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline "$@" > "gpurun_out/${name}.json" 2> "gpurun_out/${name}.err"
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.3e  frac %.3f  e2e %.3e  ms/step %.3f  stage1 %.3f ms" % (l["value"], l["roofline"]["frac"], l["e2e"]["value"], l["ms_per_step"], l["roofline"]["stage1_alone_ms"]))
except Exception as e:
    print(n, "FAILED", e)
PY
}
run ${tag}_exact
SPRUCE_VEC_ROWS=0 run ${tag}_exact_novec
run ${tag}_relaxed --arith relaxed
run ${tag}_zfull_exact --zfull
run ${tag}_zfull_relaxed --zfull --arith relaxed
