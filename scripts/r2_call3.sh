#!/bin/bash
# stage kernel v13 (bulk row copies, own-cell staging in shared memory, register-kept transports): parity first, then timing
set -u
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c3_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2c3_parity.log 2>&1; echo "parity rc=$?"; tail -15 gpurun_out/r2c3_parity.log
timeout 900 python -m pytest tests/test_gpu_extended.py -x -q -m gpu -p no:cacheprovider -k "stage_kernel or relaxed" > gpurun_out/r2c3_ext.log 2>&1; echo "ext rc=$?"; tail -8 gpurun_out/r2c3_ext.log
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
run r2c3_exact
SPRUCE_VEC_ROWS=0 run r2c3_exact_nobulk
run r2c3_relaxed --arith relaxed
run r2c3_zfull_exact --zfull
run r2c3_zfull_relaxed --zfull --arith relaxed
