#!/bin/bash
set -u
tag=${1:-b}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log | cut -c1-80
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
run ${tag}_relaxed --arith relaxed
run ${tag}_zfull_exact --zfull
