#!/bin/bash
# First GPU call of the next round: everything that was written after the round-1 GPU budget was spent, in one go.
#   gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Results land in gpurun_out/: the isolated tests (XPASS = validated -> drop the xfail marker / the switch), then the bench line in the
# eight arithmetic / instance combinations (same workload, same box, back to back).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2_build.log 2>&1
timeout 1200 python -m pytest tests/test_zz_gpu_unvalidated.py -m gpu -q -rxXs -p no:cacheprovider > gpurun_out/r2_unvalidated.log 2>&1
tail -40 gpurun_out/r2_unvalidated.log
for mode in "exact:0" "exact:1" "exact:2" "exact:3" "relaxed:0" "relaxed:1" "relaxed:2" "relaxed:3"; do
    arith=${mode%%:*}; sv=${mode#*:}
    name="r2_bench_${arith}_variants${sv}"
    timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --arith "$arith" --stage-variants "$sv" > "gpurun_out/${name}.json" 2> "gpurun_out/${name}.err"
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    l = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.3e  frac %.3f  e2e %.3e  ms/step %.3f" % (l["value"], l["roofline"]["frac"], l["e2e"]["value"], l["ms_per_step"]))
except Exception as e:
    print(n, "FAILED", e)
PY
done
# host text I/O on the box's host cores: thread scaling of the row-parallel formatter / parser
g++ -O2 -fopenmp -std=c++17 -I spruce_b200/host scripts/text_io_bench.cpp -o /tmp/text_io_bench 2> gpurun_out/r2_text_io.err && for t in 1 8 32; do
    OMP_NUM_THREADS=$t /tmp/text_io_bench 4096 4 | sed "s/^/threads $t: /" | tee -a gpurun_out/r2_text_io.jsonl
done
