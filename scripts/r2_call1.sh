#!/bin/bash
# Round 2, call 1: timing of the stage-kernel configurations that round 1 wrote but never timed (exact / relaxed x SPRUCE_STAGE_VARIANTS 0..3),
# the 12-quantity instance, and host text I/O thread scaling.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt
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
for mode in "exact:0" "exact:1" "relaxed:1"; do
    arith=${mode%%:*}; sv=${mode#*:}
    name="r2_bench_zfull_${arith}_variants${sv}"
    timeout 300 python bench.py --zfull --steps 30 --warmup 5 --no-cpu-baseline --arith "$arith" --stage-variants "$sv" > "gpurun_out/${name}.json" 2> "gpurun_out/${name}.err"
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
g++ -O2 -fopenmp -std=c++17 -I spruce_b200/host scripts/text_io_bench.cpp -o /tmp/text_io_bench 2> gpurun_out/r2_text_io.err && for t in 1 8 32; do
    OMP_NUM_THREADS=$t /tmp/text_io_bench 4096 4 | sed "s/^/threads $t: /" | tee -a gpurun_out/r2_text_io.jsonl
done
