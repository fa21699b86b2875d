#!/bin/bash
# round-end measurement on one B200: full GPU test suite, host-binary drop-in tests, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_mhd_stage_xy -s 6 -c 2 -f -o gpurun_out/prof_stage_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_full.log 2>&1
cat gpurun_out/final_pytest.log gpurun_out/final_smoke.log; cut -c1-400 gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench_ref.json
