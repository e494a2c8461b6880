#!/bin/bash
# One GPU box, everything the round's evidence needs:  tools/gpu_round.sh <prefix>   (writes gpurun_out/<prefix>_*)
# bench lines (reference arm first, as the driver does), ncu launch list + one full capture of the top kernel on
# the plain and on the Payne-Hanek workload (each only after the same command ran clean without ncu), all configs.
P=gpurun_out/$1
SHORT="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err
python bench.py --workload c2 --no-cpu-baseline > ${P}_bench_c2.json 2> ${P}_bench_c2.err
python bench.py --workload c4 --no-cpu-baseline > ${P}_bench_c4.json 2> ${P}_bench_c4.err
python bench.py $SHORT --elems 200000000 > ${P}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${P}_launches.csv \
      python bench.py $SHORT --elems 200000000 > ${P}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 2 -o ${P}_prof_fused \
      python bench.py $SHORT --elems 200000000 > ${P}_ncu_full.log 2>&1
python bench.py $SHORT --workload c4 > ${P}_plain_c4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 1 -o ${P}_prof_fused_c4 \
      python bench.py $SHORT --workload c4 > ${P}_ncu_full_c4.log 2>&1
python tools/run_all_configs.py > ${P}_all_configs.md 2> ${P}_all_configs.err
python tools/derived_timing.py > ${P}_derived.md 2> ${P}_derived.err
python tools/ab_worklist.py final 0,1 > ${P}_ab_worklist.txt 2>&1
tools/abi_bench > ${P}_abi_bench.txt 2>&1
tools/abi_selfcheck > ${P}_selfcheck.txt 2>&1
python tools/soak.py 300 7 > ${P}_soak.txt 2>&1
python tools/accuracy_report.py > ${P}_accuracy.txt 2>&1
python tools/tune_thresholds.py > ${P}_tune_thresholds.md 2>&1
python tools/exhaustive_f32.py --device --threads $(nproc) > ${P}_sincosf_exhaustive.txt 2>&1
[ -x oracle/_ref/fabe13_test_runner_b200 ] && oracle/_ref/fabe13_test_runner_b200 > ${P}_ref_test_runner.txt 2>&1
tail -c 600 ${P}_bench_c3.json
