#!/bin/bash
# One GPU box, everything the round's evidence needs:  tools/gpu_round.sh <prefix>   (writes gpurun_out/<prefix>_*)
# GPU tests and smoke, bench lines (reference arm first, as the driver does), ncu launch list + one full capture of the
# top kernel on the plain and on the Payne-Hanek workload (each only after the same command ran clean without ncu),
# all configs, derived functions, latencies, self-check, soak, accuracy, the reference's own programs relinked.
P=gpurun_out/$1
SHORT="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --sustain-seconds 0"
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "ref rc=$?" | tee -a ${P}_rc.txt
timeout 400 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --workload c2 --no-cpu-baseline --no-e2e > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 200 python bench.py --workload c4 --no-cpu-baseline --no-e2e > ${P}_bench_c4.json 2> ${P}_bench_c4.err
timeout 200 python bench.py --core 1 --no-cpu-baseline --no-e2e > ${P}_bench_c3_psi.json 2> ${P}_bench_c3_psi.err
timeout 120 python bench.py $SHORT --elems 200000000 > ${P}_plain.log 2>&1 &&
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${P}_launches.csv \
      python bench.py $SHORT --elems 200000000 > ${P}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 2 -f -o ${P}_prof_fused \
      python bench.py $SHORT --elems 200000000 > ${P}_ncu_full.log 2>&1
timeout 120 python bench.py $SHORT --workload c4 > ${P}_plain_c4.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 1 -f -o ${P}_prof_fused_c4 \
      python bench.py $SHORT --workload c4 > ${P}_ncu_full_c4.log 2>&1
timeout 120 python tools/graph_latency.py > ${P}_graph_latency.md 2>&1; timeout 120 python tools/graph_latency.py 1048576 >> ${P}_graph_latency.md 2>&1
timeout 300 python tools/host_assist_sweep.py > ${P}_host_assist_sweep.md 2>&1
# A/B of kernel-variant builds, if any were shipped under fabe_b200/lib/ab/ (FABE13_BUILD_DEFINES=... python -m fabe_b200.build <out.so>)
AB=$(ls fabe_b200/lib/ab/*.so 2>/dev/null | sed 's#.*/libfabe13_\(.*\)\.so#\1=&#' | tr '\n' ' ')
[ -n "$AB" ] && timeout 300 python tools/ab_plain.py --c4 $AB > ${P}_ab_c4.txt 2>&1
timeout 300 python tools/run_all_configs.py > ${P}_all_configs.md 2> ${P}_all_configs.err; echo "all_configs rc=$?" | tee -a ${P}_rc.txt
timeout 200 python tools/derived_timing.py > ${P}_derived.md 2> ${P}_derived.err
timeout 200 tools/abi_bench > ${P}_abi_bench.txt 2>&1
timeout 200 tools/abi_selfcheck > ${P}_selfcheck.txt 2>&1; echo "selfcheck rc=$?" | tee -a ${P}_rc.txt
timeout 300 python tools/soak.py 200 11 > ${P}_soak.txt 2>&1; echo "soak rc=$?" | tee -a ${P}_rc.txt
timeout 200 python tools/accuracy_report.py > ${P}_accuracy.txt 2>&1
[ -x oracle/_ref/fabe13_test_runner_b200 ] && timeout 60 oracle/_ref/fabe13_test_runner_b200 > ${P}_ref_test_runner.txt 2>&1
[ -x oracle/_ref/fabe13_benchmark_b200 ] && timeout 400 stdbuf -oL oracle/_ref/fabe13_benchmark_b200 > ${P}_ref_benchmark_relinked.txt 2>&1
cat ${P}_rc.txt; tail -3 ${P}_pytest.txt; tail -2 ${P}_soak.txt; tail -1 ${P}_selfcheck.txt; tail -c 400 ${P}_bench_c3.json
