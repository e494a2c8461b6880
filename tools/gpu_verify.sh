#!/bin/bash
# One short GPU call that verifies a changed library end to end:  tools/gpu_verify.sh <prefix> [previous-library.so]
# GPU tests, smoke, the default bench line, config 4 with the new and (if given) the previous library, all configs,
# then the ncu capture of the fused kernel on config 4.  Every step has its own timeout; outputs in gpurun_out/<prefix>_*.
# QUICK=1: tests, smoke, config 4 (new, previous), then the default workload without its CPU baseline -- about 2.5 minutes.
P=gpurun_out/$1
OLD=$2
SHORT="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
[ -z "$QUICK" ] && { timeout 120 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt; }
timeout 60 python bench.py --workload c4 --no-cpu-baseline > ${P}_bench_c4.json 2> ${P}_bench_c4.err
[ -n "$OLD" ] && FABE13_LIBRARY=$OLD timeout 60 python bench.py --workload c4 --no-cpu-baseline --no-e2e > ${P}_bench_c4_previous.json 2> ${P}_bench_c4_previous.err
if [ -n "$QUICK" ]; then
  timeout 60 python bench.py --no-cpu-baseline > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
  tail -3 ${P}_pytest.txt; cat ${P}_bench_c4.json ${P}_bench_c4_previous.json ${P}_bench_c3.json 2>/dev/null | cut -c1-120
  exit 0
fi
timeout 90 python tools/run_all_configs.py > ${P}_all_configs.md 2> ${P}_all_configs.err
timeout 60 python bench.py $SHORT --workload c4 > ${P}_plain_c4.log 2>&1 &&
  timeout 90 ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 1 -f -o ${P}_prof_fused_c4 \
      python bench.py $SHORT --workload c4 > ${P}_ncu_full_c4.log 2>&1
tail -3 ${P}_pytest.txt; cat ${P}_rc.txt; cat ${P}_bench_c4.json ${P}_bench_c4_previous.json 2>/dev/null | cut -c1-200
