#!/bin/bash
# One GPU call after the large-argument reduction changed:  tools/gpu_r2b.sh <prefix>   (writes gpurun_out/<prefix>_*)
# GPU tests + smoke, A/B of the previous library (exact Payne-Hanek on the hot path) against this one on the
# Payne-Hanek workloads and on plain data, both bench lines (C3 with the end-to-end legs, C4), ncu launch list and full
# captures (plain, C4), self-check, a short soak, all configs.
P=gpurun_out/$1
mkdir -p gpurun_out
OLD=fabe_b200/lib/ab/libfabe13_exact_ph.so
SHORT="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --sustain-seconds 0"
timeout 800 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
if [ -f $OLD ]; then
  timeout 300 python tools/ab_plain.py --c4 exact_ph=$OLD rolled=fabe_b200/lib/ab/libfabe13_rolled.so noel=fabe_b200/lib/ab/libfabe13_noel.so > ${P}_ab_c4.txt 2>&1; echo "ab_c4 rc=$?" | tee -a ${P}_rc.txt
fi
timeout 200 tools/abi_selfcheck > ${P}_selfcheck.txt 2>&1; echo "selfcheck rc=$?" | tee -a ${P}_rc.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "ref rc=$?" | tee -a ${P}_rc.txt
timeout 500 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --workload c4 --no-cpu-baseline --no-e2e > ${P}_bench_c4.json 2> ${P}_bench_c4.err; echo "bench c4 rc=$?" | tee -a ${P}_rc.txt
timeout 120 python bench.py $SHORT --elems 200000000 > ${P}_plain.log 2>&1 &&
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${P}_launches.csv \
      python bench.py $SHORT --elems 200000000 > ${P}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 2 -f -o ${P}_prof_fused \
      python bench.py $SHORT --elems 200000000 > ${P}_ncu_full.log 2>&1
timeout 120 python bench.py $SHORT --workload c4 > ${P}_plain_c4.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 1 -f -o ${P}_prof_fused_c4 \
      python bench.py $SHORT --workload c4 > ${P}_ncu_full_c4.log 2>&1
timeout 300 python tools/soak.py 80 11 > ${P}_soak.txt 2>&1; echo "soak rc=$?" | tee -a ${P}_rc.txt
if [ -f $OLD ]; then
  timeout 300 python tools/ab_plain.py exact_ph=$OLD > ${P}_ab_plain.txt 2>&1; echo "ab_plain rc=$?" | tee -a ${P}_rc.txt
fi
timeout 300 python tools/run_all_configs.py > ${P}_all_configs.md 2> ${P}_all_configs.err; echo "all_configs rc=$?" | tee -a ${P}_rc.txt
timeout 200 python tools/derived_timing.py > ${P}_derived.md 2> ${P}_derived.err
timeout 200 tools/abi_bench > ${P}_abi_bench.txt 2>&1; echo "abi_bench rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --workload c2 --no-cpu-baseline --no-e2e > ${P}_bench_c2.json 2> ${P}_bench_c2.err
timeout 200 python bench.py --core 1 --no-cpu-baseline --no-e2e > ${P}_bench_c3_psi.json 2> ${P}_bench_c3_psi.err
timeout 200 python tools/accuracy_report.py > ${P}_accuracy.txt 2>&1
cat ${P}_rc.txt; tail -3 ${P}_pytest.txt; tail -2 ${P}_soak.txt; tail -1 ${P}_selfcheck.txt
grep "c4 dense" ${P}_ab_c4.txt | tail -6
python - <<PY
import json
for f in ("${P}_bench_c3.json", "${P}_bench_c4.json"):
    try:
        j = json.load(open(f))
        r = j["roofline"]
        print(f, "value %.1f frac %.3f" % (j["value"], r["frac"]), "sustained", (r.get("sustained") or {}).get("value"))
        e = j.get("e2e")
        if e:
            print("  e2e %.3f frac %.3f | assist %s | pageable %.2f host-kept %s" % (e["value"], e["roofline"]["frac"], e.get("with_host_assist"),
                  j["e2e_pageable"]["value"], j["e2e_pageable"]["kept_on_the_host_instead"]))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
