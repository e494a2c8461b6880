#!/bin/bash
# Short GPU call: is the register-resident dense route + evict-last table load right?  tools/gpu_r2c.sh <prefix>
P=gpurun_out/$1
mkdir -p gpurun_out
AB=fabe_b200/lib/ab
LIB=fabe_b200/lib/libfabe13.so
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 300 python tools/ab_plain.py --c4 exact_ph=$AB/libfabe13_exact_ph.so rolled=$AB/libfabe13_rolled.so noel=$AB/libfabe13_noel.so \
    dm8=$LIB:direct_max=8 dm128=$LIB:direct_max=128 > ${P}_ab_c4.txt 2>&1; echo "ab_c4 rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --workload c4 --no-cpu-baseline --no-e2e > ${P}_bench_c4.json 2> ${P}_bench_c4.err; echo "bench c4 rc=$?" | tee -a ${P}_rc.txt
SHORT="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --sustain-seconds 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sincos_fused -s 3 -c 1 -f -o ${P}_prof_fused_c4 \
      python bench.py $SHORT --workload c4 > ${P}_ncu_full_c4.log 2>&1
timeout 200 python tools/tune_stragglers.py > ${P}_tune_stragglers.md 2>&1; echo "tune rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; tail -3 ${P}_pytest.txt; cat ${P}_ab_c4.txt | sort -k4,4 -k5,5 -s | awk '{print}' | tail -60
python - <<PY
import json
j = json.load(open("${P}_bench_c4.json")); r = j["roofline"]
print("c4 value %.1f frac %.3f sustained %s" % (j["value"], r["frac"], (r.get("sustained") or {}).get("value")))
PY
