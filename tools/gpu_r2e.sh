#!/bin/bash
# round 2, 1 GPU: kernel-variant A/B on the Payne-Hanek workloads and on plain data, the bench line (e2e against its probe)
P=gpurun_out/$1
V=libvariants
timeout 300 python tools/ab_plain.py --c4 r2c=$V/r2c.so u2=$V/u2.so u4=$V/u4.so mb5=$V/mb5.so mb4=$V/mb4.so mb4u2=$V/mb4u2.so t128=$V/t128.so > ${P}_ab_c4.txt 2>&1; echo "ab c4 rc=$?" | tee ${P}_rc.txt
timeout 300 python tools/ab_plain.py r2c=$V/r2c.so u2=$V/u2.so mb5=$V/mb5.so t128=$V/t128.so > ${P}_ab_plain.txt 2>&1; echo "ab plain rc=$?" | tee -a ${P}_rc.txt
timeout 400 python bench.py --no-cpu-baseline > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; grep "rep [12]" ${P}_ab_c4.txt; grep "rep [12]" ${P}_ab_plain.txt
python - <<PY
import json
j = json.load(open("${P}_bench_c3.json"))
e, pg, r = j["e2e"], j["e2e_pageable"], j["roofline"]
print("value %.1f frac %.3f sustained %.1f | e2e %.3f (probe %.3f, frac %.3f, %s) | pageable %.2f" % (
    j["value"], r["frac"], r["sustained"]["value"], e["value"], e["roofline"]["peak"], e["roofline"]["frac"], e["roofline"]["bound"], pg["value"]))
PY
