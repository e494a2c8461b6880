#!/bin/bash
# One short GPU call that verifies a changed library end to end:  tools/gpu_verify.sh <prefix>
# GPU tests, smoke, the default bench line, small-n placement, call latencies.  Outputs in gpurun_out/<prefix>_*.
P=gpurun_out/$1
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
timeout 400 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
timeout 200 python tools/sweep_small_n.py > ${P}_sweep_small_n.md 2>&1; echo "small_n rc=$?" | tee -a ${P}_rc.txt
timeout 200 tools/abi_bench > ${P}_abi_bench.txt 2>&1; echo "abi_bench rc=$?" | tee -a ${P}_rc.txt
tail -3 ${P}_pytest.txt; cat ${P}_rc.txt; cat ${P}_sweep_small_n.md; tail -22 ${P}_abi_bench.txt
python - <<PY
import json
j = json.load(open("${P}_bench_c3.json"))
e, pg, r = j["e2e"], j["e2e_pageable"], j["roofline"]
print("value %.1f frac %.3f sustained %.1f traffic %s | e2e %.3f (probe %.3f, frac %.3f, %s) | pageable %.2f host-core %s" % (
    j["value"], r["frac"], r["sustained"]["value"], r["traffic"], e["value"], e["roofline"]["peak"], e["roofline"]["frac"], e["roofline"]["bound"],
    pg["value"], pg["kept_on_the_host_instead"]))
PY
