P=gpurun_out/r2q
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
timeout 200 python tools/tune_dense_hint.py > ${P}_tune_dense_hint.md 2>&1; echo "dense rc=$?" | tee -a ${P}_rc.txt
timeout 500 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --elems 125000000 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > ${P}_bench_slice.json 2> ${P}_bench_slice.err; echo "slice rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --workload c4 --no-cpu-baseline --no-e2e > ${P}_bench_c4.json 2> ${P}_bench_c4.err; echo "c4 rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; tail -2 ${P}_pytest.txt; cat ${P}_smoke.txt | tail -1; cat ${P}_tune_dense_hint.md
python - <<PY
import json
for f in ("c3","slice","c4"):
    j = json.load(open("${P}_bench_%s.json" % f)); r = j["roofline"]
    print(f, "value %.1f frac %.3f" % (j["value"], r["frac"]), "step avg/median/min", r["step_ms_avg"], r["step_ms_median"], r["step_ms_min"], "sustained", (r.get("sustained") or {}).get("value"), "traffic", r["traffic"])
PY
