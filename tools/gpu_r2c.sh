#!/bin/bash
# round 2, third GPU call (1 GPU): tests on the new kernel routing, the straggler / launch-sequence sweep, the
# end-to-end pipeline against array and chunk size.
P=gpurun_out/$1
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 400 python tools/tune_stragglers.py > ${P}_tune_stragglers.md 2>&1; echo "tune rc=$?" | tee -a ${P}_rc.txt
timeout 400 python tools/e2e_sweep.py > ${P}_e2e_sweep.txt 2>&1; echo "e2e rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py --workload c4 --no-cpu-baseline --no-e2e > ${P}_bench_c4.json 2> ${P}_bench_c4.err; echo "c4 rc=$?" | tee -a ${P}_rc.txt
tail -5 ${P}_pytest.txt; cat ${P}_rc.txt; cat ${P}_tune_stragglers.md; cat ${P}_e2e_sweep.txt; cut -c1-300 ${P}_bench_c4.json
