#!/bin/bash
# Short 1-GPU call:  tools/gpu_r2d.sh <prefix> -- which kernel should serve small arrays (graph replay, small_n), the
# host_assist sweep, and a final bench line (with the stamped ncu traffic) + GPU tests on the library as committed.
P=gpurun_out/$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 120 python tools/graph_latency.py > ${P}_graph_latency.md 2>&1; echo "graph rc=$?" | tee -a ${P}_rc.txt
timeout 120 python tools/graph_latency.py 0 >> ${P}_graph_latency.md 2>&1
timeout 120 python tools/graph_latency.py 65536 >> ${P}_graph_latency.md 2>&1
timeout 300 python tools/host_assist_sweep.py > ${P}_host_assist_sweep.md 2>&1; echo "assist rc=$?" | tee -a ${P}_rc.txt
timeout 500 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; tail -2 ${P}_pytest.txt; cat ${P}_graph_latency.md ${P}_host_assist_sweep.md; cut -c1-600 ${P}_bench_c3.json
