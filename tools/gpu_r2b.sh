#!/bin/bash
# round 2, second GPU call (1 GPU): all GPU tests, the new bench line (sustained, e2e roofline, pageable leg), the
# reference arm on the full workload, single- vs two-output kernels alternating on the Payne-Hanek mix, latencies.
P=gpurun_out/$1
timeout 500 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "ref rc=$?" | tee -a ${P}_rc.txt
timeout 400 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
timeout 120 python tools/c4_single_vs_both.py > ${P}_c4_single_vs_both.md 2>&1; echo "c4 rc=$?" | tee -a ${P}_rc.txt
timeout 300 tools/abi_bench > ${P}_abi_bench.txt 2>&1; echo "abi_bench rc=$?" | tee -a ${P}_rc.txt
tail -5 ${P}_pytest.txt; cat ${P}_rc.txt; cat ${P}_c4_single_vs_both.md; tail -22 ${P}_abi_bench.txt; tail -c 1500 ${P}_bench_c3.err
