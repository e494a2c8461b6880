#!/bin/bash
# round 2, first GPU call (1 GPU): tests of the new host routes, the host-link probe, call latencies, the
# single-output vs two-output question on the Payne-Hanek mix, and the default bench line.
P=gpurun_out/$1
timeout 400 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
timeout 200 tools/pcie_probe --gpus 1 --secs 1.0 > ${P}_pcie_probe.txt 2>&1; echo "probe rc=$?" | tee -a ${P}_rc.txt
timeout 200 tools/abi_bench > ${P}_abi_bench.txt 2>&1; echo "abi_bench rc=$?" | tee -a ${P}_rc.txt
timeout 200 python tools/derived_timing.py > ${P}_derived.md 2> ${P}_derived.err; echo "derived rc=$?" | tee -a ${P}_rc.txt
timeout 200 python bench.py > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?" | tee -a ${P}_rc.txt
nvidia-smi -q | grep -iE "numa|link|pcie|gen|width" | head -30 > ${P}_nvsmi.txt 2>&1
lscpu > ${P}_lscpu.txt 2>&1; numactl -H >> ${P}_lscpu.txt 2>&1; free -g >> ${P}_lscpu.txt
tail -5 ${P}_pytest.txt; cat ${P}_rc.txt; cat ${P}_pcie_probe.txt; tail -22 ${P}_abi_bench.txt
