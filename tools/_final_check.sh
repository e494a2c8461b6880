P=gpurun_out/r2j
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" | tee -a ${P}_rc.txt
timeout 200 tools/pcie_probe --gpus 1 --secs 0.5 > ${P}_pcie_probe.txt 2>&1; echo "probe rc=$?" | tee -a ${P}_rc.txt
timeout 200 tools/abi_bench > ${P}_abi_bench.txt 2>&1; echo "abi rc=$?" | tee -a ${P}_rc.txt
tail -3 ${P}_pytest.txt; cat ${P}_rc.txt; cat ${P}_pcie_probe.txt; tail -21 ${P}_abi_bench.txt
[ -x oracle/_ref/fabe13_benchmark_b200 ] && FABE13_CUDA_MIN_N_PAGEABLE=2147483648 timeout 120 stdbuf -oL oracle/_ref/fabe13_benchmark_b200 > ${P}_ref_benchmark_relinked_hostcore.txt 2>&1
grep -E "^[0-9]" ${P}_ref_benchmark_relinked_hostcore.txt | cut -c1-120
