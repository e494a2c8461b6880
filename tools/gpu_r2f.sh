#!/bin/bash
# round 2, 1 GPU: block size of the streaming kernels (64 / 128 / 256 threads) on plain data, slices, small n and mixes
P=gpurun_out/$1
V=libvariants
timeout 400 python tools/ab_plain.py t64=$V/t64.so t128=$V/t128.so t256=$V/t256.so > ${P}_ab_threads_plain.txt 2>&1; echo "plain rc=$?" | tee ${P}_rc.txt
timeout 300 python tools/ab_plain.py --c4 t64=$V/t64.so t128=$V/t128.so t256=$V/t256.so > ${P}_ab_threads_c4.txt 2>&1; echo "c4 rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; cat ${P}_ab_threads_plain.txt ${P}_ab_threads_c4.txt
