#!/bin/bash
# Multi-GPU call:  tools/gpu_multi.sh <prefix> <counts, e.g. "2" or "8 4 2">   (run under gpurun --gpus N, N = the largest count)
# The host-link probe at every count, the multi-GPU tests, bench.py under torchrun (as the driver launches it) at each
# count, the reference arm once.
P=gpurun_out/$1
COUNTS=${2:-2}
MAXN=$(echo $COUNTS | tr ' ' '\n' | sort -n | tail -1)
LIST=$(echo 1 $COUNTS | tr ' ' '\n' | sort -n | uniq | paste -sd,)
nvidia-smi topo -m > ${P}_topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA" > ${P}_host.txt; free -g >> ${P}_host.txt
timeout 600 tools/pcie_probe --gpus $LIST --secs 0.6 > ${P}_pcie_probe.txt 2>&1; echo "probe rc=$?" | tee ${P}_rc.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_devices or all_gpus or multi_entry" > ${P}_pytest_multi.txt 2>&1; echo "pytest rc=$?" | tee -a ${P}_rc.txt
PORT=29511
for N in $COUNTS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 > ${P}_bench_${N}gpu.json 2> ${P}_bench_${N}gpu.err; echo "bench N=$N rc=$?" | tee -a ${P}_rc.txt
  PORT=$((PORT+1))
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $MAXN --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --impl reference --gpus $MAXN --steps 5 --warmup 2 > ${P}_bench_ref_${MAXN}gpu.json 2> ${P}_bench_ref_${MAXN}gpu.err; echo "ref rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; cat ${P}_host.txt; cat ${P}_pcie_probe.txt; tail -3 ${P}_pytest_multi.txt
for N in $COUNTS; do python - <<PY
import json
try:
    j = json.load(open("${P}_bench_${N}gpu.json"))
    e, pg, r = j["e2e"], j["e2e_pageable"], j["roofline"]
    print("N=$N value %.1f frac %.3f sustained %.1f | e2e %.2f (probe %.2f, frac %.3f, %s) | pageable %.2f" % (
        j["value"], r["frac"], r["sustained"]["value"], e["value"], e["roofline"]["peak"], e["roofline"]["frac"], e["roofline"]["bound"], pg["value"]))
except Exception as ex:
    print("N=$N:", repr(ex))
PY
done
cut -c1-400 ${P}_bench_ref_${MAXN}gpu.json
