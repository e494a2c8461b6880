#!/bin/bash
# Short 2-GPU call (gpurun --gpus 2):  tools/gpu_r2e.sh <prefix> -- the multi-GPU tests and bench.py under torchrun as the
# driver launches it, with the host-assisted leg of the end-to-end measurement on two ranks.
P=gpurun_out/$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_devices or all_gpus or multi_entry or host_assist" > ${P}_pytest_multi.txt 2>&1; echo "pytest rc=$?" | tee ${P}_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 > ${P}_bench_2gpu.json 2> ${P}_bench_2gpu.err; echo "bench N=2 rc=$?" | tee -a ${P}_rc.txt
cat ${P}_rc.txt; tail -3 ${P}_pytest_multi.txt; tail -5 ${P}_bench_2gpu.err
python - <<PY
import json
j = json.load(open("${P}_bench_2gpu.json"))
e, pg, r = j["e2e"], j["e2e_pageable"], j["roofline"]
print("N=2 value %.1f frac %.3f sustained %.1f traffic %s | e2e %.2f (probe %.2f, frac %.3f, %s) | assist %s | pageable %.2f host-kept %s" % (
    j["value"], r["frac"], r["sustained"]["value"], r["traffic"], e["value"], e["roofline"]["peak"], e["roofline"]["frac"], e["roofline"]["bound"],
    e.get("with_host_assist"), pg["value"], pg["kept_on_the_host_instead"]))
PY
