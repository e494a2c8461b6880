#!/bin/bash
# 8-GPU call (gpurun --gpus 8):  tools/gpu_r2f.sh <prefix> -- bench.py under torchrun at 8 and 4 ranks as the driver launches it.
P=gpurun_out/$1
mkdir -p gpurun_out
PORT=29541
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 > ${P}_bench_${N}gpu.json 2> ${P}_bench_${N}gpu.err; echo "bench N=$N rc=$?" | tee -a ${P}_rc.txt
  PORT=$((PORT+1))
  python - <<PY
import json
try:
    j = json.load(open("${P}_bench_${N}gpu.json"))
    e, pg, r = j["e2e"], j["e2e_pageable"], j["roofline"]
    a = e.get("with_host_assist") or {}
    print("N=$N value %.1f frac %.3f sustained %.1f | e2e %.2f (probe %.2f, frac %.3f, %s) | assist %s (gpu share %s, %s helpers) | pageable %.2f host-kept %s" % (
        j["value"], r["frac"], r["sustained"]["value"], e["value"], e["roofline"]["peak"], e["roofline"]["frac"], e["roofline"]["bound"],
        a.get("value"), a.get("gpu_share"), a.get("host_threads_per_rank"), pg["value"], (pg["kept_on_the_host_instead"] or {}).get("value")))
except Exception as ex:
    print("N=$N:", repr(ex))
PY
done
tail -3 ${P}_bench_8gpu.err
