#!/bin/bash
# N-GPU A/B of the bench step on ONE box: where the x update is enqueued, warps per CTA
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
for V in "--x-update main" "--x-update comm" "--x-update main --warps 16" "--x-update main"; do
  (timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 --no-rmat $V) > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$V" <<'P'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
    print("AB [%s] ms %.4f value %.1f parity_bad %d nan %d kernel %s" % (sys.argv[1], d["ms_per_step"], d["value"], d["parity"]["rows_outside_bound"], d["parity"]["nan_rows"], d["config"]["kernel"][40:80]))
except Exception as e:
    print("AB [%s] failed: %s" % (sys.argv[1], e))
P
done
