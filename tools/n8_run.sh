#!/bin/bash
# the 8-GPU validation call: every command under its own timeout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 500 python -m pytest tests/test_multi_gpu.py tests/test_mg_abi.py -m gpu -q -x) > gpurun_out/r2_pytest_n8.log 2>&1; tail -3 gpurun_out/r2_pytest_n8.log
for N in 8 4; do
  (timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
  echo "bench N=$N rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"],"floor",d["e2e"]["pcie_floor_ms"],d["e2e"]["ms_per_step"])
    print("   parity",{k:d["parity"][k] for k in ("rows_checked","rows_outside_bound","nan_rows","numpy_rows_outside_bound","pre_timing_rows_outside_bound","halo_poisoned","steps_run")})
    r=d["rmat24"]; print("   rmat",r.get("ms_per_step"),r.get("gflops"),r.get("parity"),r.get("all_gather_mode"),r.get("error"))
except Exception as e:
    print("N=$N parse failed",e)
P
  tail -c 300 gpurun_out/r2_bench_n$N.err
done
timeout 300 python tools/mg_bench.py --gpus 8 > gpurun_out/r2_mg_bench_n8.json 2> gpurun_out/r2_mg_bench_n8.err; echo "mg rc=$?"; cat gpurun_out/r2_mg_bench_n8.json | cut -c1-600; tail -c 300 gpurun_out/r2_mg_bench_n8.err
