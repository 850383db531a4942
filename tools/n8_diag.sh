# usage: bash tools/n8_diag.sh N   -- strong-scaling diagnosis runs (peer exchange, then compute only)
N=${1:-8}
run() { tag=$1; shift; HSPMV_SPIN_STATS=1 timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 "$@" > gpurun_out/n${N}_$tag.log 2>&1; echo "== $tag"; grep "halo spin\|own ms" gpurun_out/n${N}_$tag.log | sort; grep "^{" gpurun_out/n${N}_$tag.log > gpurun_out/n${N}_$tag.json; python -c "import sys,json; d=json.loads(open(sys.argv[1]).read()); print(d[\"config\"][\"exchange\"], d[\"ms_per_step\"], d[\"value\"], d[\"host_enqueue_ms_per_step\"], d[\"e2e\"][\"value\"])" gpurun_out/n${N}_$tag.json || grep -i "error" gpurun_out/n${N}_$tag.log | head -5; }
run peer --exchange peer
run noex --no-exchange
