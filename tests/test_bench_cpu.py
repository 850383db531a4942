"""bench.py pieces that run without a GPU: the reference arm (the compiled reference's omp_spmv on the host cores, arrays from
the oracle's C generator -- it must not map the GPU library), the closed forms and the bookkeeping of the parity record."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_it_does_not_load_the_gpu_library(tmp_path):
    code = ("import sys, json, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--g', '24', '--steps', '2', '--warmup', '1'];"
            "runpy.run_path(%r, run_name='__main__');"
            "print('MAPS_HSPMV', any('libhspmv' in l for l in open('/proc/self/maps')))" % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    d = json.loads([l for l in lines if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "stencil27_g24_csr" and d["config"]["nnz"] == (3 * 24 - 2) ** 3
    assert d["config"]["rows_timed"] == 24 ** 3                       # the WHOLE matrix
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "MAPS_HSPMV False" in p.stdout, "the reference arm mapped libhspmv.so"


def test_reference_arm_under_torchrun_only_rank0_prints():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--g", "16", "--steps", "1", "--warmup", "1"]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    js = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(js) == 1 and json.loads(js[0])["n_gpus"] == 2


def test_bench_bookkeeping():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.merge_segments([(5, 9), (0, 3), (3, 4), (8, 12), (20, 21)]) == [(0, 4), (5, 12), (20, 21)]
    assert [bench.step_scale(t) for t in (0, 1, 7, 8, 9)] == [1.0, 1.125, 1.875, 1.0, 1.125]
    # scales are exact in fp32 times a vector_x entry (k/128): no rounding enters the expected x
    x = (1.0 + np.arange(97) / 128.0)
    for t in range(8):
        assert np.array_equal((x * bench.step_scale(t)).astype(np.float32).astype(np.float64), x * bench.step_scale(t))
    g = 12
    assert bench.stencil27_before(g ** 3, g) == (3 * g - 2) ** 3
    assert bench.alg_bytes(10, 10, 100) == 100 * 12 + 11 * 4 + 80 + 80
    rp, ci, v = (np.array([0, 2, 2, 3]), np.array([0, 2, 1]), np.array([2.0, -1.0, 4.0]))
    y, s = bench.numpy_spmv(rp, ci, v, np.array([1.0, 2.0, 3.0]))
    assert y.tolist() == [-1.0, 0.0, 8.0] and s.tolist() == [5.0, 0.0, 8.0]
