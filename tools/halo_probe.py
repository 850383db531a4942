"""Single-GPU cost of the fused halo launch vs the plain one (no exchange involved)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs

g = int(sys.argv[1]) if len(sys.argv) > 1 else 252
A = hs.Matrix.stencil(27, g, np.float64, 0)
A.fill_x_synthetic(0)
m = A.info()["m"]
half = g * g + g + 1
st = torch.cuda.current_stream().cuda_stream
flags = hs.peer_alloc(256)
hs.stream_write(st, flags, 1 << 30)
hs.stream_write(st, flags + 4, 1 << 30)
torch.cuda.synchronize()

def timeit(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("plain range      %.4f ms" % timeit(lambda: A.spmv_range(0, m, hs.ALGO_CSR_STREAM, stream=st)))
print("halo n_wait=0    %.4f ms" % timeit(lambda: A.spmv_halo(half, m - half, stream=st)))
print("halo n_wait=2    %.4f ms" % timeit(lambda: A.spmv_halo(half, m - half, wait_flags=flags, n_wait=2, wait_step=5, timed_out=flags + 128, stream=st)))
print("3 launches       %.4f ms" % timeit(lambda: (A.spmv_range(half, m - half, hs.ALGO_CSR_STREAM, stream=st), A.spmv_range(0, half, hs.ALGO_CSR_STREAM, stream=st), A.spmv_range(m - half, m, hs.ALGO_CSR_STREAM, stream=st))))
step = [0]
def with_memops():
    step[0] += 1
    hs.stream_write(st, flags + 8, step[0])
    hs.stream_wait_ge(st, flags + 8, step[0])
    A.spmv_range(0, m, hs.ALGO_CSR_STREAM, stream=st)
print("plain + write/wait memops on the same stream %.4f ms" % timeit(with_memops))
side = torch.cuda.Stream()
def with_side_memops():
    step[0] += 1
    hs.stream_write(side.cuda_stream, flags + 12, step[0])
    hs.memcpy_async(A.y_ptr(), A.x_ptr(), 1280000, side.cuda_stream)
    A.spmv_range(0, m, hs.ALGO_CSR_STREAM, stream=st)
print("plain + write + 1.28MB copy on a side stream  %.4f ms" % timeit(with_side_memops))
