"""PCIe floor of the end-to-end call: 512 MB up, 512 MB down, alone and together (pinned memory)."""
import torch
n = 64_000_000
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def up():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    up()
    down()


def both_chunked(c=32):
    k = n // c
    for i in range(c):
        with torch.cuda.stream(s1):
            d_in[i * k:(i + 1) * k].copy_(h_in[i * k:(i + 1) * k], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i * k:(i + 1) * k].copy_(d_out[i * k:(i + 1) * k], non_blocking=True)


for name, fn in (("H2D 512 MB", up), ("D2H 512 MB", down), ("both directions", both), ("both, 32 chunks each", both_chunked)):
    for s in (s1, s2):
        s.wait_stream(torch.cuda.current_stream())
    ms = timed(fn)
    print("%-24s %.3f ms  (%.1f GB/s per direction)" % (name, ms, 0.512 / ms * 1e3), flush=True)
