import sys, numpy as np
sys.path.insert(0,'.')
import heterogeneous_spmv_b200 as hs
from heterogeneous_spmv_b200 import generators as G
for name, make in (("cfg1", lambda: G.laplace2d_5pt(1000, rcm=True, mode="varied")), ("cfg2", lambda: G.stencil27(100, mode="varied")), ("cfg3", lambda: G.shell_blocks(1_500_000))):
    rp, ci, v = make()
    A = hs.Matrix.from_csr(rp, ci, v); A.build_sts(); A.sts_compute_b()
    A.set_option(hs.OPT_STS_BARRIERS, 1); t_bar = float(np.median(A.sts_solve_timed(2, 6)))
    A.set_option(hs.OPT_STS_BARRIERS, 2)
    out = []
    for g in (0, 512, 1024, 2048, 4096, 8192, 16384, 32768, 1 << 20):
        A.set_option(hs.OPT_STS_GROUPS, g)
        t = float(np.median(A.sts_solve_timed(2, 6)))
        ok = float(np.max(np.abs(A.sts_get_x() - 1.0))) == 0.0 and not A.sts_status()["timed_out"]
        out.append((g, round(t, 3), ok))
    print(name, "packs", len(A.sts()["pack_ptr"]) - 1, "barriers", round(t_bar, 3), "flags", out, flush=True)
    A.close()
