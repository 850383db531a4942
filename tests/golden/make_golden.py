"""Regenerates tests/golden/*.npz by running the REAL reference (oracle/_ref, compiled from the
unmodified /root/reference sources by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden.py

Each case stores the inputs' identity (generator + arguments, so the test regenerates them), the
reference's Band-k outputs (perm, sr_ptr[, ssr_ptr]), CRC-32s of the reordered row_ptr/col arrays
and of y's bit pattern, and y itself from the reference's OpenMP CSR-k SpMV and OpenMP CSR SpMV
(fp32 = the reference as written; fp64 = the same source with float widened, see oracle/Makefile).
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402

CASES = {
    # name: (generator, kwargs, k, sizes)
    "lap30_k3": ("laplace2d_5pt", {"g": 30}, 3, [7, 8]),
    "lap30_k2": ("laplace2d_5pt", {"g": 30}, 2, [7]),
    "lap64rcm_k2": ("laplace2d_5pt", {"g": 64, "rcm": True}, 2, [16]),
    "lap64rcm_k3": ("laplace2d_5pt", {"g": 64, "rcm": True, "mode": "varied"}, 3, [8, 8]),
    "st27_12_k2": ("stencil27", {"g": 12, "mode": "varied"}, 2, [32]),
    "st27_12_k3": ("stencil27", {"g": 12}, 3, [20, 10]),
    "shell4k_k3": ("banded_shell", {"n": 4000, "band": 120}, 3, [9, 12]),
    "shell4k_k2": ("banded_shell", {"n": 4000, "band": 120}, 2, [24]),
    "blocks3k_k2": ("shell_blocks", {"n": 3000, "node_band": 60}, 2, [16]),       # config-3 stand-in: 5-dof node blocks
    "blocks3k_k3": ("shell_blocks", {"n": 3000, "node_band": 60}, 3, [12, 6]),
}


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def main():
    assert O.have_ref(), "oracle/_ref is missing: make -C oracle ref"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])          # python make_golden.py [case ...]: regenerate just those
    for name, (gen, kw, k, sizes) in CASES.items():
        if only and name not in only:
            continue
        rec = {"k": np.int32(k), "sizes": np.array(sizes, np.int32)}
        for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
            rp, ci, v = getattr(G, gen)(dtype=dt, **kw)
            x = G.vector_x(len(rp) - 1, dt)
            R = O.RefCSRk(rp, ci, v, k, sizes)
            maps = R.maps()
            perm, r2, c2, v2 = R.matrix()
            y_k = R.spmv(x)                      # permuted row order
            y_csr = O.ref_csr_spmv(rp, ci, v, x)  # reference omp_spmv
            if tag == "f32":
                rec["crc_in_row_ptr"] = crc(rp)   # detects generator drift (numpy / scipy versions)
                rec["crc_in_col"] = crc(ci)
                rec["perm"] = perm
                rec["sr_ptr"] = maps[0]
                if k == 3:
                    rec["ssr_ptr"] = maps[1]
                rec["crc_row_ptr"] = crc(r2)
                rec["crc_col"] = crc(c2)
            else:
                assert np.array_equal(perm, rec["perm"])  # index work is type independent
            rec["crc_val_" + tag] = crc(v2)
            rec["y_csrk_" + tag] = y_k
            rec["y_csr_" + tag] = y_csr
            R.close()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "n=%d nSR=%d" % (len(rec["perm"]), len(rec["sr_ptr"]) - 1))


if __name__ == "__main__":
    main()
