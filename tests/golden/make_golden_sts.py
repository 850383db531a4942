"""Regenerates tests/golden/sts/*.npz by running the REAL reference's STS path (kernel "STS", order "LS", k = 2:
putInCSRkFormat, incomplete_choloskey, compute_b, lowerSTS) through oracle/_ref.  Run in the build container:

    python tests/golden/make_golden_sts.py

Each case stores the reference's level-set permutation and pack bounds, CRC-32s of L (row pointers, columns, values) and
of b = L * 1, and the solution x for a right-hand side that is not L * 1 (b_i = 1 + (i mod 13) / 16), fp32 (the
reference as written) and fp64 (the same source with float widened, see oracle/Makefile)."""
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
    # name: (generator, kwargs)
    "sts_st27_9": ("stencil27", {"g": 9, "mode": "varied"}),
    "sts_lap40rcm": ("laplace2d_5pt", {"g": 40, "rcm": True, "mode": "varied"}),
    "sts_blocks6k": ("shell_blocks", {"n": 6000}),
    "sts_shell4k": ("banded_shell", {"n": 4000, "band": 200}),
}


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def rhs(n, dt):
    return (1.0 + (np.arange(n) % 13) / 16.0).astype(dt)


def main():
    assert O.have_ref(), "oracle/_ref is missing: make -C oracle ref"
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sts")   # (tests/golden/*.npz are the Band-k / SpMV fixtures)
    os.makedirs(out_dir, exist_ok=True)
    for name, (gen, kw) in CASES.items():
        rec = {}
        for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
            rp, ci, v = getattr(G, gen)(dtype=dt, **kw)
            R = O.RefSTS(rp, ci, v)
            a = R.arrays()
            if tag == "f32":
                rec["crc_in_row_ptr"] = crc(rp)
                rec["crc_in_col"] = crc(ci)
                rec["perm"] = a["perm"]
                rec["pack_ptr"] = a["pack_ptr"]
                rec["crc_l_rp"] = crc(a["l_rp"])
                rec["crc_l_ci"] = crc(a["l_ci"])
            else:
                assert np.array_equal(a["perm"], rec["perm"]) and np.array_equal(a["pack_ptr"], rec["pack_ptr"])
            rec["crc_l_val_" + tag] = crc(a["l_val"])
            rec["crc_b_" + tag] = crc(a["b"])
            rec["x_ones_is_exact_" + tag] = np.bool_(np.all(R.solve() == 1))
            R.set_b(rhs(len(rp) - 1, dt))
            rec["x_" + tag] = R.solve()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "n=%d packs=%d" % (len(rec["perm"]), len(rec["pack_ptr"]) - 1))


if __name__ == "__main__":
    main()
