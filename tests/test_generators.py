"""Shapes of the synthetic stand-ins match BASELINE.json / SURVEY.md section 8."""
import numpy as np


def test_config_shapes(gen):
    rp, ci, v = gen.laplace2d_5pt(100)
    assert len(rp) - 1 == 10000 and len(ci) == 5 * 10000 - 4 * 100
    rp, ci, v = gen.stencil27(10)
    assert len(ci) == 28 ** 3 and np.array_equal(rp.astype(np.int64), gen.stencil27_row_ptr(10))
    # cfg 1 / cfg 2 / cfg 5 closed forms
    assert 5 * 1000 ** 2 - 4 * 1000 == 4996000 and (3 * 100 - 2) ** 3 == 26463592
    assert (3 * 400 - 2) ** 3 == 1719374392
    # row blocks are consistent with the full matrix
    full = gen.stencil27(8, mode="varied")
    part = gen.stencil27(8, mode="varied", row_lo=100, row_hi=300)
    e0, e1 = full[0][100], full[0][300]
    assert np.array_equal(part[0], full[0][100:301] - e0)
    assert np.array_equal(part[1], full[1][e0:e1]) and np.array_equal(part[2], full[2][e0:e1])


def test_values_survive_text_format(gen):
    for rp, ci, v in (gen.stencil27(5, mode="varied"), gen.banded_shell(500, band=40), gen.rmat(8, 4)):
        assert np.array_equal(np.array([float("%f" % t) for t in v]), v)
        assert np.array_equal(v.astype(np.float32).astype(np.float64), v)
        for r in range(len(rp) - 1):
            seg = ci[rp[r]:rp[r + 1]]
            assert np.all(np.diff(seg) > 0)


def test_rmat_is_power_law_and_directed(gen):
    rp, ci, v = gen.rmat(12, 8)
    lens = np.diff(rp)
    assert lens.max() > 20 * lens.mean() and (lens == 0).sum() > 0
    assert np.all(v != 0)
