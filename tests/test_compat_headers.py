"""include/spmv.h and include/csrk.h are source-compatible mirrors of the reference headers: drivers written the
way the reference's are (tests/compat/*.c*) must compile and link against them on any machine, and run on a B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIB = os.path.join(ROOT, "heterogeneous-spmv_b200", "lib")
SRC = os.path.join(ROOT, "tests", "compat")
CASES = [("c_spmv_h", "/usr/bin/gcc", "driver_spmv_h.c", ["-std=c99"]),
         ("cxx_csrk_h_f64", "/usr/bin/g++", "driver_csrk_h.cpp", ["-std=c++11", "-DHSPMV_VAL_DOUBLE"]),
         ("cxx_csrk_h_f32", "/usr/bin/g++", "driver_csrk_h.cpp", ["-std=c++11"])]


def build(tmp, name, cc, src, flags):
    exe = os.path.join(str(tmp), name)
    cmd = [cc, "-O1", "-Wall", "-Werror"] + flags + ["-I" + INC, os.path.join(SRC, src), "-o", exe,
                                                    "-L" + LIB, "-lhspmv", "-Wl,-rpath," + LIB, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_style_drivers_compile_and_link(case, tmp_path, hs):
    exe = build(tmp_path, *case)
    r = subprocess.run(["ldd", exe], capture_output=True, text=True)
    assert "libhspmv.so" in r.stdout and "not found" not in r.stdout


@pytest.mark.gpu
def test_reference_style_drivers_run(tmp_path, hs, gen):
    rp, ci, v = gen.stencil27(12, mode="varied")
    f0, f1 = str(tmp_path / "a0.mtx.rcm.csr"), str(tmp_path / "a1.mtx.rcm.csr")
    gen.write_csr_text(f0, rp, ci, v, base=0)
    gen.write_csr_text(f1, rp, ci, v, base=1)
    exe = build(tmp_path, *CASES[0])
    r = subprocess.run([exe, f0], capture_output=True, text=True)
    assert r.returncode == 0 and "Number Wrong: 0" in r.stdout and "size 1728 1728 %d" % len(ci) in r.stdout, r.stdout + r.stderr
    for case in CASES[1:]:
        exe = build(tmp_path, *case)
        for args in (["2", "16"], ["3", "16", "4"]):
            r = subprocess.run([exe, f1] + args, capture_output=True, text=True)
            assert r.returncode == 0 and "Number Wrong: 0" in r.stdout, r.stdout + r.stderr
