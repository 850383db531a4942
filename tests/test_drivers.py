"""The drop-in executables keep the reference's CLI, stdout protocol and file contracts."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "heterogeneous-spmv_b200", "bin")
REF_REFORMAT = os.path.join(ROOT, "oracle", "_ref", "reformat-auto")


def scrape(stdout):
    """What run_scripts/run_cuda.py:110-122 does: slice after 'TimeMin:' etc. up to the newline."""
    out = {}
    for key in ("TimeMin:", "TimeMax:", "TimeAvg:"):
        i = stdout.find(key)
        assert i >= 0, key + " missing from driver output"
        out[key] = float(stdout[i + 8: stdout.find("\n", i)])
    m = re.search(r"Number Wrong: (\d+) \n", stdout)
    assert m, "'Number Wrong: %d ' line missing"
    out["wrong"] = int(m.group(1))
    return out


def test_reformat_auto_is_byte_identical_to_the_reference_converter(tmp_path, gen):
    if not os.path.exists(REF_REFORMAT):
        pytest.skip("oracle/_ref/reformat-auto not built")
    for name, (rp, ci, v) in (("lap30", gen.laplace2d_5pt(30, dtype=np.float32)),
                              ("st27", gen.stencil27(11, dtype=np.float32, mode="varied")),
                              ("shell", gen.banded_shell(3000, dtype=np.float32, band=80))):
        src = str(tmp_path / (name + ".mtx.rcm.csr"))
        gen.write_csr_text(src, rp, ci, v, base=0)
        ours, theirs = str(tmp_path / (name + ".ours.csr3")), str(tmp_path / (name + ".ref.csr3"))
        o = subprocess.run([os.path.join(BIN, "reformat-csr-to-csr3", "reformat-auto"), src, ours], capture_output=True, text=True)
        r = subprocess.run([REF_REFORMAT, src, theirs], capture_output=True, text=True)
        assert o.returncode == 0 and r.returncode == 0
        line = [l for l in r.stdout.splitlines() if l.startswith("using ssrs")]
        assert line and line[0] in o.stdout            # same "using ssrs %d, srs %d" line
        assert open(ours, "rb").read() == open(theirs, "rb").read(), name


def test_usage_lines_match_the_reference():
    r = subprocess.run([os.path.join(BIN, "cuda-spmv-csr", "spmv.exe")], capture_output=True, text=True)
    assert r.returncode == 0 and "./spmv.exe inputfie.csr num_runs" in r.stdout        # spmv-csr/spmv.c:119 (sic)
    r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv")], capture_output=True, text=True)
    assert r.returncode == 0 and "inputfile num_runs super_super_row_size super_row_size" in r.stdout
    r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv-auto-ampere"), "x"], capture_output=True, text=True)
    assert r.returncode == 0 and "inputfile num_runs" in r.stdout


@pytest.mark.gpu
def test_drivers_on_gpu(tmp_path, gen):
    rp, ci, v = gen.stencil27(20, mode="varied")
    f0, f1 = str(tmp_path / "m0.mtx.rcm.csr"), str(tmp_path / "m1.mtx.rcm.csr")
    gen.write_csr_text(f0, rp, ci, v, base=0)      # cuda-spmv-csr reads 0-based
    gen.write_csr_text(f1, rp, ci, v, base=1)      # cuda-spmv-csrk reads 1-based (Q1)
    for extra in ([], ["--fp32"]):
        r = subprocess.run([os.path.join(BIN, "cuda-spmv-csr", "spmv.exe"), f0, "5"] + extra, capture_output=True, text=True)
        s = scrape(r.stdout)
        assert r.returncode == 0 and s["wrong"] == 0 and 0 < s["TimeMin:"] <= s["TimeAvg:"] <= s["TimeMax:"]
        r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv"), f1, "5", "16"] + extra, capture_output=True, text=True)
        s = scrape(r.stdout)
        assert r.returncode == 0 and s["wrong"] == 0 and "size 8000 8000 %d" % len(ci) in r.stdout
        r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv"), f1, "5", "16", "8"] + extra, capture_output=True, text=True)
        assert scrape(r.stdout)["wrong"] == 0 and "k=3" in r.stdout
        for auto in ("volta", "ampere", "b200"):
            r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv-auto-" + auto), f1, "3"] + extra, capture_output=True, text=True)
            assert scrape(r.stdout)["wrong"] == 0 and "using ssrs" in r.stdout
            assert ("Tuned: csrk_tma" in r.stdout) == (auto == "b200")     # the B200 flavour measures, the others use formulas
    # --cache: the second run of the harness on the same file loads the binary side file
    exe = os.path.join(BIN, "cuda-spmv-csr", "spmv.exe")
    r1 = subprocess.run([exe, f0, "2", "--cache"], capture_output=True, text=True)
    r2 = subprocess.run([exe, f0, "2", "--cache"], capture_output=True, text=True)
    assert "Matrix from text" in r1.stdout and "Matrix from the binary side file" in r2.stdout
    assert scrape(r1.stdout)["wrong"] == 0 and scrape(r2.stdout)["wrong"] == 0 and os.path.exists(f0 + ".hspmvbin")
    # a 0-based file fed to the 1-based reader is rejected, not mis-read silently
    r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv"), f0, "2", "16"], capture_output=True, text=True)
    assert "Error" in r.stdout
