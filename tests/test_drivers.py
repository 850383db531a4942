"""The drop-in executables keep the reference's CLI, stdout protocol and file contracts."""
import os
import re
import subprocess
import sys

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


def test_run_scripts_compatible_runner_records(tmp_path):
    """heterogeneous-spmv_b200/run_scripts/run_b200.py scrapes and records like run_scripts/run_cuda.py:110-125 and
    run_tuning.py:125; a stand-in driver that prints the protocol lets this run without a GPU"""
    fake_bin = tmp_path / "bin"
    for sub, name in (("cuda-spmv-csr", "spmv.exe"), ("cuda-spmv-csrk/cuda", "spmv-auto-b200"), ("cuda-spmv-csrk/cuda", "spmv")):
        d = fake_bin / sub
        d.mkdir(parents=True, exist_ok=True)
        exe = d / name
        exe.write_text("#!/bin/sh\necho \"banner $@\"\necho \"TimeMin: 1.5e-04\"\necho \"TimeMax: 3e-04\"\necho \"TimeAvg: 2e-04\"\necho \"Number Wrong: 0 \"\n")
        exe.chmod(0o755)
    mats = tmp_path / "mats"
    mats.mkdir()
    for n in ("b.mtx.rcm.csr", "a.mtx.rcm.csr", "a.mtx.rcm.csr.hspmvbin"):
        (mats / n).write_text("1 1 1\n")
    runner = os.path.join(ROOT, "heterogeneous-spmv_b200", "run_scripts", "run_b200.py")
    rec = tmp_path / "rec.csv"
    common = [sys.executable, runner, "--matrices", str(mats), "--bin", str(fake_bin), "--record", str(rec), "--out", str(tmp_path / "runs"),
              "--threads", "16", "--runs", "7"]
    for kernel, extra in (("cuda-spmv-csr", []), ("cuda-spmv-csrk", ["--extra", "--fp32 --cache"]), ("tuning", ["--sizes", "16,24"])):
        r = subprocess.run(common + ["--kernel", kernel] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lines = rec.read_text().splitlines()
    assert lines[0] == "cuda-spmv-csr, a.mtx.rcm.csr, cuda, 1, 16,  1.5e-04,  3e-04,  2e-04, "          # run_cuda.py:125
    assert lines[1].startswith("cuda-spmv-csr, b.mtx.rcm.csr") and lines[2].startswith("cuda-spmv-csrk, a.mtx.rcm.csr")
    assert lines[4] == "cuda-spmv-csrk/cuda, a.mtx.rcm.csr, cuda, 1, 16, (16),  1.5e-04,  3e-04,  2e-04, "  # run_tuning.py:125
    assert len(lines) == 2 + 2 + 4 and not any("hspmvbin" in t for t in lines)
    out = (tmp_path / "runs" / "cuda-spmv-csrk" / "a.mtx.rcm.csr_cuda_1_16.txt").read_text()
    assert "a.mtx.rcm.csr 7 --fp32 --cache" in out


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
