"""AddressSanitizer + UndefinedBehaviorSanitizer passes over the product's host code (no device needed): the host
sources are compiled a second time with -fsanitize=address,undefined next to small harnesses (tools/asan_*.cpp) and run
on random inputs -- the index work (Band-k with the serial AND the level-synchronous RCM walk, STS level sets, bins,
blocks, tiles, partition, pipeline plan) and the text readers on valid, truncated, corrupted and absurd-header files."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "heterogeneous-spmv_b200", "csrc")


def _build(tmp, name, harness, sources):
    exe = os.path.join(str(tmp), name)
    cmd = ["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", os.path.join(ROOT, "tools", harness)] + [os.path.join(CSRC, s) for s in sources] + ["-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("cannot find -lasan" in r.stderr or "cannot find -lubsan" in r.stderr or "libasan" in r.stderr):
        pytest.skip("sanitizer runtimes not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _run(exe, extra_env=None, args=()):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    env.update(extra_env or {})
    r = subprocess.run([exe] + list(args), capture_output=True, text=True, env=env, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]
    return r.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_host_index_work_is_clean_under_asan_ubsan(tmp_path):
    exe = _build(tmp_path, "asan_host", "asan_host_index.cpp", ["bandk.cpp", "sts.cpp", "partition.cpp"])
    assert _run(exe, args=["70"]).strip().endswith("ok")                                                       # serial RCM walk (small graphs)
    assert _run(exe, {"HSPMV_RCM": "parallel", "HSPMV_RCM_CUTOFF": "3"}, ["70"]).strip().endswith("ok")        # level-synchronous walk
    assert _run(exe, {"HSPMV_STS_LEVELS": "literal"}, ["40"]).strip().endswith("ok")                           # literal STS level search


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_text_readers_are_clean_under_asan_ubsan(tmp_path):
    exe = _build(tmp_path, "asan_io", "asan_io.cpp", ["io.cpp"])
    out = _run(exe)
    assert "csr ok" in out and "csr3 ok" in out
    # the valid files must have been accepted, the absurd ones rejected
    words = out.replace(";", " ").split()
    ok, rej, ok3, rej3 = int(words[2]), int(words[4]), int(words[7]), int(words[9])
    assert ok >= 400 and rej >= 400 and ok3 >= 200 and rej3 >= 100, out
