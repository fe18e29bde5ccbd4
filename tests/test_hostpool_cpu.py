"""CPU: the library's host thread pool and the packer cut into pool pieces (tests/native/pool_selftest.cpp, compiled here against
pgr_tk_b200/csrc/hostpack.cpp): exactly-once visits over hundreds of back-to-back jobs, concurrent callers, pieces == one piece."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_pool_selftest(tmp_path, threads):
    exe = str(tmp_path / "pool_selftest")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "native", "pool_selftest.cpp"),
                           os.path.join(ROOT, "pgr_tk_b200", "csrc", "hostpack.cpp")])
    out = subprocess.run([exe], env=dict(os.environ, PGR_B200_HOST_THREADS=threads), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok threads=%s " % threads), out.stdout
