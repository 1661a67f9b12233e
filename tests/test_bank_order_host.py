"""CPU-only: the lane phases of the bank-aware row order (parm_b200/csrc/bank_order.cuh) compiled for the host --
every entry of a row survives exactly once, free slots carry their class sentinel, conflicts drop."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bank_order_phases_on_the_host(tmp_path):
    exe = str(tmp_path / "bank_order_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "parm_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "bank_order_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "failures 0" in out.stdout
    cur = float(re.search(r"build order .* = ([0-9.]+) wavefronts", out.stdout).group(1))
    new = float(re.search(r"bank order .* = ([0-9.]+) wavefronts", out.stdout).group(1))
    assert new < 0.65 * cur
