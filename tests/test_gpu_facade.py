"""The C++ façade (lala-pc_b200/facade/b200pc/pir.hpp) driven by the reference's own test scenarios, re-expressed in
lala-pc_b200/facade/tests/pir_facade_test.cpp. The binary is built by __graft_entry__.build()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACADE = os.path.join(ROOT, "lala-pc_b200", "facade")
BIN = os.path.join(FACADE, "tests", "pir_facade_test")


def test_facade_test_driver_builds():
    r = subprocess.run(["make", "-C", FACADE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_reference_scenarios_through_the_cpp_facade():
    subprocess.run(["make", "-C", FACADE], capture_output=True, text=True)
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout
