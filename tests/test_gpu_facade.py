"""The C++ façades (lala-pc_b200/facade/b200pc/{pir,pc}.hpp) driven by the reference's own test scenarios, re-expressed
in lala-pc_b200/facade/tests/{pir,pc}_facade_test.cpp. The binaries are built by __graft_entry__.build()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACADE = os.path.join(ROOT, "lala-pc_b200", "facade")
BIN = os.path.join(FACADE, "tests", "pir_facade_test")
PC_BIN = os.path.join(FACADE, "tests", "pc_facade_test")


def test_facade_test_driver_builds():
    r = subprocess.run(["make", "-C", FACADE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(BIN) and os.path.exists(PC_BIN)


@pytest.mark.gpu
def test_reference_scenarios_through_the_cpp_facade():
    subprocess.run(["make", "-C", FACADE], capture_output=True, text=True)
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout


@pytest.mark.gpu
def test_reference_pc_scenarios_through_the_cpp_facade():
    """tests/pc_test.cpp and tests/pc_bitset_test.cpp scenarios through b200pc::PC<VStore> / PC<BitVStore>."""
    subprocess.run(["make", "-C", FACADE], capture_output=True, text=True)
    r = subprocess.run([PC_BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout


@pytest.mark.gpu
def test_reference_flatzinc_models_through_the_cpp_ternariser():
    """tests/pir_test.cpp's FlatZinc models: b200pc::Ternarizer -> PIR::interpret_tell -> device fixpoint, with the
    reference's expected intervals and deduction counts."""
    subprocess.run(["make", "-C", FACADE], capture_output=True, text=True)
    r = subprocess.run([os.path.join(FACADE, "tests", "ternarize_facade_test")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout
