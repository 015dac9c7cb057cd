"""The device propagator rules (lala-pc_b200/csrc/pir_device.cuh + pir_div.cuh, incl. the flattened den_fdiv /
den_cdiv) compiled for the HOST by nvcc and compared with the oracle on every interval triple: a CPU-side check of
the register-level code the kernels run, independent of any GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "devhost.cu")
LIB = os.path.join(HERE, "native", "libdevhost.so")
OPS = dict(ADD=2, MUL=4, MIN=6, MAX=7, TDIV=25, FDIV=27, CDIV=29, EDIV=31, EQ=46, LEQ=48)


@pytest.fixture(scope="module")
def devhost():
    csrc = os.path.join(os.path.dirname(HERE), "lala-pc_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("pir_device.cuh", "pir_div.cuh", "pc_device.cuh", "pc_tree.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                        "-o", LIB, SRC], check=True, capture_output=True)
    return ctypes.CDLL(LIB)


@pytest.mark.parametrize("name", list(OPS))
def test_device_rules_on_host_match_oracle(devhost, name):
    lo, hi = -6, 6
    _, fix = O.pir_exhaustive(OPS[name], lo, hi, False, want_fixpoints=True)
    out = np.zeros_like(fix)
    devhost.devhost_exhaustive(OPS[name], lo, hi, out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(out[:, 6], fix[:, 6]), "bot flags differ"
    ok = fix[:, 6] == 0
    assert np.array_equal(out[ok, :6], fix[ok, :6]), "fixpoints differ"


def test_device_ask_on_host_matches_oracle(devhost):
    rng = np.random.default_rng(5)
    for name, op in OPS.items():
        rec = np.array([op, 0, 1, 2], dtype=np.int32)
        for _ in range(3000):
            a = rng.integers(-4, 5, (3, 2))
            store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
            if rng.random() < 0.5:
                store[:, 1] = store[:, 0]
            got = devhost.devhost_ask(store.ctypes.data_as(ctypes.c_void_p), rec.ctypes.data_as(ctypes.c_void_p))
            assert bool(got) == O.pir_ask(store, rec), (name, store.tolist())

