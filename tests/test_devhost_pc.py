"""The flat PC device rules (lala-pc_b200/csrc/pc_device.cuh) compiled for the HOST and compared with the tree-walking
oracle: goldens, the config-3 / config-5 shapes and random networks. A CPU-side check of the code the PC kernel runs."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O
from test_devhost import devhost  # noqa: F401  (fixture: builds tests/native/libdevhost.so)
from test_gpu_pc import random_pc, to_tree


def host_fixpoint(D, props, terms, store):
    s = np.ascontiguousarray(store, dtype=np.int32).copy()
    p = np.ascontiguousarray(props, dtype=np.int32)
    t = np.ascontiguousarray(terms if len(terms) else np.zeros((1, 2)), dtype=np.int32)
    bot, ch = ctypes.c_int(0), ctypes.c_int(0)
    D.devhost_pc_fixpoint(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(p)), t.ctypes.data_as(ctypes.c_void_p),
                          s.ctypes.data_as(ctypes.c_void_p), len(s), ctypes.byref(bot), ctypes.byref(ch))
    return s, bool(bot.value), bool(ch.value)


def host_ask_bits(D, props, terms, store):
    p = np.ascontiguousarray(props, dtype=np.int32)
    t = np.ascontiguousarray(terms, dtype=np.int32)
    s = np.ascontiguousarray(store, dtype=np.int32)
    return np.array([D.devhost_pc_ask(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(i), t.ctypes.data_as(ctypes.c_void_p),
                                      s.ctypes.data_as(ctypes.c_void_p)) for i in range(len(p))], dtype=np.uint8)


def compare(D, formulas, store, label):
    from lala_pc_b200 import pcflat
    props, terms = pcflat.flatten(formulas)
    m = O.PCModel(formulas)
    want, st = m.fixpoint(store)
    got, bot, ch = host_fixpoint(D, props, terms, store)
    assert bot == bool(st.is_bot), label
    if not bot:
        assert np.array_equal(got, want), (label, got.tolist(), want.tolist())
        assert ch == bool(st.has_changed), label
        _, bits = m.ask_all(want, want_bits=True)
        assert np.array_equal(host_ask_bits(D, props, terms, want), bits), label
    return st


def test_goldens_and_flattener(devhost):
    from lala_pc_b200 import pcflat
    ran = 0
    for k in load_golden("pc_kat.json")["props"]:
        formulas = [to_tree(p) for p in k["props"]]
        try:
            pcflat.flatten(formulas)
        except pcflat.Unsupported:
            continue
        ran += 1
        compare(devhost, formulas, np.array(k["store"], dtype=np.int32), k["name"])
    assert ran >= 18
    # shapes without a flat kind are refused, not approximated
    for f in (("le", ("add", ("add", ("var", 0), ("var", 1)), ("var", 2)), ("const", 3)),
              ("le", ("var", 0), ("var", 1)), ("gt", ("var", 0), ("const", 1)),
              ("equiv", ("lit", 0), ("and", ("lit", 1), ("lit", 2)))):
        with pytest.raises(pcflat.Unsupported):
            pcflat.flatten([f])


@pytest.mark.parametrize("cfg", ["config3", "config5"])
def test_config_shapes(devhost, cfg):
    from lala_pc_b200 import workloads as W
    net = getattr(W, cfg)(0.05)
    st = compare(devhost, net.formulas(), net.store, cfg)
    assert not st.is_bot and st.sweeps >= 3


def test_random_networks(devhost):
    rng = np.random.default_rng(7)
    n_ok = 0
    for trial in range(300):
        nvars = int(rng.integers(4, 10))
        forms = random_pc(rng, nvars)
        a = rng.integers(-6, 12, (nvars, 2))
        store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
        store[rng.random(nvars) < 0.4] = (0, 1)
        if rng.random() < 0.2:
            store[int(rng.integers(0, nvars))] = (-2**31, 2**31 - 1)
        st = compare(devhost, forms, store, f"random {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 60
