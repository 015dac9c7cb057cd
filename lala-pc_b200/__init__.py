"""lala-pc_b200 — host-side binding of the B200 propagation engine (the C-ABI of include/lpc.h).

Thin ctypes layer, no compute of its own: every call goes to liblpc.so (hand-written CUDA for sm_100a). There is
no CPU fallback: importing works without a GPU (so that symbols can be inspected), but creating a table, a store
or a batch raises `LpcError` when no CUDA device is present, and a missing liblpc.so raises at import time.

Mirrors the reference's PIR interface (lala-pc include/lala/pir.hpp): `PIR.num_deductions`, `load_deduce`,
`deduce(i)`, `ask(i)`, `embed`, `is_bot`, `is_top`, `__getitem__`, `vars`, `snapshot/restore`,
`is_extractable`, `extract`, plus `fixpoint()` for the GaussSeidelIteration loop the tests drive by hand.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LPC_LIB") or os.path.join(_HERE, "liblpc.so")   # LPC_LIB: A/B runs of two builds on one box

# lala-core Sig values of the operators PIR accepts (include/lpc.h: enum lpc_sig)
ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48
SIG_NAMES = {ADD: "ADD", MUL: "MUL", MIN: "MIN", MAX: "MAX", TDIV: "TDIV", FDIV: "FDIV", CDIV: "CDIV",
             EDIV: "EDIV", EQ: "EQ", LEQ: "LEQ"}
MODE_AUTO, MODE_SWEEP, MODE_WORKLIST = 0, 1, 2
INT_MIN, INT_MAX = -2**31, 2**31 - 1


class LpcError(RuntimeError):
    pass


class FixpointOpts(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("max_sweeps", ctypes.c_int32), ("stop_on_bot", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("stream", ctypes.c_uint64)]


class FixpointResult(ctypes.Structure):
    _fields_ = [("has_changed", ctypes.c_int32), ("is_bot", ctypes.c_int32), ("sweeps", ctypes.c_int32),
                ("dense_sweeps", ctypes.c_int32), ("deductions", ctypes.c_int64), ("device_ms", ctypes.c_float),
                ("overflow_hazard", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SearchOpts(ctypes.Structure):
    _fields_ = [("max_nodes", ctypes.c_int64), ("max_depth", ctypes.c_int32), ("objective_var", ctypes.c_int32),
                ("stream", ctypes.c_uint64), ("change_driven", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class SearchResult(ctypes.Structure):
    _fields_ = [("n_solutions", ctypes.c_int64), ("n_nodes", ctypes.c_int64), ("n_fails", ctypes.c_int64),
                ("n_unknown_leaves", ctypes.c_int64), ("n_incomplete", ctypes.c_int64), ("sweeps_total", ctypes.c_int64),
                ("deductions", ctypes.c_int64), ("best_bound", ctypes.c_int32), ("max_depth_seen", ctypes.c_int32),
                ("device_ms", ctypes.c_float), ("reserved", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class BatchResult(ctypes.Structure):
    _fields_ = [("n_bot", ctypes.c_int64), ("n_solution", ctypes.c_int64), ("n_unknown", ctypes.c_int64),
                ("best_bound", ctypes.c_int32), ("max_sweeps_seen", ctypes.c_int32),
                ("sweeps_total", ctypes.c_int64), ("deductions", ctypes.c_int64), ("device_ms", ctypes.c_float),
                ("overflow_hazard", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class EpsResult(ctypes.Structure):
    _fields_ = [("n_bot", ctypes.c_int64), ("n_solution", ctypes.c_int64), ("n_unknown", ctypes.c_int64),
                ("best_bound", ctypes.c_int32), ("max_sweeps_seen", ctypes.c_int32), ("sweeps_total", ctypes.c_int64),
                ("deductions", ctypes.c_int64), ("n_survivors", ctypes.c_int64), ("n_live_records", ctypes.c_int32),
                ("device_ms", ctypes.c_float), ("overflow_hazard", ctypes.c_int32), ("n_first_sweep_records", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback.")

_L = ctypes.CDLL(LIB_PATH)
_vp, _i32, _i64, _u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
_pi32 = ctypes.POINTER(ctypes.c_int32)
_pint = ctypes.POINTER(ctypes.c_int)
_pu8 = ctypes.POINTER(ctypes.c_uint8)
_pvp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); also the list the symbol-export test checks against include/lpc.h
SIGNATURES = {
    "lpc_version": (ctypes.c_char_p, []),
    "lpc_last_error": (ctypes.c_char_p, []),
    "lpc_device_init": (ctypes.c_int, [ctypes.c_int]),
    "lpc_device_count": (ctypes.c_int, [_pint]),
    "lpc_launch_count": (_i64, []),
    "lpc_measure_l2_copy_gbs": (ctypes.c_int, [_i64, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "lpc_table_create": (ctypes.c_int, [_vp, _i64, _i32, _pvp]),
    "lpc_table_destroy": (ctypes.c_int, [_vp]),
    "lpc_table_create_empty": (ctypes.c_int, [_i32, _pvp]),
    "lpc_table_append": (ctypes.c_int, [_vp, _vp, _i64]),
    "lpc_table_set_nvars": (ctypes.c_int, [_vp, _i32]),
    "lpc_table_truncate": (ctypes.c_int, [_vp, _i64]),
    "lpc_table_finalize": (ctypes.c_int, [_vp, _i32]),
    "lpc_table_uploaded_bytes": (_i64, [_vp]),
    "lpc_table_size": (_i64, [_vp]),
    "lpc_table_nvars": (_i32, [_vp]),
    "lpc_table_load": (ctypes.c_int, [_vp, _i64, _pi32]),
    "lpc_table_clamp_reified": (ctypes.c_int, [_vp, _vp]),
    "lpc_store_create": (ctypes.c_int, [_i32, _pvp]),
    "lpc_store_wrap_device": (ctypes.c_int, [_vp, _i32, _pvp]),
    "lpc_store_destroy": (ctypes.c_int, [_vp]),
    "lpc_store_nvars": (_i32, [_vp]),
    "lpc_store_device_ptr": (_vp, [_vp]),
    "lpc_store_write": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_store_read": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_store_embed": (ctypes.c_int, [_vp, _i32, _i32, _i32, _pint]),
    "lpc_store_copy": (ctypes.c_int, [_vp, _vp]),
    "lpc_store_is_bot": (ctypes.c_int, [_vp, _pint]),
    "lpc_store_is_top": (ctypes.c_int, [_vp, _pint]),
    "lpc_fixpoint_default_opts": (None, [ctypes.POINTER(FixpointOpts)]),
    "lpc_fixpoint": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), ctypes.POINTER(FixpointResult)]),
    "lpc_fixpoint_async": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts)]),
    "lpc_fixpoint_collect": (ctypes.c_int, [_vp, ctypes.POINTER(FixpointResult)]),
    "lpc_fixpoint_host": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), ctypes.POINTER(FixpointResult)]),
    "lpc_deduce_one": (ctypes.c_int, [_vp, _vp, _i64, _pint]),
    "lpc_ask_one": (ctypes.c_int, [_vp, _vp, _i64, _pint]),
    "lpc_ask_all": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_int64)]),
    "lpc_ask_bits": (ctypes.c_int, [_vp, _vp, _pu8]),
    "lpc_batch_create": (ctypes.c_int, [_vp, _i32, _pvp]),
    "lpc_batch_destroy": (ctypes.c_int, [_vp]),
    "lpc_batch_device_ptr": (_vp, [_vp]),
    "lpc_batch_write": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_batch_read": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_batch_init_split": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i64]),
    "lpc_batch_init_split_ids": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "lpc_batch_set_seeds": (ctypes.c_int, [_vp, _vp, _i32]),
    "lpc_batch_fixpoint": (ctypes.c_int, [_vp, ctypes.POINTER(FixpointOpts), _i32, ctypes.POINTER(BatchResult)]),
    "lpc_batch_fixpoint_async": (ctypes.c_int, [_vp, ctypes.POINTER(FixpointOpts), _i32]),
    "lpc_batch_collect": (ctypes.c_int, [_vp, ctypes.POINTER(BatchResult)]),
    "lpc_batch_fixpoint_host": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), _i32,
                                               ctypes.POINTER(BatchResult)]),
    "lpc_batch_flags": (ctypes.c_int, [_vp, _pu8]),
    "lpc_batch_reduction_device_ptr": (_vp, [_vp]),
    "lpc_batch_set_rank": (ctypes.c_int, [_vp, _i32, _i32]),
    "lpc_batch_payload_device_ptr": (_vp, [_vp, _pi32]),
    "lpc_eps_create": (ctypes.c_int, [_vp, _i32, _i32, _pvp]),
    "lpc_eps_destroy": (ctypes.c_int, [_vp]),
    "lpc_eps_set_rank": (ctypes.c_int, [_vp, _i32, _i32]),
    "lpc_eps_payload_device_ptr": (_vp, [_vp, _pi32]),
    "lpc_eps_solve_host": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, ctypes.POINTER(FixpointOpts), _i32,
                                          _vp, _vp, _vp, _i32, _pi32, ctypes.POINTER(EpsResult)]),
    "lpc_eps_upload": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _i32]),
    "lpc_eps_run_async": (ctypes.c_int, [_vp, ctypes.POINTER(FixpointOpts), _i32]),
    "lpc_eps_collect": (ctypes.c_int, [_vp, ctypes.POINTER(EpsResult)]),
    "lpc_eps_download": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _pi32]),
    "lpc_eps_sweeps": (ctypes.c_int, [_vp, _vp]),
    "lpc_eps_peer_export": (ctypes.c_int, [_vp, _vp]),
    "lpc_eps_peer_connect": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_eps_peer_disconnect": (ctypes.c_int, [_vp]),
    "lpc_search_default_opts": (None, [ctypes.POINTER(SearchOpts)]),
    "lpc_batch_search": (ctypes.c_int, [_vp, _vp, _i32, ctypes.POINTER(SearchOpts), ctypes.POINTER(SearchResult), _vp]),
    # include/lpc_pc.h
    "lpc_pc_table_create": (ctypes.c_int, [_vp, _i64, _vp, _i64, _i32, _pvp]),
    "lpc_pc_table_destroy": (ctypes.c_int, [_vp]),
    "lpc_pc_table_size": (_i64, [_vp]),
    "lpc_pc_table_terms": (_i64, [_vp]),
    "lpc_pc_fixpoint": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), ctypes.POINTER(FixpointResult)]),
    "lpc_pc_fixpoint_host": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), ctypes.POINTER(FixpointResult)]),
    "lpc_pc_deduce_one": (ctypes.c_int, [_vp, _vp, _i64, _pint]),
    "lpc_pc_ask_all": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_int64), _pu8]),
    "lpc_store_write_bits": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_store_read_bits": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "lpc_nbit_range": (_u64, [_i32, _i32]),
    "lpc_store_embed_bits": (ctypes.c_int, [_vp, _i32, _u64, _pint]),
    "lpc_store_is_bot_bits": (ctypes.c_int, [_vp, _pint]),
    "lpc_store_is_top_bits": (ctypes.c_int, [_vp, _pint]),
    "lpc_pc_fixpoint_bits": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), ctypes.POINTER(FixpointResult)]),
    "lpc_pc_fixpoint_bits_host": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(FixpointOpts), ctypes.POINTER(FixpointResult)]),
    "lpc_pc_deduce_one_bits": (ctypes.c_int, [_vp, _vp, _i64, _pint]),
    "lpc_pc_ask_all_bits": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_int64), _pu8]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(_L, _name)
    _f.restype = _res
    _f.argtypes = _args

lib = _L


def _check(rc):
    if rc != 0:
        raise LpcError(f"lpc error {rc}: {_L.lpc_last_error().decode()}")


def device_count():
    n = ctypes.c_int(0)
    _L.lpc_device_count(ctypes.byref(n))
    return n.value


def device_init(device=0):
    _check(_L.lpc_device_init(device))


def launch_count():
    return int(_L.lpc_launch_count())


def sort_records(records):
    """The host-side ordering of PIR::deduce(tell) (pir.hpp:343-347): stable sort by (op, y, x, z)."""
    r = np.ascontiguousarray(records, dtype=np.int32)
    order = np.lexsort((r[:, 3], r[:, 1], r[:, 2], r[:, 0]))   # last key is primary
    return r[order]


def _opts(mode=MODE_AUTO, max_sweeps=0, stop_on_bot=True, stream=0, switch_div=0):
    o = FixpointOpts()
    _L.lpc_fixpoint_default_opts(ctypes.byref(o))
    o.mode, o.max_sweeps, o.stop_on_bot, o.stream, o.reserved = mode, max_sweeps, int(stop_on_bot), stream, switch_div
    return o


class Table:
    """Immutable device propagator table (battery::vector<bytecode_type>, pir.hpp:104,115-118)."""

    def __init__(self, records=None, nvars=0):
        """`records` None: an empty table to be grown by append() + finalize() (PIR::deduce(tell), pir.hpp:326-352)."""
        self._h = ctypes.c_void_p()
        if records is None:
            _check(_L.lpc_table_create_empty(nvars, ctypes.byref(self._h)))
        else:
            r = np.ascontiguousarray(records, dtype=np.int32).reshape(-1, 4)
            _check(_L.lpc_table_create(r.ctypes.data, r.shape[0], nvars, ctypes.byref(self._h)))
        self.nvars = nvars

    def append(self, records):
        r = np.ascontiguousarray(records, dtype=np.int32).reshape(-1, 4)
        _check(_L.lpc_table_append(self._h, r.ctypes.data, r.shape[0]))

    def set_nvars(self, nvars):
        _check(_L.lpc_table_set_nvars(self._h, nvars))
        self.nvars = nvars

    def truncate(self, n):
        _check(_L.lpc_table_truncate(self._h, n))

    def finalize(self, sort=True):
        """Bring the device image up to date; sort=True: stable sort by (op, y, x, z) first (pir.hpp:343-347)."""
        _check(_L.lpc_table_finalize(self._h, int(sort)))

    def uploaded_bytes(self):
        return int(_L.lpc_table_uploaded_bytes(self._h))

    def records(self):
        n = len(self)
        out = np.empty((n, 4), dtype=np.int32)
        row = (ctypes.c_int32 * 4)()
        for i in range(n):
            _check(_L.lpc_table_load(self._h, i, row))
            out[i] = tuple(row)
        return out

    def __len__(self):
        return int(_L.lpc_table_size(self._h))

    def load(self, i):
        out = (ctypes.c_int32 * 4)()
        _check(_L.lpc_table_load(self._h, i, out))
        return tuple(out)

    def close(self):
        if getattr(self, "_h", None):
            _L.lpc_table_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Store:
    """Device interval store (VStore<Interval<ZLB>>): nvars pairs {lb, ub}."""

    def __init__(self, nvars=None, values=None, device_ptr=None):
        self._h = ctypes.c_void_p()
        if device_ptr is not None:
            _check(_L.lpc_store_wrap_device(device_ptr, nvars, ctypes.byref(self._h)))
        else:
            if values is not None:
                values = np.ascontiguousarray(values, dtype=np.int32).reshape(-1, 2)
                nvars = values.shape[0]
            _check(_L.lpc_store_create(nvars, ctypes.byref(self._h)))
            if values is not None:
                self.write(values)
        self.nvars = nvars

    def write(self, values, first=0):
        v = np.ascontiguousarray(values, dtype=np.int32).reshape(-1, 2)
        _check(_L.lpc_store_write(self._h, first, v.shape[0], v.ctypes.data))

    def read(self, first=0, n=None):
        n = self.nvars - first if n is None else n
        out = np.empty((n, 2), dtype=np.int32)
        _check(_L.lpc_store_read(self._h, first, n, out.ctypes.data))
        return out

    def write_bits(self, cells, first=0):
        """The same cells as one uint64 NBitset<64> per variable (include/lpc_pc.h)."""
        v = np.ascontiguousarray(cells, dtype=np.uint64).reshape(-1)
        _check(_L.lpc_store_write_bits(self._h, first, v.shape[0], v.ctypes.data))

    def read_bits(self, first=0, n=None):
        n = self.nvars - first if n is None else n
        out = np.empty(n, dtype=np.uint64)
        _check(_L.lpc_store_read_bits(self._h, first, n, out.ctypes.data))
        return out

    def embed(self, var, lb, ub):
        c = ctypes.c_int(0)
        _check(_L.lpc_store_embed(self._h, var, lb, ub, ctypes.byref(c)))
        return bool(c.value)

    def copy_from(self, other):
        _check(_L.lpc_store_copy(self._h, other._h))

    def is_bot(self):
        c = ctypes.c_int(0)
        _check(_L.lpc_store_is_bot(self._h, ctypes.byref(c)))
        return bool(c.value)

    def is_top(self):
        c = ctypes.c_int(0)
        _check(_L.lpc_store_is_top(self._h, ctypes.byref(c)))
        return bool(c.value)

    @property
    def device_ptr(self):
        return _L.lpc_store_device_ptr(self._h)

    def close(self):
        if getattr(self, "_h", None):
            _L.lpc_store_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def measure_l2_copy_gbs(nbytes=32 << 20, iters=20):
    """Copy bandwidth (read + write, GB/s) of an L2-resident buffer pair: the on-chip ceiling bench.py reports next to HBM."""
    g = ctypes.c_double(0.0)
    _check(_L.lpc_measure_l2_copy_gbs(nbytes, iters, ctypes.byref(g)))
    return g.value


def fixpoint(table, store, **kw):
    """GaussSeidelIteration{}.fixpoint(n, deduce) on the device (tests/pir_test.cpp:60-62)."""
    o, r = _opts(**kw), FixpointResult()
    _check(_L.lpc_fixpoint(table._h, store._h, ctypes.byref(o), ctypes.byref(r)))
    return r


def fixpoint_async(table, store, **kw):
    o = _opts(**kw)
    _check(_L.lpc_fixpoint_async(table._h, store._h, ctypes.byref(o)))


def fixpoint_collect(store):
    r = FixpointResult()
    _check(_L.lpc_fixpoint_collect(store._h, ctypes.byref(r)))
    return r


def fixpoint_host(table, values, **kw):
    """Host-buffer entry point: values ([nvars,2] int32, or a raw address) is updated in place."""
    o, r = _opts(**kw), FixpointResult()
    ptr = values if isinstance(values, int) else values.ctypes.data
    _check(_L.lpc_fixpoint_host(table._h, ptr, ctypes.byref(o), ctypes.byref(r)))
    return r


PC_LIN_LE, PC_REIF_LIN_LE, PC_EQ, PC_NEQ, PC_CLAUSE, PC_ABS_EQ = 1, 2, 3, 4, 5, 6
PC_LIN_GE, PC_LIN_GT, PC_LIN_EQ, PC_LIN_EQ_VAR = 7, 8, 9, 10


def nbit_range(lb, ub):
    """NBitset<64>(lb, ub) as a uint64 cell (scalar or arrays)."""
    lb, ub = np.asarray(lb, dtype=np.int64), np.asarray(ub, dtype=np.int64)
    frm = np.where(lb < 0, 0, np.where(lb >= 62, 63, lb + 1)).astype(np.uint64)
    to = np.where(ub < 0, 0, np.where(ub >= 62, 63, ub + 1)).astype(np.uint64)
    ones = np.uint64(0xFFFFFFFFFFFFFFFF)
    cells = (ones << frm) & (ones >> (np.uint64(63) - to))
    return np.where(lb > ub, np.uint64(0), cells).astype(np.uint64)


def nbit_from_intervals(store):
    """An interval store [n,2] as NBitset<64> cells (exact for domains inside [0, 61])."""
    s = np.asarray(store).reshape(-1, 2)
    return nbit_range(s[:, 0], s[:, 1])


class PcTable:
    """Flattened PC propagators (include/lpc_pc.h): props [n,5] int32 {kind, first_term, n_terms, rhs, bvar},
    terms [m,2] int32 {coef, var}."""

    def __init__(self, props, terms, nvars):
        p = np.ascontiguousarray(props, dtype=np.int32).reshape(-1, 5)
        t = np.ascontiguousarray(terms, dtype=np.int32).reshape(-1, 2)
        self._h = ctypes.c_void_p()
        self.nvars = nvars
        _check(_L.lpc_pc_table_create(p.ctypes.data, p.shape[0], t.ctypes.data, t.shape[0], nvars, ctypes.byref(self._h)))

    def __len__(self):
        return int(_L.lpc_pc_table_size(self._h))

    def num_terms(self):
        return int(_L.lpc_pc_table_terms(self._h))

    def fixpoint(self, store, bitset=False, **kw):
        """`bitset=True`: the store holds NBitset<64> cells (Store.write_bits), see include/lpc_pc.h."""
        o, r = _opts(**kw), FixpointResult()
        f = _L.lpc_pc_fixpoint_bits if bitset else _L.lpc_pc_fixpoint
        _check(f(self._h, store._h, ctypes.byref(o), ctypes.byref(r)))
        return r

    def fixpoint_host(self, values, bitset=False, **kw):
        o, r = _opts(**kw), FixpointResult()
        ptr = values if isinstance(values, int) else values.ctypes.data
        f = _L.lpc_pc_fixpoint_bits_host if bitset else _L.lpc_pc_fixpoint_host
        _check(f(self._h, ptr, ctypes.byref(o), ctypes.byref(r)))
        return r

    def deduce(self, store, i, bitset=False):
        c = ctypes.c_int(0)
        f = _L.lpc_pc_deduce_one_bits if bitset else _L.lpc_pc_deduce_one
        _check(f(self._h, store._h, i, ctypes.byref(c)))
        return bool(c.value)

    def ask_all(self, store, want_bits=False, bitset=False):
        n = ctypes.c_int64(0)
        bits = np.zeros(max(1, len(self)), dtype=np.uint8) if want_bits else None
        f = _L.lpc_pc_ask_all_bits if bitset else _L.lpc_pc_ask_all
        _check(f(self._h, store._h, ctypes.byref(n), bits.ctypes.data_as(_pu8) if want_bits else None))
        return (int(n.value), bits[:len(self)]) if want_bits else int(n.value)

    def close(self):
        if getattr(self, "_h", None):
            _L.lpc_pc_table_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Batch:
    """n_stores independent stores over one shared table (one thread block per store)."""

    def __init__(self, table, n_stores):
        self._h = ctypes.c_void_p()
        self.table, self.n_stores, self.nvars = table, n_stores, table.nvars
        _check(_L.lpc_batch_create(table._h, n_stores, ctypes.byref(self._h)))

    def write(self, values, first=0):
        v = np.ascontiguousarray(values, dtype=np.int32).reshape(-1, self.nvars, 2)
        _check(_L.lpc_batch_write(self._h, first, v.shape[0], v.ctypes.data))

    def read(self, first=0, n=None):
        n = self.n_stores - first if n is None else n
        out = np.empty((n, self.nvars, 2), dtype=np.int32)
        _check(_L.lpc_batch_read(self._h, first, n, out.ctypes.data))
        return out

    def init_split(self, base, decision_vars, first_id=0, ids=None):
        """EPS decomposition on the device; `ids` (int64 [n_stores]) overrides the consecutive ids first_id + k."""
        b = np.ascontiguousarray(base, dtype=np.int32).reshape(-1, 2)
        d = np.ascontiguousarray(decision_vars, dtype=np.int32)
        if ids is None:
            _check(_L.lpc_batch_init_split(self._h, b.ctypes.data, d.ctypes.data, d.shape[0], first_id))
        else:
            i = np.ascontiguousarray(ids, dtype=np.int64)
            assert i.shape == (self.n_stores,)
            _check(_L.lpc_batch_init_split_ids(self._h, b.ctypes.data, d.ctypes.data, d.shape[0], i.ctypes.data))

    def fixpoint(self, objective_var=-1, **kw):
        o, r = _opts(**kw), BatchResult()
        _check(_L.lpc_batch_fixpoint(self._h, ctypes.byref(o), objective_var, ctypes.byref(r)))
        return r

    def fixpoint_async(self, objective_var=-1, **kw):
        o = _opts(**kw)
        _check(_L.lpc_batch_fixpoint_async(self._h, ctypes.byref(o), objective_var))

    def collect(self):
        r = BatchResult()
        _check(_L.lpc_batch_collect(self._h, ctypes.byref(r)))
        return r

    def fixpoint_host(self, values, objective_var=-1, **kw):
        o, r = _opts(**kw), BatchResult()
        ptr = values if isinstance(values, int) else values.ctypes.data
        _check(_L.lpc_batch_fixpoint_host(self._h, ptr, ctypes.byref(o), objective_var, ctypes.byref(r)))
        return r

    def flags(self):
        out = np.empty(self.n_stores, dtype=np.uint8)
        _check(_L.lpc_batch_flags(self._h, out.ctypes.data_as(_pu8)))
        return out

    @property
    def device_ptr(self):
        return _L.lpc_batch_device_ptr(self._h)

    def set_seeds(self, variables):
        """Promise that the stores are fixpoints of the table except on `variables` (None withdraws it): include/lpc.h."""
        if variables is None:
            _check(_L.lpc_batch_set_seeds(self._h, None, -1))
        else:
            v = np.ascontiguousarray(variables, dtype=np.int32)
            _check(_L.lpc_batch_set_seeds(self._h, v.ctypes.data, v.shape[0]))

    @property
    def reduction_device_ptr(self):
        return _L.lpc_batch_reduction_device_ptr(self._h)

    def set_rank(self, rank, world):
        _check(_L.lpc_batch_set_rank(self._h, rank, world))

    @property
    def payload(self):
        """(device address, length) of the int64 all-reduce payload (include/lpc.h: lpc_batch_set_rank)."""
        n = ctypes.c_int32(0)
        p = _L.lpc_batch_payload_device_ptr(self._h, ctypes.byref(n))
        return p, n.value

    def search(self, branch_vars, objective_var=-1, max_nodes=0, max_depth=64, stream=0, want_per_store=True,
               change_driven=None):
        """Depth-first search from every store of the batch (include/lpc.h: lpc_batch_search).
        Returns (SearchResult, int64 [n_stores, 6] {solutions, nodes, fails, best, incomplete, unknown_leaves} or None)."""
        bv = np.ascontiguousarray(branch_vars, dtype=np.int32)
        o, r = SearchOpts(), SearchResult()
        _L.lpc_search_default_opts(ctypes.byref(o))
        o.max_nodes, o.max_depth, o.objective_var, o.stream = max_nodes, max_depth, objective_var, stream
        o.change_driven = -1 if change_driven is None else int(change_driven)   # None: by table size
        per = np.zeros((self.n_stores, 6), dtype=np.int64) if want_per_store else None
        _check(_L.lpc_batch_search(self._h, bv.ctypes.data, bv.shape[0], ctypes.byref(o), ctypes.byref(r),
                                   per.ctypes.data if want_per_store else None))
        return r, per

    def close(self):
        if getattr(self, "_h", None):
            _L.lpc_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Eps:
    """The EPS-native call (include/lpc.h: lpc_eps_*): root store + decision variables + subproblem ids in; flags, the
    reduction record and the compacted non-failed stores out. `survivor_cap` bounds the stores kept on the device."""

    def __init__(self, table, max_subproblems, survivor_cap=None):
        self._h = ctypes.c_void_p()
        self.table, self.nvars, self.max_n = table, table.nvars, max_subproblems
        self.survivor_cap = max_subproblems if survivor_cap is None else survivor_cap
        _check(_L.lpc_eps_create(table._h, max_subproblems, self.survivor_cap, ctypes.byref(self._h)))
        self.n = 0

    def set_rank(self, rank, world):
        _check(_L.lpc_eps_set_rank(self._h, rank, world))

    @property
    def payload(self):
        n = ctypes.c_int32(0)
        p = _L.lpc_eps_payload_device_ptr(self._h, ctypes.byref(n))
        return p, n.value

    @staticmethod
    def _ptr(a):
        return None if a is None else (a if isinstance(a, int) else a.ctypes.data)

    def solve_host(self, root, decision_vars, ids=None, first_id=0, n=None, objective_var=-1, flags=None, survivors=None,
                   survivor_index=None, max_survivors=None, **kw):
        """One call from host buffers. `root`, `flags`, `survivors`, `survivor_index` may be numpy arrays or raw (pinned)
        addresses; returns (EpsResult, number of survivor stores written)."""
        if not isinstance(root, int):
            root = np.ascontiguousarray(root, dtype=np.int32)
        d = np.ascontiguousarray(decision_vars, dtype=np.int32)
        if ids is not None and not isinstance(ids, int):
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            n = len(ids) if n is None else n
        assert n is not None
        self.n = n
        max_surv = 0
        if survivors is not None:
            max_surv = (self.survivor_cap if isinstance(survivors, int) else len(survivors)) if max_survivors is None else max_survivors
        o, r, nw = _opts(**kw), EpsResult(), ctypes.c_int32(0)
        _check(_L.lpc_eps_solve_host(self._h, self._ptr(root), d.ctypes.data, d.shape[0], self._ptr(ids), first_id, n,
                                     ctypes.byref(o), objective_var, self._ptr(flags), self._ptr(survivors),
                                     self._ptr(survivor_index), max_surv, ctypes.byref(nw), ctypes.byref(r)))
        return r, nw.value

    def upload(self, root, decision_vars, ids=None, first_id=0, n=None):
        root = np.ascontiguousarray(root, dtype=np.int32)
        d = np.ascontiguousarray(decision_vars, dtype=np.int32)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            n = len(ids) if n is None else n
        self.n = n
        _check(_L.lpc_eps_upload(self._h, root.ctypes.data, d.ctypes.data, d.shape[0], self._ptr(ids), first_id, n))

    def run_async(self, objective_var=-1, **kw):
        o = _opts(**kw)
        _check(_L.lpc_eps_run_async(self._h, ctypes.byref(o), objective_var))

    def collect(self):
        r = EpsResult()
        _check(_L.lpc_eps_collect(self._h, ctypes.byref(r)))
        return r

    def run(self, objective_var=-1, **kw):
        self.run_async(objective_var, **kw)
        return self.collect()

    def download(self, max_survivors=None):
        """(flags u8 [n], survivors int32 [k, nvars, 2], survivor_index int32 [k]) of the last collected run."""
        cap = self.survivor_cap if max_survivors is None else max_survivors
        flags = np.empty(max(self.n, 1), dtype=np.uint8)
        surv = np.empty((max(cap, 1), self.nvars, 2), dtype=np.int32)
        idx = np.empty(max(cap, 1), dtype=np.int32)
        nw = ctypes.c_int32(0)
        _check(_L.lpc_eps_download(self._h, flags.ctypes.data, surv.ctypes.data, idx.ctypes.data, cap, ctypes.byref(nw)))
        return flags[:self.n], surv[:nw.value], idx[:nw.value]

    def sweeps(self):
        out = np.empty(max(self.n, 1), dtype=np.int32)
        _check(_L.lpc_eps_sweeps(self._h, out.ctypes.data))
        return out[:self.n]

    def peer_disconnect(self):
        _check(_L.lpc_eps_peer_disconnect(self._h))

    # -- record exchange over peer memory (include/lpc.h: lpc_eps_peer_*) --
    PEER_HANDLE_BYTES = 64

    def peer_export(self):
        buf = ctypes.create_string_buffer(self.PEER_HANDLE_BYTES)
        _check(_L.lpc_eps_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, rank, world, handles):
        """`handles`: the ranks' exported handles in rank order (list of bytes)."""
        blob = b"".join(handles)
        assert len(blob) == world * self.PEER_HANDLE_BYTES
        _check(_L.lpc_eps_peer_connect(self._h, rank, world, blob))

    def close(self):
        if getattr(self, "_h", None):
            _L.lpc_eps_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class PIR:
    """Python mirror of `PIR<VStore<Interval<ZLB>>>` over the C-ABI (same member names as pir.hpp)."""

    name = "PIR"

    def __init__(self, nvars):
        self._n = 0
        self.store = Store(nvars)
        self._table = Table(None, nvars)

    # -- build step: PIR::deduce(const tell_type&) (pir.hpp:326-352) --
    def tell(self, records=(), domains=()):
        """records: iterable of (op, x, y, z); domains: iterable of (var, lb, ub) store tells."""
        changed = False
        for v, lb, ub in domains:
            changed |= self.store.embed(v, lb, ub)
        recs = np.asarray(list(records), dtype=np.int32).reshape(-1, 4)
        if len(recs):
            # incremental: only the appended records are sorted and merged in, only the changed part of the device image
            # is uploaded (lpc_table_append / lpc_table_finalize)
            self._table.append(recs)
            self._table.finalize(sort=True)
            self._n += len(recs)
            _check(_L.lpc_table_clamp_reified(self._table._h, self.store._h))
            changed = True
        return changed

    @property
    def table(self):
        return self._table

    def num_deductions(self):
        return self._n

    def load_deduce(self, i):
        return self.table.load(i)

    load_deductions = load_deduce   # spelling used by BASELINE.json's north_star

    def deduce(self, i):
        c = ctypes.c_int(0)
        _check(_L.lpc_deduce_one(self.table._h, self.store._h, i, ctypes.byref(c)))
        return bool(c.value)

    def ask(self, i):
        c = ctypes.c_int(0)
        _check(_L.lpc_ask_one(self.table._h, self.store._h, i, ctypes.byref(c)))
        return bool(c.value)

    def embed(self, var, lb, ub):
        return self.store.embed(var, lb, ub)

    def fixpoint(self, **kw):
        return fixpoint(self.table, self.store, **kw)

    def is_bot(self):
        return self.store.is_bot()

    def is_top(self):
        return self.store.is_top() and self.num_deductions() == 0

    def __getitem__(self, v):
        lb, ub = self.store.read(v, 1)[0]
        return int(lb), int(ub)

    project = __getitem__

    def vars(self):
        return self.store.nvars

    def snapshot(self):
        return self._n, self.store.read()

    def restore(self, snap):
        """pir.hpp:863-870: pops the records behind the snapshot's count from the back of the (sorted) table."""
        n, values = snap
        if n < self._n:
            self._table.truncate(n)
            self._table.finalize(sort=True)
            self._n = n
        self.store.write(values)

    def is_extractable(self):
        if self.is_bot():
            return False
        n = ctypes.c_int64(0)
        _check(_L.lpc_ask_all(self.table._h, self.store._h, ctypes.byref(n)))
        return n.value == self.num_deductions()

    def extract(self):
        return self.store.read()
