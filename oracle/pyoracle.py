"""ctypes front-end of the plain-C oracle (oracle/ocffm_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

SIDE_U, SIDE_V, SIDE_T = 0, 1, 2
TOPK = (5, 10, 20, 40, 80)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and, when /root/reference is present, oracle/_ref)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ocffm_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u64p, u32p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double)
        L.oc_create.restype = C.c_void_p
        L.oc_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_double,
                                C.c_double, C.c_double, C.c_int, C.c_int]
        L.oc_destroy.argtypes = [C.c_void_p]
        L.oc_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, u64p, u32p, f64p]
        L.oc_set_labels.argtypes = [C.c_void_p, C.c_int, C.c_uint64, u64p, u32p]
        L.oc_set_test_nnx.argtypes = [C.c_void_p, u64p]
        L.oc_init_model_rng.argtypes = [C.c_void_p]
        L.oc_block_rows.restype = C.c_uint64
        L.oc_block_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oc_set_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64p]
        L.oc_get_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64p]
        L.oc_init_state.argtypes = [C.c_void_p]
        L.oc_get_vec.restype = C.c_uint64
        L.oc_get_vec.argtypes = [C.c_void_p, C.c_char_p, f64p]
        L.oc_get_embed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64p]
        L.oc_get_csc.argtypes = [C.c_void_p, u64p, u32p]
        L.oc_grad.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64p]
        L.oc_hess_vec.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64p, f64p]
        L.oc_cg.restype = C.c_int
        L.oc_cg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64p, f64p]
        L.oc_solve_block.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oc_one_epoch.argtypes = [C.c_void_p]
        L.oc_cg_iters_total.restype = C.c_uint64
        L.oc_cg_iters_total.argtypes = [C.c_void_p]
        L.oc_func.restype = C.c_double
        L.oc_func.argtypes = [C.c_void_p]
        L.oc_validate.argtypes = [C.c_void_p, f64p, f64p, f64p, u32p, f64p]
        L.oc_ndcg_at.restype = C.c_double
        L.oc_ndcg_at.argtypes = [u32p, C.c_uint32, u32p, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One reference-semantics problem held by the C oracle."""

    def __init__(self, ds, k: int, lam: float, omega: float, r: float = -1.0,
                 self_side: bool = True, freq: bool = False):
        L = lib()
        self.L, self.ds, self.k = L, ds, k
        self.fu, self.fv = ds.users.f, ds.items.f
        self.f = self.fu + self.fv
        self.m, self.n = ds.users.rows, ds.items.rows
        self.self_side = self_side
        self.h = L.oc_create(self.fu, self.fv, self.m, self.n, k, lam, omega, r,
                             int(self_side), int(freq))
        self._keep = []
        for side, s in ((SIDE_U, ds.users), (SIDE_V, ds.items)):
            for fi, fld in enumerate(s.fields):
                self._set_field(side, fi, s.rows, fld)
        self._set_labels(SIDE_U, ds.train)
        self.mt = 0
        if getattr(ds, "test", None) is not None:
            for fi, fld in enumerate(ds.test_users.fields):
                self._set_field(SIDE_T, fi, ds.test_users.rows, fld)
            self._set_labels(SIDE_T, ds.test)
            self.mt = ds.test_users.rows

    def _set_field(self, side, fi, rows, fld):
        rp, ix, vl = _u64(fld.rowptr), _u32(fld.idx), _f64(fld.val)
        self.L.oc_set_field(self.h, side, fi, rows, fld.D, _p(rp, C.c_uint64), _p(ix, C.c_uint32),
                            _p(vl, C.c_double))

    def _set_labels(self, which, lab):
        rp, ix = _u64(lab.rowptr), _u32(lab.idx)
        self.L.oc_set_labels(self.h, which, lab.rows, _p(rp, C.c_uint64), _p(ix, C.c_uint32))

    def close(self):
        if self.h:
            self.L.oc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- model ---------------------------------------------------------------------------
    def blocks(self):
        for f1 in range(self.f):
            for f2 in range(f1, self.f):
                if self.self_side or (f1 < self.fu <= f2):
                    yield f1, f2

    def is_side(self, f1, f2):
        return (f1 < self.fu and f2 < self.fu) or (f1 >= self.fu and f2 >= self.fu)

    def init_model_rng(self):
        self.L.oc_init_model_rng(self.h)

    def block_rows(self, f1, f2, which):
        return int(self.L.oc_block_rows(self.h, f1, f2, ord(which)))

    def set_block(self, f1, f2, which, data):
        d = _f64(data).reshape(-1)
        assert d.size == self.block_rows(f1, f2, which) * self.k
        self.L.oc_set_block(self.h, f1, f2, ord(which), _p(d, C.c_double))

    def get_block(self, f1, f2, which):
        out = np.empty((self.block_rows(f1, f2, which), self.k), dtype=np.float64)
        self.L.oc_get_block(self.h, f1, f2, ord(which), _p(out, C.c_double))
        return out

    def init_state(self):
        self.L.oc_init_state(self.h)

    def vec(self, name: str) -> np.ndarray:
        n = int(self.L.oc_get_vec(self.h, name.encode(), None))
        out = np.empty(n, dtype=np.float64)
        self.L.oc_get_vec(self.h, name.encode(), _p(out, C.c_double))
        return out

    def embed(self, f1, f2, which):
        rows = (self.m if f1 < self.fu else self.n) if which == "P" else (self.m if f2 < self.fu else self.n)
        out = np.empty((rows, self.k), dtype=np.float64)
        self.L.oc_get_embed(self.h, f1, f2, ord(which), _p(out, C.c_double))
        return out

    def csc(self):
        nnz = int(self.L.oc_get_vec(self.h, b"ytilde_csc", None))
        colptr = np.empty(self.n + 1, dtype=np.uint64)
        rowidx = np.empty(nnz, dtype=np.uint32)
        self.L.oc_get_csc(self.h, _p(colptr, C.c_uint64), _p(rowidx, C.c_uint32))
        return colptr, rowidx

    # -- solver phases -------------------------------------------------------------------
    def grad(self, f1, f2, which):
        G = np.empty((self.block_rows(f1, f2, which), self.k), dtype=np.float64)
        self.L.oc_grad(self.h, f1, f2, ord(which), _p(G, C.c_double))
        return G

    def hess_vec(self, f1, f2, which, V):
        V = _f64(V)
        Hv = np.empty_like(V)
        self.L.oc_hess_vec(self.h, f1, f2, ord(which), _p(V, C.c_double), _p(Hv, C.c_double))
        return Hv

    def cg(self, f1, f2, which, G):
        G = _f64(G)
        S = np.zeros_like(G)
        it = self.L.oc_cg(self.h, f1, f2, ord(which), _p(G, C.c_double), _p(S, C.c_double))
        return S, int(it)

    def solve_block(self, f1, f2):
        self.L.oc_solve_block(self.h, f1, f2)

    def one_epoch(self):
        self.L.oc_one_epoch(self.h)

    def cg_iters_total(self) -> int:
        return int(self.L.oc_cg_iters_total(self.h))

    def func(self) -> float:
        return float(self.L.oc_func(self.h))

    def validate(self, want_topk=True, want_scores=False) -> Dict[str, np.ndarray]:
        prec, ndcg = np.zeros(5), np.zeros(5)
        ploss = C.c_double(0)
        topk = np.zeros((self.mt, 80), dtype=np.uint32) if want_topk else None
        Z = np.zeros((self.mt, self.n), dtype=np.float64) if want_scores else None
        self.L.oc_validate(self.h, _p(prec, C.c_double), _p(ndcg, C.c_double), C.byref(ploss),
                           _p(topk, C.c_uint32) if want_topk else None,
                           _p(Z, C.c_double) if want_scores else None)
        return dict(prec=prec, ndcg=ndcg, ploss=ploss.value, topk=topk, Z=Z)


def ndcg_at(ranking, labels, K: int) -> float:
    r, l = _u32(ranking), _u32(labels)
    return float(lib().oc_ndcg_at(_p(r, C.c_uint32), r.size, _p(l, C.c_uint32), l.size, K))


# -----------------------------------------------------------------------------------------
# dump container written by oracle/ref_harness.cpp
# -----------------------------------------------------------------------------------------
def load_ocfd(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as fh:
        assert fh.readline() == b"OCFD1\n", "not an OCFD1 dump"
        while True:
            hdr = fh.readline()
            if not hdr:
                break
            parts = hdr.split()
            name, dt, nd = parts[0].decode(), parts[1].decode(), int(parts[2])
            dims = [int(x) for x in parts[3:3 + nd]]
            cnt = int(np.prod(dims)) if dims else 1
            out[name] = np.frombuffer(fh.read(cnt * np.dtype(dt).itemsize), dtype=dt).reshape(dims).copy()
    return out


def dataset_from_dump(d: Dict[str, np.ndarray], name: str = "golden"):
    """Rebuild a synth.Dataset-shaped object from the arrays the reference's reader produced."""
    import importlib.util
    import sys
    if "ocffm_synth" not in sys.modules:
        spec = importlib.util.spec_from_file_location(
            "ocffm_synth", os.path.join(_HERE, "..", "one-class-ffm_b200", "synth.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ocffm_synth"] = mod
        spec.loader.exec_module(mod)
    S = sys.modules["ocffm_synth"]

    def side(pfx):
        rows, f = int(d[pfx + ".hdr"][0]), int(d[pfx + ".hdr"][2])
        flds = [S.Field(int(d[pfx + ".Ds"][fi]), d[f"{pfx}.X{fi}.rowptr"], d[f"{pfx}.X{fi}.idx"].astype(np.uint32),
                        d[f"{pfx}.X{fi}.val"]) for fi in range(f)]
        return S.Side(rows, flds)

    def labels(pfx):
        rows, n = int(d[pfx + ".hdr"][0]), int(d[pfx + ".hdr"][1])
        return S.Labels(rows, n, d[pfx + ".Y.rowptr"], d[pfx + ".Y.idx"].astype(np.uint32))

    ds = S.Dataset(name, side("U"), side("V"), labels("U"))
    if "T.hdr" in d:
        ts = side("T")
        # the test reader keeps the TRAINING vocabulary size (ffm.cpp:104-105)
        for fld, tr in zip(ts.fields, ds.users.fields):
            fld.D = tr.D
        ds.test_users, ds.test = ts, labels("T")
    return ds
