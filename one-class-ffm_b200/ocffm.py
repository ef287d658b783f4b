"""ctypes binding of libocffm_cuda.so (include/ocffm.h) used by tests/, bench.py and smoke().

This is a caller of the C ABI, nothing more: every method maps to one `ocffm_*` entry point and
there is no Python or CPU implementation of any of them.  If the library is missing or no CUDA
device is present, construction fails loudly (`OcffmError`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Iterator, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCFFM_LIB", os.path.join(_HERE, "libocffm_cuda.so"))   # OCFFM_LIB: tuning builds only
F32, F64 = 0, 1
SIDE_U, SIDE_V, SIDE_T = 0, 1, 2
TOPK = (5, 10, 20, 40, 80)
_lib = None


class OcffmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ocffm error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("omega", C.c_double), ("r", C.c_double), ("k", C.c_uint32),
                ("self_side", C.c_int32), ("freq", C.c_int32), ("dtype", C.c_int32), ("device", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("cg_iters", C.c_uint64), ("nnz_traversed", C.c_uint64),
                ("algo_bytes", C.c_uint64), ("ms_side_grad", C.c_double), ("ms_side_cg", C.c_double),
                ("ms_side_update", C.c_double), ("ms_cross_grad", C.c_double), ("ms_cross_cg", C.c_double),
                ("ms_cross_update", C.c_double), ("hv_launches", C.c_uint64), ("hv_algo_bytes", C.c_uint64),
                ("hv_ms", C.c_double), ("omega_device_bytes", C.c_uint64), ("row_gram_bytes", C.c_uint64),
                ("row_gram_builds", C.c_uint64), ("cg_kernel_ms", C.c_double), ("cg_kernel_algo_bytes", C.c_uint64),
                ("cg_kernel_launches", C.c_uint64), ("cg_kernel_iters", C.c_uint64)]


EXPORTS = [
    "ocffm_abi_version", "ocffm_last_error", "ocffm_device_count", "ocffm_create", "ocffm_destroy",
    "ocffm_comm_unique_id", "ocffm_shard_range", "ocffm_comm_init", "ocffm_set_field", "ocffm_set_labels",
    "ocffm_set_test_labels", "ocffm_set_hyper", "ocffm_init_model", "ocffm_mirror_block", "ocffm_set_block", "ocffm_get_block", "ocffm_init_state",
    "ocffm_solve_block", "ocffm_one_epoch", "ocffm_grad", "ocffm_hess_vec", "ocffm_cg",
    "ocffm_objective", "ocffm_validate", "ocffm_get_vec", "ocffm_get_embed", "ocffm_get_csc",
    "ocffm_get_stats", "ocffm_reset_stats", "ocffm_synchronize", "ocffm_stream",
]


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-j4", "-C", src_dir])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OcffmError(-2, f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no fallback)")
        L = C.CDLL(LIB_PATH)
        u64p, u32p, f64p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.c_void_p
        L.ocffm_last_error.restype = C.c_char_p
        L.ocffm_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]
        L.ocffm_destroy.argtypes = [vp]
        L.ocffm_comm_unique_id.argtypes = [vp]
        L.ocffm_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ocffm_shard_range.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p, u64p]
        L.ocffm_set_field.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, u64p, u32p, f64p]
        L.ocffm_set_labels.argtypes = [vp, C.c_uint64, u64p, u32p, u64p, u32p, C.c_uint64, f64p]
        L.ocffm_set_test_labels.argtypes = [vp, C.c_uint64, u64p, u32p, u64p]
        L.ocffm_set_block.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, C.c_uint64]
        L.ocffm_get_block.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, C.c_uint64]
        L.ocffm_init_state.argtypes = [vp]
        L.ocffm_init_model.argtypes = [vp, C.c_uint64]
        L.ocffm_mirror_block.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, C.c_uint64]
        L.ocffm_set_hyper.argtypes = [vp, C.c_double, C.c_double, C.c_double]
        L.ocffm_solve_block.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.ocffm_one_epoch.argtypes = [vp]
        L.ocffm_grad.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, C.c_uint64]
        L.ocffm_hess_vec.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, f64p, C.c_uint64]
        L.ocffm_cg.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, f64p, C.c_uint64, C.POINTER(C.c_int32)]
        L.ocffm_objective.argtypes = [vp, f64p]
        L.ocffm_validate.argtypes = [vp, f64p, f64p, f64p, u32p]
        L.ocffm_get_vec.argtypes = [vp, C.c_char_p, f64p, u64p]
        L.ocffm_get_embed.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, f64p, C.c_uint64]
        L.ocffm_get_csc.argtypes = [vp, u64p, u32p]
        L.ocffm_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.ocffm_reset_stats.argtypes = [vp]
        L.ocffm_synchronize.argtypes = [vp]
        L.ocffm_stream.argtypes = [vp, C.POINTER(vp)]
        _lib = L
    return _lib


def device_count() -> int:
    return int(lib().ocffm_device_count())


def _p(a: Optional[np.ndarray], ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def shard_range(rows: int, nranks: int, rank: int) -> Tuple[int, int]:
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    rc = lib().ocffm_shard_range(rows, nranks, rank, C.byref(lo), C.byref(hi))
    if rc:
        raise OcffmError(rc, lib().ocffm_last_error().decode())
    return int(lo.value), int(hi.value)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().ocffm_comm_unique_id(buf)
    if rc:
        raise OcffmError(rc, lib().ocffm_last_error().decode())
    return buf.raw


class Problem:
    """Python-side mirror of the reference's ImpProblem (ffm.h:82-151) over the C ABI."""

    def __init__(self, ds, k: int, lam: float, omega: float, r: float = -1.0, self_side: bool = True,
                 freq: bool = False, dtype: int = F32, device: int = -1,
                 comm: Optional[Tuple[int, int, bytes]] = None, upload: bool = True):
        L = lib()
        self.L, self.ds, self.k = L, ds, k
        self.fu, self.fv = ds.users.f, ds.items.f
        self.f = self.fu + self.fv
        self.m, self.n = ds.users.rows, ds.items.rows
        self.self_side, self.dtype = self_side, dtype
        self.mt = 0
        prm = Params(lam, omega, r, k, int(self_side), int(freq), dtype, device)
        h = C.c_void_p()
        self._ck(L.ocffm_create(C.byref(h), C.byref(prm), self.fu, self.fv, self.m, self.n))
        self.h = h
        if comm is not None:
            nranks, rank, uid = comm
            self._ck(L.ocffm_comm_init(self.h, nranks, rank, C.c_char_p(uid)))
        if upload:
            self.upload(ds)

    def _ck(self, rc: int):
        if rc != 0:
            raise OcffmError(rc, self.L.ocffm_last_error().decode())

    def upload(self, ds):
        for side, s in ((SIDE_U, ds.users), (SIDE_V, ds.items)):
            for fi, fld in enumerate(s.fields):
                self.set_field(side, fi, s.rows, fld)
        self.set_labels(ds.train)
        if getattr(ds, "test", None) is not None:
            for fi, (fld, tr) in enumerate(zip(ds.test_users.fields, ds.users.fields)):
                self.set_field(SIDE_T, fi, ds.test_users.rows, fld, D=tr.D)
            self.set_test_labels(ds.test)

    def set_field(self, side, fi, rows, fld, D=None):
        rp, ix, vl = _u64(fld.rowptr), _u32(fld.idx), _f64(fld.val)
        self._ck(self.L.ocffm_set_field(self.h, side, fi, rows, fld.D if D is None else D,
                                        _p(rp, C.c_uint64), _p(ix, C.c_uint32), _p(vl, C.c_double)))

    def set_labels(self, lab, csc=None):
        rp, ix = _u64(lab.rowptr), _u32(lab.idx)
        cp = ri = None
        if csc is not None:
            cp, ri = _u64(csc[0]), _u32(csc[1])
        self._ck(self.L.ocffm_set_labels(self.h, lab.rows, _p(rp, C.c_uint64), _p(ix, C.c_uint32),
                                         _p(cp, C.c_uint64), _p(ri, C.c_uint32), 0, None))

    def set_test_labels(self, lab, nnx=None):
        rp, ix = _u64(lab.rowptr), _u32(lab.idx)
        nx = None if nnx is None else _u64(nnx)
        self._ck(self.L.ocffm_set_test_labels(self.h, lab.rows, _p(rp, C.c_uint64), _p(ix, C.c_uint32),
                                              _p(nx, C.c_uint64)))
        self.mt = lab.rows

    def close(self):
        if getattr(self, "h", None):
            self.L.ocffm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- model ---------------------------------------------------------------------------
    def blocks(self) -> Iterator[Tuple[int, int]]:
        for f1 in range(self.f):
            for f2 in range(f1, self.f):
                if self.self_side or (f1 < self.fu <= f2):
                    yield f1, f2

    def _field(self, fg):
        return self.ds.users.fields[fg] if fg < self.fu else self.ds.items.fields[fg - self.fu]

    def block_rows(self, f1, f2, which) -> int:
        return self._field(f1 if which == "W" else f2).D

    def set_block(self, f1, f2, which, data):
        d = _f64(data).reshape(-1)   # no copy for a contiguous float64 array (keeps pinned memory pinned)
        rows = self.block_rows(f1, f2, which)
        assert d.size == rows * self.k
        self._ck(self.L.ocffm_set_block(self.h, f1, f2, ord(which), _p(d, C.c_double), rows))

    def get_block(self, f1, f2, which, out: Optional[np.ndarray] = None) -> np.ndarray:
        """`out`: an existing float64 [rows, k] array to fill (e.g. pinned memory for DMA)."""
        rows = self.block_rows(f1, f2, which)
        if out is None:
            out = np.empty((rows, self.k), dtype=np.float64)
        assert out.dtype == np.float64 and out.size == rows * self.k and out.flags["C_CONTIGUOUS"]
        self._ck(self.L.ocffm_get_block(self.h, f1, f2, ord(which), _p(out, C.c_double), rows))
        return out

    def init_model(self, seed: int = 1):
        """Random blocks with the reference's scale U(-0.1/sqrt(k), 0.1/sqrt(k)) (ffm.cpp:71-78);
        numpy's generator, NOT the reference's libstdc++ stream (the C++ host does that)."""
        rng = np.random.default_rng(seed)
        s = 0.1 / np.sqrt(self.k)
        out = {}
        for f1, f2 in self.blocks():
            for which in "WH":
                out[(f1, f2, which)] = rng.uniform(-s, s, size=(self.block_rows(f1, f2, which), self.k))
                self.set_block(f1, f2, which, out[(f1, f2, which)])
        return out

    def mirror_block(self, f1, f2, which, pinned: Optional[np.ndarray]):
        """Register (or, with None, drop) a pinned float64 [rows, k] host mirror of a block half:
        one_epoch() keeps it current, streaming the block out while the rest of the iteration runs."""
        rows = self.block_rows(f1, f2, which)
        if pinned is not None:
            assert pinned.dtype == np.float64 and pinned.size == rows * self.k and pinned.flags["C_CONTIGUOUS"]
        self._ck(self.L.ocffm_mirror_block(self.h, f1, f2, ord(which), _p(pinned, C.c_double), rows))

    def init_model_device(self, seed: int = 1):
        """Counter-based init on the GPU (ocffm_init_model): same distribution as init_mat, no PCIe."""
        self._ck(self.L.ocffm_init_model(self.h, seed))

    def init_state(self):
        self._ck(self.L.ocffm_init_state(self.h))

    def set_hyper(self, lam: float, omega: float, r: float = -1.0):
        """New (lambda, omega, r) on the resident data; set blocks + init_state() afterwards."""
        self._ck(self.L.ocffm_set_hyper(self.h, lam, omega, r))

    def vec(self, name: str) -> np.ndarray:
        cnt = C.c_uint64(0)
        self._ck(self.L.ocffm_get_vec(self.h, name.encode(), None, C.byref(cnt)))
        out = np.empty(cnt.value, dtype=np.float64)
        self._ck(self.L.ocffm_get_vec(self.h, name.encode(), _p(out, C.c_double), C.byref(cnt)))
        return out

    def embed(self, f1, f2, which) -> np.ndarray:
        rows = (self.m if f1 < self.fu else self.n) if which == "P" else (self.m if f2 < self.fu else self.n)
        out = np.empty((rows, self.k), dtype=np.float64)
        self._ck(self.L.ocffm_get_embed(self.h, f1, f2, ord(which), _p(out, C.c_double), rows))
        return out

    def csc(self):
        nnz = int(self.ds.train.idx.size)
        colptr = np.empty(self.n + 1, dtype=np.uint64)
        rowidx = np.empty(nnz, dtype=np.uint32)
        self._ck(self.L.ocffm_get_csc(self.h, _p(colptr, C.c_uint64), _p(rowidx, C.c_uint32)))
        return colptr, rowidx

    # -- solver --------------------------------------------------------------------------
    def grad(self, f1, f2, which) -> np.ndarray:
        rows = self.block_rows(f1, f2, which)
        G = np.empty((rows, self.k), dtype=np.float64)
        self._ck(self.L.ocffm_grad(self.h, f1, f2, ord(which), _p(G, C.c_double), rows))
        return G

    def hess_vec(self, f1, f2, which, V) -> np.ndarray:
        V = _f64(V)
        Hv = np.empty_like(V)
        rows = self.block_rows(f1, f2, which)
        self._ck(self.L.ocffm_hess_vec(self.h, f1, f2, ord(which), _p(V, C.c_double), _p(Hv, C.c_double), rows))
        return Hv

    def cg(self, f1, f2, which, G):
        G = _f64(G)
        S = np.zeros_like(G)
        it = C.c_int32(0)
        rows = self.block_rows(f1, f2, which)
        self._ck(self.L.ocffm_cg(self.h, f1, f2, ord(which), _p(G, C.c_double), _p(S, C.c_double), rows,
                                 C.byref(it)))
        return S, int(it.value)

    def solve_block(self, f1, f2):
        self._ck(self.L.ocffm_solve_block(self.h, f1, f2))

    def one_epoch(self):
        self._ck(self.L.ocffm_one_epoch(self.h))

    def objective(self) -> float:
        v = C.c_double(0)
        self._ck(self.L.ocffm_objective(self.h, C.byref(v)))
        return float(v.value)

    def validate(self, want_topk: bool = True) -> Dict[str, np.ndarray]:
        prec, ndcg = np.zeros(5), np.zeros(5)
        ploss = C.c_double(0)
        topk = np.zeros((self.mt, 80), dtype=np.uint32) if want_topk else None
        self._ck(self.L.ocffm_validate(self.h, _p(prec, C.c_double), _p(ndcg, C.c_double), C.byref(ploss),
                                       _p(topk, C.c_uint32)))
        return dict(prec=prec, ndcg=ndcg, ploss=ploss.value, topk=topk)

    # -- instrumentation -----------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        self._ck(self.L.ocffm_get_stats(self.h, C.byref(s)))
        return s

    def reset_stats(self):
        self._ck(self.L.ocffm_reset_stats(self.h))

    def synchronize(self):
        self._ck(self.L.ocffm_synchronize(self.h))

    def stream(self) -> int:
        s = C.c_void_p()
        self._ck(self.L.ocffm_stream(self.h, C.byref(s)))
        return int(s.value or 0)
