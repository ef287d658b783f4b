"""Seeded synthetic one-class data in the reference's shapes and text format.

The reference ships no dataset (shapes C1-C5); every benchmark and parity test
runs on data from this generator.  Output follows the reference's data contract
(ffm.cpp:80-183): one text line per user / item,

    train/test :  ``j1,j2,... fid:idx:val fid:idx:val ...``   (labels = item row numbers)
    item file  :  ``fid:idx:val ...``                         (line number = item id)

and the same arrays are available as CSR (``Side``/``Dataset``) so the GPU path can be fed
without a text round trip.  Popularity of items is Zipf(``zipf``) over a random permutation
of item ids; positives per user are Poisson(``pos_per_user``), de-duplicated and sorted inside
a row; every feature value is 1.0; field 0 on each side is id-like (one nnz per row, idx=row)
unless the spec says otherwise.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclasses.dataclass
class FieldSpec:
    D: int                 # number of distinct feature ids
    nnz_lo: int = 1        # features per row, inclusive range
    nnz_hi: int = 1
    id_like: bool = False  # idx = row % D, one nnz per row
    zipf: float = 0.0      # >0: skewed feature popularity


@dataclasses.dataclass
class Field:
    D: int
    rowptr: np.ndarray  # uint64 [rows+1]
    idx: np.ndarray     # uint32 [nnz]
    val: np.ndarray     # float64 [nnz]


@dataclasses.dataclass
class Side:
    rows: int
    fields: List[Field]

    @property
    def f(self) -> int:
        return len(self.fields)


@dataclasses.dataclass
class Labels:
    rows: int
    n_items: int           # U->n = max label + 1 (ffm.cpp:97)
    rowptr: np.ndarray     # uint64 [rows+1]
    idx: np.ndarray        # uint32 [nnz]


@dataclasses.dataclass
class Dataset:
    name: str
    users: Side
    items: Side
    train: Labels
    test_users: Optional[Side] = None
    test: Optional[Labels] = None
    meta: dict = dataclasses.field(default_factory=dict)

    @property
    def m(self) -> int:
        return self.users.rows

    @property
    def n(self) -> int:
        return self.items.rows


SHAPES = {
    # SURVEY.md §8 table.  (users, items, user fields, item fields, positives/user)
    "tiny": dict(m=60, n=40, fu=[FieldSpec(60, id_like=True), FieldSpec(7, 1, 3)],
                 fv=[FieldSpec(40, id_like=True), FieldSpec(5, 1, 2)], pos=4.0),
    "C1": dict(m=10_000, n=5_000, fu=[FieldSpec(10_000, id_like=True), FieldSpec(500, 1, 3)],
               fv=[FieldSpec(5_000, id_like=True), FieldSpec(200, 1, 3)], pos=8.0),
    "C2": dict(m=30_000, n=360_000, fu=[FieldSpec(30_000, id_like=True), FieldSpec(2_000, 1, 3)],
               fv=[FieldSpec(360_000, id_like=True), FieldSpec(5_000, 1, 3)], pos=300.0),
    "C3": dict(m=8_000_000, n=20_000,
               fu=[FieldSpec(3_000, 3, 3, zipf=1.1), FieldSpec(900_000, 3, 3, zipf=1.05)],
               fv=[FieldSpec(50_000, 3, 3), FieldSpec(8_000, 2, 2)], pos=1.0, min_pos=1),
    "C4": dict(m=4_000_000, n=50_000,
               fu=[FieldSpec(2_500_000, id_like=True), FieldSpec(1_200_000, 2, 2, zipf=1.05)],
               fv=[FieldSpec(40_000, 1, 1), FieldSpec(35_000, 1, 1), FieldSpec(45_000, 1, 1),
                   FieldSpec(60_000, 3, 3)], pos=1.3, min_pos=1),
    # C4 (KDD12-shaped) cut down for the oracle: same field structure (fu=2, fv=4 -> 8 cross pairs,
    # id-like user field, Zipf query field whose top feature passes the default
    # hot-replica threshold of 16 384 occurrences), ~1.3 positives per row
    "C4s": dict(m=70_000, n=2_000,
                fu=[FieldSpec(70_000, id_like=True), FieldSpec(5_000, 2, 2, zipf=1.05)],
                fv=[FieldSpec(1_600, 1, 1), FieldSpec(1_400, 1, 1), FieldSpec(1_800, 1, 1),
                    FieldSpec(2_400, 3, 3)], pos=1.3, min_pos=1),
    # C3 (Outbrain-shaped) cut down: no id field at all, Zipf user fields, 1 positive per row, k=16
    "C3s": dict(m=120_000, n=1_000,
                fu=[FieldSpec(300, 3, 3, zipf=1.1), FieldSpec(14_000, 3, 3, zipf=1.05)],
                fv=[FieldSpec(2_500, 3, 3), FieldSpec(400, 2, 2)], pos=1.0, min_pos=1),
}


def _zipf_cdf(n: int, s: float) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def _gen_field(rng: np.random.Generator, rows: int, spec: FieldSpec) -> Field:
    if spec.id_like:
        rowptr = np.arange(rows + 1, dtype=np.uint64)
        idx = (np.arange(rows, dtype=np.uint64) % spec.D).astype(np.uint32)
    else:
        cnt = rng.integers(spec.nnz_lo, spec.nnz_hi + 1, size=rows, dtype=np.int64)
        rowptr = np.zeros(rows + 1, dtype=np.uint64)
        rowptr[1:] = np.cumsum(cnt)
        nnz = int(rowptr[-1])
        if spec.zipf > 0:
            cdf = _zipf_cdf(spec.D, spec.zipf)
            perm = rng.permutation(spec.D)
            idx = perm[np.searchsorted(cdf, rng.random(nnz))].astype(np.uint32)
        else:
            idx = rng.integers(0, spec.D, size=nnz, dtype=np.int64).astype(np.uint32)
        # make sure the largest id is present so Ds[fid] == spec.D (ffm.cpp:221)
        if nnz:
            idx[-1] = spec.D - 1
    return Field(spec.D, rowptr, idx, np.ones(idx.shape[0], dtype=np.float64))


def _gen_labels(rng: np.random.Generator, rows: int, n_items: int, pos: float, zipf: float,
                perm: np.ndarray, min_pos: int = 0) -> Labels:
    cnt = np.maximum(rng.poisson(pos, size=rows), min_pos).astype(np.int64)
    cnt = np.minimum(cnt, n_items)
    cdf = _zipf_cdf(n_items, zipf)
    # Heavy users would lose most of their Zipf draws to duplicates: oversample, de-duplicate,
    # then keep a uniformly random `cnt` of each user's distinct items.
    over = 6 if pos > 20 else 1
    draws = cnt * over
    tot = int(draws.sum())
    items = perm[np.searchsorted(cdf, rng.random(tot))].astype(np.int64)
    users = np.repeat(np.arange(rows, dtype=np.int64), draws)
    key = np.unique(users * n_items + items)          # de-duplicate, sort by (user, item)
    if over > 1:
        users = key // n_items
        prio = rng.random(key.size)
        order = np.lexsort((prio, users))
        start = np.zeros(rows + 1, dtype=np.int64)
        start[1:] = np.cumsum(np.bincount(users, minlength=rows))
        rank_in_user = np.arange(key.size, dtype=np.int64) - start[users[order]]
        key = np.sort(key[order[rank_in_user < cnt[users[order]]]])
    users, items = key // n_items, key % n_items
    rowptr = np.zeros(rows + 1, dtype=np.uint64)
    rowptr[1:] = np.cumsum(np.bincount(users, minlength=rows))
    return Labels(rows, int(items.max()) + 1 if items.size else 0, rowptr, items.astype(np.uint32))


def generate(shape: str = "C1", seed: int = 1, k_hint: int = 0, scale: float = 1.0,
             test_rows: int = 0, cold_rows: int = 0, zipf: float = 1.3,
             pos_override: Optional[float] = None) -> Dataset:
    """Build one synthetic set.  ``scale`` shrinks users, items and the id-like fields together
    (used for the bounded CPU-baseline sample); ``test_rows`` adds a test split whose rows reuse
    the feature distribution of the users; ``cold_rows`` of them carry only out-of-vocabulary
    features (dropped by the reader, ffm.cpp:104-105 -> nnx == 0 -> ranked by popularity)."""
    cfg = SHAPES[shape]
    rng = np.random.default_rng(seed)
    m = max(8, int(round(cfg["m"] * scale)))
    n = max(8, int(round(cfg["n"] * scale)))

    def scaled(specs: Sequence[FieldSpec], rows: int) -> List[FieldSpec]:
        out = []
        for s in specs:
            D = rows if s.id_like else max(2, int(round(s.D * (scale if s.D > 5000 else 1.0))))
            out.append(dataclasses.replace(s, D=D))
        return out

    fu, fv = scaled(cfg["fu"], m), scaled(cfg["fv"], n)
    users = Side(m, [_gen_field(rng, m, s) for s in fu])
    items = Side(n, [_gen_field(rng, n, s) for s in fv])
    perm = rng.permutation(n)
    pos = cfg["pos"] if pos_override is None else pos_override
    train = _gen_labels(rng, m, n, pos, zipf, perm, cfg.get("min_pos", 0))
    ds = Dataset(shape, users, items, train,
                 meta=dict(shape=shape, seed=seed, scale=scale, zipf=zipf, pos_per_user=pos,
                           m=m, n=n, nnz_y=int(train.idx.size)))
    if test_rows:
        tfields = []
        for s, trained in zip(fu, users.fields):
            fld = _gen_field(rng, test_rows, s)
            # never exceed the training vocabulary: the reader drops idx >= Ds[fid]
            fld.idx = np.minimum(fld.idx, np.uint32(trained.D - 1))
            tfields.append(fld)
        if cold_rows:
            # empty the last `cold_rows` rows (equivalent to all-OOV rows after the reader's filter)
            for fld in tfields:
                cut = int(fld.rowptr[test_rows - cold_rows])
                fld.idx, fld.val = fld.idx[:cut], fld.val[:cut]
                fld.rowptr[test_rows - cold_rows:] = cut
        ds.test_users = Side(test_rows, tfields)
        ds.test = _gen_labels(rng, test_rows, n, max(2.0, min(pos, 10.0)), zipf, perm, 1)
        ds.meta.update(test_rows=test_rows, cold_rows=cold_rows)
    return ds


# ---------------------------------------------------------------------------------------
# text format (what the reference's reader and our host reader parse)
# ---------------------------------------------------------------------------------------
def _feature_tokens(side: Side, row: int, oov: Optional[Sequence[int]] = None) -> List[str]:
    toks = []
    for fid, fld in enumerate(side.fields):
        a, b = int(fld.rowptr[row]), int(fld.rowptr[row + 1])
        for p in range(a, b):
            v = fld.val[p]
            toks.append(f"{fid}:{int(fld.idx[p])}:{int(v) if v == int(v) else repr(float(v))}")
        if oov is not None and a == b:
            toks.append(f"{fid}:{int(oov[fid])}:1")
    return toks


def write_text(ds: Dataset, out_dir: str) -> Tuple[str, str, Optional[str]]:
    """Write item / train / (test) files; returns their paths (item first, the CLI's order)."""
    os.makedirs(out_dir, exist_ok=True)
    item_p = os.path.join(out_dir, f"{ds.name}.item")
    tr_p = os.path.join(out_dir, f"{ds.name}.tr")
    te_p = os.path.join(out_dir, f"{ds.name}.te") if ds.test is not None else None
    with open(item_p, "w") as fh:
        for j in range(ds.n):
            fh.write(" ".join(_feature_tokens(ds.items, j)) + "\n")

    def write_labelled(path: str, side: Side, lab: Labels, oov):
        rp, idx = lab.rowptr, lab.idx
        with open(path, "w") as fh:
            for i in range(side.rows):
                labs = ",".join(map(str, idx[int(rp[i]):int(rp[i + 1])].tolist()))
                fh.write(labs + " " + " ".join(_feature_tokens(side, i, oov)) + "\n")

    write_labelled(tr_p, ds.users, ds.train, None)
    if te_p:
        # cold rows get one out-of-vocabulary feature per field so that the reader's filter,
        # not an empty line, is what produces nnx == 0
        oov = [fld.D + 7 for fld in ds.users.fields]
        write_labelled(te_p, ds.test_users, ds.test, oov)
    return item_p, tr_p, te_p


def csc_of(lab: Labels, n_rows_other: int) -> Tuple[np.ndarray, np.ndarray]:
    """CSC of Omega sorted by (item, user) -- the reference's transY (ffm.cpp:259-294).
    Labels >= n_rows_other are skipped (ffm.cpp:267-268)."""
    users = np.repeat(np.arange(lab.rows, dtype=np.int64), np.diff(lab.rowptr.astype(np.int64)))
    items = lab.idx.astype(np.int64)
    keep = items < n_rows_other
    users, items = users[keep], items[keep]
    order = np.lexsort((users, items))
    colptr = np.zeros(n_rows_other + 1, dtype=np.uint64)
    colptr[1:] = np.cumsum(np.bincount(items, minlength=n_rows_other))
    return colptr, users[order].astype(np.uint32)


# ---------------------------------------------------------------------------------------
# weak-scaling sets: the item side is fixed, users come in independent seeded blocks of the
# shape's m rows each (one block per GPU), so ranks can generate their blocks in parallel
# ---------------------------------------------------------------------------------------
def _item_side(shape: str, seed: int):
    cfg = SHAPES[shape]
    rng = np.random.default_rng([seed, 31337])
    n = cfg["n"]
    items = Side(n, [_gen_field(rng, n, dataclasses.replace(s, D=n if s.id_like else s.D)) for s in cfg["fv"]])
    return items, rng.permutation(n)


def user_block(shape: str, seed: int, block: int, n_blocks: int, zipf: float = 1.3):
    """Users [block*m, (block+1)*m) of an n_blocks*m-user set: (Side, Labels) with global ids."""
    cfg = SHAPES[shape]
    m, n = cfg["m"], cfg["n"]
    _, perm = _item_side(shape, seed)
    rng = np.random.default_rng([seed, 7919, block])
    fields = []
    for s in cfg["fu"]:
        if s.id_like:
            fld = Field(m * n_blocks, np.arange(m + 1, dtype=np.uint64),
                        (np.arange(m, dtype=np.uint64) + block * m).astype(np.uint32), np.ones(m))
        else:
            fld = _gen_field(rng, m, s)
        fields.append(fld)
    labels = _gen_labels(rng, m, n, cfg["pos"], zipf, perm, cfg.get("min_pos", 0))
    return Side(m, fields), labels


def assemble_blocks(shape: str, seed: int, blocks, test_rows: int = 0, zipf: float = 1.3) -> Dataset:
    """Concatenate user blocks (in block order) over the shared item side."""
    cfg = SHAPES[shape]
    items, perm = _item_side(shape, seed)
    n = cfg["n"]
    sides, labs = [b[0] for b in blocks], [b[1] for b in blocks]
    m = sum(s.rows for s in sides)

    def cat_ptr(ptrs):
        out, base = [np.zeros(1, dtype=np.uint64)], 0
        for p in ptrs:
            out.append(p[1:].astype(np.uint64) + np.uint64(base))
            base += int(p[-1])
        return np.concatenate(out)

    fields = []
    for fi in range(len(cfg["fu"])):
        fl = [s.fields[fi] for s in sides]
        fields.append(Field(max(f.D for f in fl), cat_ptr([f.rowptr for f in fl]),
                            np.concatenate([f.idx for f in fl]), np.concatenate([f.val for f in fl])))
    users = Side(m, fields)
    train = Labels(m, max(l.n_items for l in labs), cat_ptr([l.rowptr for l in labs]),
                   np.concatenate([l.idx for l in labs]))
    ds = Dataset(shape, users, items, train,
                 meta=dict(shape=shape, seed=seed, blocks=len(blocks), zipf=zipf, m=m, n=n,
                           nnz_y=int(train.idx.size)))
    if test_rows:
        rng = np.random.default_rng([seed, 104729])
        tfields = []
        for s, trained in zip(cfg["fu"], users.fields):
            spec = dataclasses.replace(s, D=trained.D)
            if s.id_like:
                ids = rng.integers(0, trained.D, size=test_rows, dtype=np.int64).astype(np.uint32)
                fld = Field(trained.D, np.arange(test_rows + 1, dtype=np.uint64), ids, np.ones(test_rows))
            else:
                fld = _gen_field(rng, test_rows, spec)
                fld.idx = np.minimum(fld.idx, np.uint32(trained.D - 1))
            tfields.append(fld)
        ds.test_users = Side(test_rows, tfields)
        ds.test = _gen_labels(rng, test_rows, n, 10.0, zipf, perm, 1)
        ds.meta.update(test_rows=test_rows)
    return ds
