"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU oracle (oracle/libppp_oracle.so = cephes_port.c + ppp_oracle.c), the
C restatement of the reference hot path `PhyloProf::Algorithm::ppp::get_match_score`
(/root/reference/lib/PhyloProf/Algorithm/ppp.pm:130-245) and of Math::Cephes 0.47 `bdtrc`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (prophylo_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libppp_oracle.so")


class OracleResult(ctypes.Structure):
    _fields_ = [
        ("score", ctypes.c_double),
        ("yes", ctypes.c_int32),
        ("number", ctypes.c_int32),
        ("depth", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
        ("p", ctypes.c_double),
    ]


RESULT_DTYPE = np.dtype(
    [("score", "<f8"), ("yes", "<i4"), ("number", "<i4"), ("depth", "<i4"), ("pad_", "<i4"), ("p", "<f8")]
)


def build(force: bool = False) -> str:
    """Compile the oracle's C restatement (gcc only; no reference tree needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("cephes_port.c", "ppp_oracle.c", "cephes_port.h", "ppp_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libppp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        D, I, P = ctypes.c_double, ctypes.c_int, ctypes.c_void_p
        L.cp_bdtrc.restype = D
        L.cp_bdtrc.argtypes = [I, I, D]
        L.cp_incbet.restype = D
        L.cp_incbet.argtypes = [D, D, D]
        L.cp_bdtrc_batch.restype = None
        L.cp_bdtrc_batch.argtypes = [P, P, P, P, ctypes.c_long]
        L.ppp_oracle_list.restype = ctypes.c_long
        L.ppp_oracle_list.argtypes = [P, P, ctypes.c_int32, ctypes.c_int32, P, ctypes.c_int32, D, I, P]
        L.ppp_oracle_bits.restype = ctypes.c_long
        L.ppp_oracle_bits.argtypes = [P, P, P, ctypes.c_int32, D, P]
        L.ppp_oracle_bits_batch.restype = ctypes.c_long
        L.ppp_oracle_bits_batch.argtypes = [P, P, P, ctypes.c_int64, P, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, P, I]
        L.ppp_oracle_bits_batch_memo.restype = ctypes.c_long
        L.ppp_oracle_bits_batch_memo.argtypes = L.ppp_oracle_bits_batch.argtypes
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def bdtrc(k: int, n: int, p: float) -> float:
    """Math::Cephes::bdtrc(k, n, p) -- call site ppp.pm:225."""
    return lib().cp_bdtrc(int(k), int(n), float(p))


def bdtrc_batch(k, n, p) -> np.ndarray:
    k = np.ascontiguousarray(k, dtype=np.int32)
    n = np.ascontiguousarray(n, dtype=np.int32)
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty(k.shape, dtype=np.float64)
    lib().cp_bdtrc_batch(_ptr(k), _ptr(n), _ptr(p), _ptr(out), k.size)
    return out


def score_list(present, isyes, max_yes: int, target_ids, p: float, dup: bool = False):
    """get_match_score over integer taxon ids. Returns (score, yes, number, depth, p, steps)."""
    present = np.ascontiguousarray(present, dtype=np.uint8)
    isyes = np.ascontiguousarray(isyes, dtype=np.uint8)
    ids = np.ascontiguousarray(target_ids, dtype=np.int32)
    r = OracleResult()
    steps = lib().ppp_oracle_list(_ptr(present), _ptr(isyes), present.size, int(max_yes), _ptr(ids), ids.size,
                                  float(p), int(bool(dup)), ctypes.byref(r))
    return (r.score, r.yes, r.number, r.depth, r.p, steps)


def score_case(query: dict, target, prob=None, dup: bool = False, rank=None, collapse=None):
    """Same inputs as oracle/run_reference.pl: query = {taxon: 0|1}, target = ordered list of taxa
    (ARRAY branch, ppp.pm:151-153) or {taxon: rank} (HASH branch, ppp.pm:147-149; ranks must be distinct).
    rank / collapse = {rank: {taxon: collapsed}}: the -rank pre-pass of ppp.pm:328-370 on every list element
    (collapse_taxon(taxon, rank), else (taxon, "genus"), else the raw id; croak without a taxonomy, :345-349)."""
    taxa = {t: i for i, t in enumerate(query)}
    present = np.ones(len(taxa), dtype=np.uint8)
    isyes = np.array([1 if float(query[t]) != 0 else 0 for t in query], dtype=np.uint8)
    if isinstance(target, dict):
        items = sorted(target.items(), key=lambda kv: kv[1])
        ranks = [v for _, v in items]
        assert len(set(ranks)) == len(ranks), "HASH target with tied ranks is hash-order dependent in the reference"
        tlist = [k for k, _ in items]
    else:
        tlist = list(target)
    if rank:
        if collapse is None:
            raise RuntimeError("Taxonomy does not exist")  # ppp.pm:348
        tlist = [collapse.get(rank, {}).get(t) or collapse.get("genus", {}).get(t) or t for t in tlist]
    ids = []
    for t in tlist:
        if t not in taxa:
            taxa[t] = len(taxa)
        ids.append(taxa[t])
    yes = int(isyes.sum())
    total = len(query)
    p = float(prob) if prob else yes / total  # ppp.pm:138-142 (`if exists _prob`: set only when prob is true, ppp.pm:101)
    s = score_list(present, isyes, yes, ids, p, dup)
    return s[:5]


def score_bits(P, Y, T, G: int, p: float):
    """Canonical-order bit form. P,Y,T uint32 word arrays."""
    P = np.ascontiguousarray(P, dtype=np.uint32)
    Y = np.ascontiguousarray(Y, dtype=np.uint32)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    r = OracleResult()
    steps = lib().ppp_oracle_bits(_ptr(P), _ptr(Y), _ptr(T), int(G), float(p), ctypes.byref(r))
    return (r.score, r.yes, r.number, r.depth, r.p, steps)


def score_bits_batch(P, Y, p, T, G: int, nthreads: int = 1, memo: bool = False):
    """All pairs: P,Y [nq, wpr] uint32; p [nq] float64; T [nt, wpr] uint32 -> structured array [nq, nt], steps.
    memo=True caches bdtrc(yes-1, number, p) per query (a pure function; same loop, same results -- see
    ppp_oracle.c::oracle_list_core): for the full-size parity tests only, never for a timed CPU baseline."""
    P = np.ascontiguousarray(P, dtype=np.uint32)
    Y = np.ascontiguousarray(Y, dtype=np.uint32)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    p = np.ascontiguousarray(p, dtype=np.float64)
    nq, wpr = P.shape
    nt = T.shape[0]
    assert T.shape[1] == wpr and Y.shape == P.shape and p.shape == (nq,)
    out = np.zeros((nq, nt), dtype=RESULT_DTYPE)
    fn = lib().ppp_oracle_bits_batch_memo if memo else lib().ppp_oracle_bits_batch
    steps = fn(_ptr(P), _ptr(Y), _ptr(p), nq, _ptr(T), nt, int(G), wpr, _ptr(out), int(nthreads))
    return out, steps
