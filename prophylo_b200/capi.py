"""ctypes binding of libppp_b200.so (include/ppp_b200.h) -- the same C ABI the Perl XS module binds.

The library is the product: there is no Python/CPU scoring path.  Loading fails loudly when the shared object has
not been built (`python -c "import __graft_entry__ as g; g.build()"` or ./build_native.sh) and every call fails
with PPPError when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPP_B200_LIB") or os.path.join(_HERE, "libppp_b200.so")  # env override: A/B builds

PPP_OK = 0
FILTER_ALL, FILTER_THRESH_LOG10, FILTER_TOPK = 0, 1, 2

HIT_DTYPE = np.dtype(
    [("q", "<u4"), ("t", "<u4"), ("score", "<f8"), ("yes", "<u4"), ("number", "<u4"), ("depth", "<u4"), ("margin", "<f4")]
)
assert HIT_DTYPE.itemsize == 32

EXPORTS = [
    "ppp_init", "ppp_destroy", "ppp_last_error", "ppp_abi_version", "ppp_set_targets_bits", "ppp_set_targets_bits_streamed", "ppp_set_targets_bits_device", "ppp_wait_targets", "ppp_set_targets_ranked",
    "ppp_set_queries", "ppp_run", "ppp_result_count", "ppp_fetch", "ppp_score", "ppp_get_stats", "ppp_set_option",
    "ppp_get_counter", "ppp_words_per_row", "ppp_debug_bdtrc", "ppp_topk_device", "ppp_merge_topk_device", "ppp_read_bla", "ppp_read_bla_flags", "ppp_printed_cutoffs",
    "ppp_free_ranked_lists", "ppp_debug_enclosure", "ppp_query_count", "ppp_row_words", "ppp_pipe_peak", "ppp_read_bla_cached", "ppp_read_profiles", "ppp_free_profile_set",
]


class PPPError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libppp_b200 error {code}: {msg}")
        self.code = code


class Filter(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_uint32), ("k", ctypes.c_uint32), ("thresh", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [
        ("pairs", ctypes.c_uint64), ("candidates", ctypes.c_uint64), ("accurate_evals", ctypes.c_uint64),
        ("exact_pairs", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("sweep_kernel_ms", ctypes.c_float),
        ("total_device_ms", ctypes.c_float), ("sweep_launches", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
    ]


class ProfileSet(ctypes.Structure):
    _fields_ = [("universe_buf", ctypes.c_void_p), ("universe_off", ctypes.POINTER(ctypes.c_size_t)), ("P", ctypes.POINTER(ctypes.c_uint32)),
                ("Y", ctypes.POINTER(ctypes.c_uint32)), ("p", ctypes.POINTER(ctypes.c_double)), ("yes", ctypes.POINTER(ctypes.c_uint32)),
                ("total", ctypes.POINTER(ctypes.c_uint32)), ("n", ctypes.c_size_t), ("G", ctypes.c_size_t), ("wpr", ctypes.c_size_t)]


class RankedLists(ctypes.Structure):
    _fields_ = [("idx", ctypes.POINTER(ctypes.c_uint16)), ("rawdepth", ctypes.POINTER(ctypes.c_uint16)),
                ("row_off", ctypes.POINTER(ctypes.c_uint64)), ("n_targets", ctypes.c_size_t), ("n_entries", ctypes.c_size_t)]


_lib = None


def load():
    """dlopen libppp_b200.so and declare the prototypes of include/ppp_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built; run ./build_native.sh (nvcc, sm_100a). There is no fallback path.")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, u32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32
    L.ppp_init.restype = ctypes.c_int
    L.ppp_init.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    L.ppp_destroy.restype = None
    L.ppp_destroy.argtypes = [vp]
    L.ppp_last_error.restype = ctypes.c_char_p
    L.ppp_last_error.argtypes = [vp]
    L.ppp_abi_version.restype = ctypes.c_int
    L.ppp_set_targets_bits.restype = ctypes.c_int
    L.ppp_set_targets_bits.argtypes = [vp, vp, sz, u32]
    L.ppp_set_targets_bits_streamed.restype = ctypes.c_int
    L.ppp_set_targets_bits_streamed.argtypes = [vp, vp, sz, u32]
    L.ppp_set_targets_bits_device.restype = ctypes.c_int
    L.ppp_set_targets_bits_device.argtypes = [vp, vp, sz, u32]
    L.ppp_wait_targets.restype = ctypes.c_int
    L.ppp_wait_targets.argtypes = [vp]
    L.ppp_set_targets_ranked.restype = ctypes.c_int
    L.ppp_set_targets_ranked.argtypes = [vp, vp, vp, vp, sz, u32]
    L.ppp_set_queries.restype = ctypes.c_int
    L.ppp_set_queries.argtypes = [vp, vp, vp, vp, sz]
    L.ppp_run.restype = ctypes.c_int
    L.ppp_run.argtypes = [vp, ctypes.POINTER(Filter)]
    L.ppp_result_count.restype = ctypes.c_int
    L.ppp_result_count.argtypes = [vp, ctypes.POINTER(sz)]
    L.ppp_fetch.restype = ctypes.c_int
    L.ppp_fetch.argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
    L.ppp_score.restype = ctypes.c_int
    L.ppp_score.argtypes = [vp, vp, vp, vp, sz, ctypes.POINTER(Filter), vp, sz, ctypes.POINTER(sz)]
    L.ppp_get_stats.restype = ctypes.c_int
    L.ppp_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.ppp_set_option.restype = ctypes.c_int
    L.ppp_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    L.ppp_get_counter.restype = ctypes.c_int
    L.ppp_get_counter.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)]
    L.ppp_words_per_row.restype = u32
    L.ppp_words_per_row.argtypes = [u32]
    L.ppp_topk_device.restype = ctypes.c_int
    L.ppp_topk_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(u32), ctypes.POINTER(sz)]
    L.ppp_merge_topk_device.restype = ctypes.c_int
    L.ppp_merge_topk_device.argtypes = [vp, vp, vp, u32, vp, sz, u32, vp, vp]
    L.ppp_debug_bdtrc.restype = ctypes.c_int
    L.ppp_debug_bdtrc.argtypes = [vp, vp, vp, vp, vp, sz]
    L.ppp_debug_enclosure.restype = ctypes.c_int
    L.ppp_debug_enclosure.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz]
    L.ppp_pipe_peak.restype = ctypes.c_int
    L.ppp_pipe_peak.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
    L.ppp_read_bla.restype = ctypes.c_int
    L.ppp_read_bla.argtypes = [ctypes.POINTER(ctypes.c_char_p), sz, ctypes.POINTER(ctypes.c_char_p), sz, ctypes.c_int,
                               ctypes.POINTER(RankedLists), ctypes.c_char_p, sz]
    L.ppp_read_bla_flags.restype = ctypes.c_int
    L.ppp_read_bla_flags.argtypes = [ctypes.POINTER(ctypes.c_char_p), sz, ctypes.POINTER(ctypes.c_char_p), sz, ctypes.c_int, u32,
                                     ctypes.POINTER(RankedLists), ctypes.c_char_p, sz]
    L.ppp_read_bla_cached.restype = ctypes.c_int
    L.ppp_read_bla_cached.argtypes = [ctypes.POINTER(ctypes.c_char_p), sz, ctypes.POINTER(ctypes.c_char_p), sz, ctypes.c_int, u32, ctypes.c_char_p,
                                      ctypes.POINTER(RankedLists), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, sz]
    L.ppp_read_profiles.restype = ctypes.c_int
    L.ppp_read_profiles.argtypes = [ctypes.POINTER(ctypes.c_char_p), sz, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ProfileSet), ctypes.c_char_p, sz]
    L.ppp_free_profile_set.restype = None
    L.ppp_free_profile_set.argtypes = [ctypes.POINTER(ProfileSet)]
    L.ppp_printed_cutoffs.restype = ctypes.c_int
    L.ppp_printed_cutoffs.argtypes = [vp, ctypes.c_double, vp, vp, vp, vp, vp]
    L.ppp_free_ranked_lists.restype = None
    L.ppp_free_ranked_lists.argtypes = [ctypes.POINTER(RankedLists)]
    _lib = L
    return L


BLA_DUP = 1  # PPP_BLA_DUP
CUT_OK, CUT_DOUBT, CUT_DIV0, CUT_LOG0 = 0, 1, 2, 3
CUT_NONE = -(2 ** 31)


def read_profile_files(paths, header: bool = False, nthreads: int = 0):
    """Native reader + bit-plane packer of a family of profile files (ppp_read_profiles, prophylo_b200/csrc/ppp_pack.cpp).
    Returns (universe list, P, Y, p, yes, total): shared universe in ascending numeric taxon order, packed planes [n, wpr]
    uint32, p = yes / total per file.  No GPU needed."""
    L = load()
    pa = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    out = ProfileSet()
    err = ctypes.create_string_buffer(512)
    rc = L.ppp_read_profiles(pa, len(paths), int(bool(header)), nthreads, ctypes.byref(out), err, len(err))
    if rc != PPP_OK:
        raise PPPError(rc, err.value.decode(errors="replace"))
    try:
        n, G, wpr = out.n, out.G, out.wpr
        offs = np.ctypeslib.as_array(out.universe_off, shape=(max(G, 1),))[:G]
        universe = [ctypes.string_at(out.universe_buf + int(o)).decode() for o in offs]
        P = np.ctypeslib.as_array(out.P, shape=(max(n * wpr, 1),))[: n * wpr].reshape(n, wpr).copy()
        Y = np.ctypeslib.as_array(out.Y, shape=(max(n * wpr, 1),))[: n * wpr].reshape(n, wpr).copy()
        p = np.ctypeslib.as_array(out.p, shape=(max(n, 1),))[:n].copy()
        yes = np.ctypeslib.as_array(out.yes, shape=(max(n, 1),))[:n].copy()
        total = np.ctypeslib.as_array(out.total, shape=(max(n, 1),))[:n].copy()
    finally:
        L.ppp_free_profile_set(ctypes.byref(out))
    return universe, P, Y, p, yes, total


def read_bla_files(paths, universe, nthreads: int = 0, dup: bool = False, cache: str | None = None, info: dict | None = None):
    """Native reader of `.bla` BLAST caches (ppp_read_bla, prophylo_b200/csrc/ppp_bla.cpp): one ranked target per path,
    taxa mapped through `universe` (ordered taxon strings).  Returns (idx, rawdepth, row_off) as ppp_set_targets_ranked
    takes them; dup=True packs for the reference's -dup flag (see packer.pack_ranked_targets).  cache = path of a persistent
    packed cache of the whole path set (ppp_read_bla_cached; info["from_cache"] tells whether it was used).  No GPU needed."""
    L = load()
    pa = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    ua = (ctypes.c_char_p * len(universe))(*[str(u).encode() for u in universe])
    out = RankedLists()
    err = ctypes.create_string_buffer(512)
    hit = ctypes.c_int(0)
    rc = L.ppp_read_bla_cached(pa, len(paths), ua, len(universe), nthreads, BLA_DUP if dup else 0, os.fsencode(cache) if cache else None,
                               ctypes.byref(out), ctypes.byref(hit), err, len(err))
    if info is not None:
        info["from_cache"] = bool(hit.value)
    if rc != PPP_OK:
        raise PPPError(rc, err.value.decode(errors="replace"))
    try:
        n, m = out.n_entries, out.n_targets
        idx = np.ctypeslib.as_array(out.idx, shape=(max(n, 1),))[:n].copy()
        raw = np.ctypeslib.as_array(out.rawdepth, shape=(max(n, 1),))[:n].copy()
        off = np.ctypeslib.as_array(out.row_off, shape=(m + 1,)).copy()
    finally:
        L.ppp_free_ranked_lists(ctypes.byref(out))
    return idx, raw, off


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_filter(mode=FILTER_ALL, k=0, thresh=0.0):
    return Filter(mode, k, thresh)


class Context:
    """One ppp_ctx.  device = one CUDA ordinal (one GPU, the one-process-per-GPU form) or a list of ordinals (ONE process
    drives them all: targets sharded over the devices, results assembled by the library).  Create after any fork."""

    def __init__(self, device=0):
        self.lib = load()
        self._h = ctypes.c_void_p()
        devs = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        dev = (ctypes.c_int * len(devs))(*devs)
        rc = self.lib.ppp_init(ctypes.byref(self._h), dev, len(devs))
        if rc != PPP_OK:
            raise PPPError(rc, self.lib.ppp_last_error(None).decode())
        self.G = None
        self.nT = 0
        self.nQ = 0

    def close(self):
        if self._h:
            self.lib.ppp_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != PPP_OK:
            raise PPPError(rc, self.lib.ppp_last_error(self._h).decode())

    def words_per_row(self, G):
        return int(self.lib.ppp_words_per_row(G))

    def set_targets_bits(self, T: np.ndarray, G: int, streamed: bool = False):
        """Upload packed target planes.  streamed=True (ppp_set_targets_bits_streamed) only queues the copies: T should be
        page-locked, is kept referenced by this object, and must not be modified until the next run() / score() returned."""
        T = np.ascontiguousarray(T, dtype=np.uint32)
        assert T.ndim == 2 and T.shape[1] == self.words_per_row(G), (T.shape, G)
        if streamed:
            self._streamed_T = T  # keeps the host buffer alive while the copies are in flight
            self._check(self.lib.ppp_set_targets_bits_streamed(self._h, _ptr(T), T.shape[0], G))
        else:
            self._check(self.lib.ppp_set_targets_bits(self._h, _ptr(T), T.shape[0], G))
            self._streamed_T = None
        self.G, self.nT = G, T.shape[0]

    def set_targets_bits_device(self, dptr: int, nT: int, G: int):
        """Packed target planes that already sit in device memory ([nT][words_per_row(G)] uint32 at address dptr, e.g. a torch
        tensor's data_ptr() after an NCCL all-gather): copied device to device (ppp_set_targets_bits_device).  The producer's
        stream must have been synchronised."""
        self._check(self.lib.ppp_set_targets_bits_device(self._h, ctypes.c_void_p(dptr), nT, G))
        self._streamed_T = None
        self.G, self.nT = G, nT

    def wait_targets(self):
        self._check(self.lib.ppp_wait_targets(self._h))

    def set_targets_ranked(self, idx, rawdepth, row_off, G: int):
        """Ordered, de-duplicated hit lists (packer.pack_ranked_targets)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint16)
        rawdepth = np.ascontiguousarray(rawdepth, dtype=np.uint16)
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
        assert idx.shape == rawdepth.shape and row_off.ndim == 1 and row_off.size >= 1
        self._check(self.lib.ppp_set_targets_ranked(self._h, _ptr(idx), _ptr(rawdepth), _ptr(row_off), row_off.size - 1, G))
        self.G, self.nT = G, row_off.size - 1

    def set_queries(self, P, Y, p):
        P = np.ascontiguousarray(P, dtype=np.uint32)
        Y = np.ascontiguousarray(Y, dtype=np.uint32)
        p = np.ascontiguousarray(p, dtype=np.float64)
        assert P.shape == Y.shape and P.ndim == 2 and p.shape == (P.shape[0],)
        self._check(self.lib.ppp_set_queries(self._h, _ptr(P), _ptr(Y), _ptr(p), P.shape[0]))
        self.nQ = P.shape[0]

    def run(self, flt: Filter | None = None):
        self._check(self.lib.ppp_run(self._h, ctypes.byref(flt) if flt is not None else None))

    def result_count(self) -> int:
        n = ctypes.c_size_t()
        self._check(self.lib.ppp_result_count(self._h, ctypes.byref(n)))
        return n.value

    def fetch(self, out: np.ndarray | None = None) -> np.ndarray:
        n = self.result_count()
        if out is None:
            out = np.empty(n, dtype=HIT_DTYPE)
        got = ctypes.c_size_t()
        self._check(self.lib.ppp_fetch(self._h, _ptr(out), out.size, ctypes.byref(got)))
        return out[: got.value]

    def score(self, P, Y, p, flt: Filter | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """One-call form with host buffers (what the Perl binding calls)."""
        P = np.ascontiguousarray(P, dtype=np.uint32)
        Y = np.ascontiguousarray(Y, dtype=np.uint32)
        p = np.ascontiguousarray(p, dtype=np.float64)
        if out is None:
            if flt is None or flt.mode == FILTER_ALL:
                cap = P.shape[0] * self.nT
            elif flt.mode == FILTER_TOPK:
                cap = P.shape[0] * min(flt.k, self.nT)
            else:
                cap = 0
            out = np.empty(cap, dtype=HIT_DTYPE)
        got = ctypes.c_size_t()
        rc = self.lib.ppp_score(self._h, _ptr(P), _ptr(Y), _ptr(p), P.shape[0], ctypes.byref(flt) if flt is not None else None,
                                _ptr(out), out.size, ctypes.byref(got))
        if rc == -5:  # PPP_ERR_CAPACITY: results are resident, fetch them into a big enough buffer
            out = np.empty(got.value, dtype=HIT_DTYPE)
            self._check(self.lib.ppp_fetch(self._h, _ptr(out), out.size, ctypes.byref(got)))
        else:
            self._check(rc)
        self.nQ = P.shape[0]
        return out[: got.value]

    def printed_cutoffs(self, slope: float, want_milli: bool = True) -> dict:
        """Printed scores and slope cut-offs of the last threshold run, computed on the device (ppp_printed_cutoffs):
        row_off [nQ+1], milli [n hits] (sprintf "%.3f" of log10 score, in thousandths), marker [nQ] (ppp.pl -m, -1 = none),
        cutoff_milli [nQ] (ppp_cutoff.pl, CUT_NONE = none), status [nQ] (CUT_OK / CUT_DOUBT / CUT_DIV0 / CUT_LOG0)."""
        n, nq = self.result_count(), self.nQ
        off = np.empty(nq + 1, dtype=np.uint64)
        milli = np.empty(n if want_milli else 0, dtype=np.int32)
        marker = np.empty(nq, dtype=np.int32)
        cut = np.empty(nq, dtype=np.int32)
        status = np.empty(nq, dtype=np.uint8)
        self._check(self.lib.ppp_printed_cutoffs(self._h, float(slope), _ptr(off), _ptr(milli) if want_milli else None, _ptr(marker),
                                                 _ptr(cut), _ptr(status)))
        return {"row_off": off, "milli": milli if want_milli else None, "marker": marker, "cutoff_milli": cut, "status": status}

    def topk_device(self):
        """(lists_ptr, counts_ptr, k, nQ): device addresses of the last top-k run's [nQ][k] hits and [nQ] counts."""
        lp, cp, k, nq = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_size_t()
        self._check(self.lib.ppp_topk_device(self._h, ctypes.byref(lp), ctypes.byref(cp), ctypes.byref(k), ctypes.byref(nq)))
        return lp.value, cp.value, k.value, nq.value

    def merge_topk_device(self, lists_ptr: int, counts_ptr: int, nlists: int, t_offsets, n_queries: int, k: int, out=None, cnt=None):
        """Merge `nlists` gathered device lists into the global per-query top-k; returns (hits [nQ*k], counts [nQ]) on the
        host (`out` / `cnt`: caller's buffers, e.g. pinned memory)."""
        offs = np.ascontiguousarray(t_offsets, dtype=np.uint32)
        if out is None:
            out = np.empty(n_queries * k, dtype=HIT_DTYPE)
        if cnt is None:
            cnt = np.empty(n_queries, dtype=np.uint32)
        assert out.dtype == HIT_DTYPE and out.size >= n_queries * k and cnt.dtype == np.uint32 and cnt.size >= n_queries
        self._check(self.lib.ppp_merge_topk_device(self._h, ctypes.c_void_p(lists_ptr), ctypes.c_void_p(counts_ptr), nlists, _ptr(offs),
                                                   n_queries, k, _ptr(out), _ptr(cnt)))
        return out, cnt

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.ppp_get_stats(self._h, ctypes.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_ if f[0] != "reserved"}

    def set_option(self, name: str, value: int):
        self._check(self.lib.ppp_set_option(self._h, name.encode(), int(value)))

    def counter(self, name: str) -> int:
        v = ctypes.c_int64()
        self._check(self.lib.ppp_get_counter(self._h, name.encode(), ctypes.byref(v)))
        return v.value

    def debug_bdtrc(self, k, n, p) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.int32)
        n = np.ascontiguousarray(n, dtype=np.int32)
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.empty(k.shape, dtype=np.float64)
        self._check(self.lib.ppp_debug_bdtrc(self._h, _ptr(k), _ptr(n), _ptr(p), _ptr(out), k.size))
        return out

    def debug_enclosure(self, yes, number, p):
        """(lo, hi, lnf): tier A's float enclosure and tier B's FP64 value of ln P(X >= yes | number, p) at each point."""
        yes = np.ascontiguousarray(yes, dtype=np.int32)
        number = np.ascontiguousarray(number, dtype=np.int32)
        p = np.ascontiguousarray(p, dtype=np.float64)
        lo = np.empty(yes.shape, dtype=np.float32)
        hi = np.empty(yes.shape, dtype=np.float32)
        lnf = np.empty(yes.shape, dtype=np.float64)
        self._check(self.lib.ppp_debug_enclosure(self._h, _ptr(yes), _ptr(number), _ptr(p), _ptr(lo), _ptr(hi), _ptr(lnf), yes.size))
        return lo, hi, lnf

    def pipe_peak(self, pipe: str) -> float:
        """measured lane-operations per second of one issue pipe ("popc", "lop3", "iadd", "dfma") on this GPU"""
        v = ctypes.c_double()
        self._check(self.lib.ppp_pipe_peak(self._h, pipe.encode(), ctypes.byref(v)))
        return v.value
