"""Synthetic PPP workloads: deterministic numpy generators for packed bit-plane
profiles in canonical genome order.  Used by tests, the golden-vector generator and bench.py.

Bit layout (the packer's layout, see prophylo_b200/packer.py): genome g is bit (g & 31) of uint32 word
(g >> 5); rows are padded with zero words to a multiple of 4 words (128-bit loads).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 20261019


def words_per_row(G: int) -> int:
    return ((G + 127) // 128) * 4


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """bool/uint8 [n, G] -> uint32 [n, words_per_row(G)] (little-endian bit order)."""
    bits = np.asarray(bits, dtype=np.uint8)
    n, G = bits.shape
    wpr = words_per_row(G)
    by = np.packbits(bits, axis=1, bitorder="little")
    out = np.zeros((n, wpr * 4), dtype=np.uint8)
    out[:, : by.shape[1]] = by
    return out.view("<u4").reshape(n, wpr)


def unpack_bits(words: np.ndarray, G: int) -> np.ndarray:
    by = np.ascontiguousarray(words, dtype="<u4").view(np.uint8)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :G]


def make_queries(G: int, nq: int, seed: int, kind: str = "mixed"):
    """Query bit planes.  Returns (P, Y, p) with Y subset of P, p = |Y|/|P| (ppp.pm:138-142).

    kind = "single"  : P all-ones, Y ~ Bernoulli(0.10)                            (config 2)
           "mixed"   : 75% P all-ones, 25% P ~ Bernoulli(0.6); Y density log-uniform(0.005, 0.3) within P;
                       5% degenerate cross-score form P = Y with p = d/G          (config 3)
           "prefix"  : nested prefixes of one random permutation, Y_d = first d genomes (d = 5..), P all-ones,
                       p = d/G  (bin/ppp2d.pl:315)                                (config 4)
           "family"  : P all-ones, Y density log-uniform(0.002, 0.2)              (config 5)
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    Pb = np.ones((nq, G), dtype=np.uint8)
    Yb = np.zeros((nq, G), dtype=np.uint8)
    p = np.zeros(nq, dtype=np.float64)
    if kind == "single":
        Yb[:] = rng.random((nq, G)) < 0.10
    elif kind in ("mixed", "family"):
        lo, hi = (0.005, 0.3) if kind == "mixed" else (0.002, 0.2)
        dens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=nq))
        if kind == "mixed":
            partial = rng.random(nq) < 0.25
            Pb[partial] = rng.random((int(partial.sum()), G)) < 0.6
        Yb[:] = (rng.random((nq, G)) < dens[:, None]) & (Pb != 0)
        if kind == "mixed":
            degen = rng.random(nq) < 0.05
            Pb[degen] = Yb[degen]
    elif kind == "prefix":
        perm = rng.permutation(G)
        for i in range(nq):
            d = min(5 + i, G)
            Yb[i, perm[:d]] = 1
    else:
        raise ValueError(kind)
    for i in range(nq):
        if Yb[i].sum() == 0:  # keep every query scoreable (total > 0, yes > 0)
            g = int(rng.integers(G))
            Yb[i, g] = 1
            Pb[i, g] = 1
    yes = Yb.sum(axis=1).astype(np.float64)
    tot = Pb.sum(axis=1).astype(np.float64)
    p[:] = yes / tot
    if kind == "prefix":
        p[:] = yes / G
    if kind == "mixed":
        degen_rows = (Pb == Yb).all(axis=1)
        p[degen_rows] = yes[degen_rows] / G
    return pack_bits(Pb), pack_bits(Yb), p


def make_targets(G: int, nt: int, seed: int, plant_from: np.ndarray | None = None, plant_frac: float = 0.01):
    """Target bit planes: density d_t ~ U(0.02, 0.9), bits ~ Bernoulli(d_t); a fraction `plant_frac` of rows
    is planted as (row of plant_from) xor Bernoulli(0.05)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((nt, words_per_row(G)), dtype=np.uint32)
    chunk = max(1, (1 << 26) // max(G, 1))
    for s in range(0, nt, chunk):
        e = min(nt, s + chunk)
        dens = rng.uniform(0.02, 0.9, size=e - s)
        bits = rng.random((e - s, G), dtype=np.float32) < dens[:, None].astype(np.float32)
        if plant_from is not None and plant_frac > 0:
            planted = np.nonzero(rng.random(e - s) < plant_frac)[0]
            if planted.size:
                src = unpack_bits(plant_from, G)
                pick = rng.integers(0, src.shape[0], size=planted.size)
                noise = rng.random((planted.size, G)) < 0.05
                bits[planted] = (src[pick] != 0) ^ noise
        out[s:e] = pack_bits(bits)
    return out
