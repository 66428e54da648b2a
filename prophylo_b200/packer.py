"""Host-side packers: reference data model -> the HBM layouts of libppp_b200.so.

* query profiles  (taxon -> 0/1 hashes, lib/PhyloProf/Profile.pm:433-514)  -> presence / yes bit planes over a universe
* target sets in canonical (universe) order                                 -> bit planes      (ppp_set_targets_bits)
* target hit lists in their own order (lib/PhyloProf/BlastResult.pm:192-237, value-sorted HMMERResult hashes
  lib/PhyloProf/Algorithm/ppp.pm:147-149)                                   -> ranked lists    (ppp_set_targets_ranked)

Everything here is integer bookkeeping that stays on the host by design (north_star: "BLAST/HMMER parsing, gi2taxid
lookup and profile construction stay on the host").
"""
from __future__ import annotations

import numpy as np

from .synth import pack_bits, words_per_row


def perl_num(v) -> float:
    from .algorithm import perl_num as f  # the value test of Profile.pm:463 (`!= 0` on Perl scalars)

    return f(v)

FOREIGN = 0xFFFF


def universe_index(universe):
    """taxon -> genome index for an ordered universe (ascending taxon id, as data/taxdis.txt lists it)."""
    idx = {}
    for i, t in enumerate(universe):
        if t in idx:
            raise ValueError(f"duplicate taxon {t!r} in universe")
        idx[t] = i
    if len(idx) > 8192:
        raise ValueError("universe larger than 8192 genomes")
    return idx


def pack_queries(queries, index, probs=None):
    """queries: iterable of {taxon: 0|1} (taxa outside the universe are dropped: they can never be hit).
    Returns (P, Y, p): uint32 planes [nq, words_per_row(G)] and p = yes/total per query (ppp.pm:138-142, computed on
    the FULL profile as the reference does) unless probs gives it."""
    G = len(index)
    nq = len(queries)
    Pb = np.zeros((nq, G), dtype=np.uint8)
    Yb = np.zeros((nq, G), dtype=np.uint8)
    p = np.zeros(nq, dtype=np.float64)
    for i, q in enumerate(queries):
        yes = total = 0
        for t, v in q.items():
            total += 1
            isyes = perl_num(v) != 0
            yes += isyes
            g = index.get(t)
            if g is not None:
                Pb[i, g] = 1
                Yb[i, g] = isyes
        p[i] = yes / total if total else float("nan")
    if probs is not None:
        for i, pr in enumerate(probs):
            if pr:  # ppp.pm:101: -prob only counts when true
                p[i] = pr
    return pack_bits(Pb), pack_bits(Yb), p


def pack_target_sets(targets, index):
    """targets: iterable of taxon collections whose list order IS the universe order -> bit planes."""
    G = len(index)
    Tb = np.zeros((len(targets), G), dtype=np.uint8)
    for i, t in enumerate(targets):
        for x in t:
            g = index.get(x)
            if g is not None:
                Tb[i, g] = 1
    return pack_bits(Tb)


def pack_ranked_targets(target_lists, index, dup=False):
    """target_lists: iterable of ordered taxon lists, duplicates and foreign taxa allowed (BLAST hit order).
    De-duplicates keeping the first occurrence (ppp.pm:189-192), keeps the raw 1-based position of every kept entry
    (ppp.pm:169,231) and maps taxa to genome indices (FOREIGN = 0xFFFF for taxa outside the universe).

    dup=True packs for the reference's -dup flag (ppp.pm:189-200, 214-217), whose `$seen{key}` is ONE literal-key
    counter for the whole list: the first repeated entry of the list is counted again like a new taxon (kept here, with
    its genome index repeated); from the second repeated entry on `yes` is frozen while `number` can only grow, so no
    later step can improve the score and the list ends there.  Run the context with set_option("dup", 1): the sweep then
    drops the candidate that would exceed |Y| (ppp.pm:222-224).  (A taxon literally named "key" would alias the
    counter in the reference; such ids do not occur and are not modelled.)
    Returns (idx uint16[total], rawdepth uint16[total], row_off uint64[nT+1])."""
    idx, raw, off = [], [], [0]
    for lst in target_lists:
        seen = set()
        repeats = 0
        if len(lst) > 65535:
            raise ValueError("target list longer than 65535 positions")
        for pos, t in enumerate(lst):
            if t in seen:
                if not dup:
                    continue
                repeats += 1
                if repeats >= 2:
                    break
            seen.add(t)
            idx.append(index.get(t, FOREIGN))
            raw.append(pos + 1)
        off.append(len(idx))
    return (np.asarray(idx, dtype=np.uint16), np.asarray(raw, dtype=np.uint16), np.asarray(off, dtype=np.uint64))


__all__ = ["FOREIGN", "universe_index", "pack_queries", "pack_target_sets", "pack_ranked_targets", "words_per_row"]
