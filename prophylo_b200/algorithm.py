"""Python mirror of the reference's operator interface for the PPP path -- same names, argument meaning and error
behaviour as lib/PhyloProf/Profile.pm and lib/PhyloProf/Algorithm/ppp.pm -- scored on the B200 through the C ABI.

    prof1 = Profile(profile={"a": 1, "b": 0})            # Profile->new(-profile => {...})     Profile.pm:120-203
    prof2 = Profile(raw=["a", "zz", "a", "b"], yes=4, no=0)  # Profile->new(-raw=>..., -yes, -no)  Profile.pm:93-104
    score, yes, number, depth, p = ppp(prof1, prof2, prob=None).get_match_score()     # ppp.pm:130-245

There is no CPU scoring path here: get_match_score() needs libppp_b200.so and an sm_100 GPU.
"""
from __future__ import annotations

import numpy as np

from . import capi, packer


class Profile:
    """PhyloProf::Profile: a query is a taxon -> 0/1 hash with yes/no counts; a target carrier holds an ordered list
    (ARRAY) or a taxon -> rank hash (HASH) under -raw."""

    def __init__(self, file=None, profile=None, raw=None, yes=None, no=None, header=False):
        self._prof, self._yes, self._no = None, 0, 0
        if raw is not None and yes is not None and no is not None:  # Profile.pm:93-104
            self._prof, self._yes, self._no = raw, yes, no
            return
        if file is not None:
            profile = self._parse_file(file, header)
        if profile is not None:
            h = {}
            for k, v in profile.items():
                if not k:
                    continue
                if k in h and float(h[k]) != 0:  # a non-zero value is never overwritten (Profile.pm:463-468)
                    continue
                h[k] = v
            self._prof = h
            self._yes = sum(1 for v in h.values() if float(v) != 0)
            self._no = len(h) - self._yes

    @staticmethod
    def _parse_file(path, header):
        """2 tab-separated columns taxon, 0/1; '#' comments; optional header line (Profile.pm:433-514)."""
        out = {}
        first = True
        with open(path) as fh:
            for line in fh:
                line = line.rstrip("\r\n")
                if not line or line.startswith("#"):
                    continue
                if first and header:
                    first = False
                    continue
                first = False
                f = line.split("\t")
                if len(f) != 2:
                    raise ValueError(f"{path} must contain only two fields")
                if f[0] in out and float(out[f[0]]) != 0:
                    continue
                out[f[0]] = f[1]
        return out

    def get_hash(self):
        return self._prof

    def get_yes(self):
        return self._yes

    def get_no(self):
        return self._no

    def get_total(self):
        return self._yes + self._no

    def is_present(self, taxon):
        return taxon in self._prof

    def is_yes(self, taxon):
        return taxon in self._prof and float(self._prof[taxon]) != 0


_ctx = None


def context(device: int = 0) -> capi.Context:
    global _ctx
    if _ctx is None:
        _ctx = capi.Context(device)
    return _ctx


class ppp:
    """PhyloProf::Algorithm::ppp (ppp.pm:58-110): ppp(prof1, prof2, prob, rank, taxonomy, dup)."""

    def __init__(self, prof1=None, prof2=None, prob=None, rank=None, taxonomy=None, dup=None):
        if not (prof1 or prof2):
            raise ValueError("Parameter is undefined")  # ppp.pm:85
        if not (isinstance(prof1, Profile) or isinstance(prof2, Profile)):
            raise TypeError("Parameter mismatch")  # ppp.pm:87-91
        self._prof1, self._prof2 = prof1, prof2
        self._prob = prob if prob else None  # ppp.pm:101: only a true -prob is kept
        self._rank, self._dbh, self._dup = rank, taxonomy, bool(dup)

    def _taxon(self, t):  # ppp.pm:328-370
        if self._rank and self._dbh is not None:
            return self._dbh.collapse_taxon(t, self._rank) or self._dbh.collapse_taxon(t, "genus") or t
        return t

    def get_match_score(self):
        q = self._prof1.get_hash()
        total = self._prof1.get_total()
        if self._prob is None and total == 0:
            raise ZeroDivisionError("Illegal division by zero")  # ppp.pm:141
        p = self._prob if self._prob is not None else self._prof1.get_yes() / total
        t = self._prof2.get_hash()
        if isinstance(t, dict):
            lst = [k for k, _ in sorted(t.items(), key=lambda kv: kv[1])]  # ppp.pm:147-149
        else:
            lst = list(t or [])
        lst = [self._taxon(x) for x in lst]
        if any(not x for x in lst):
            raise ValueError("Taxon not found")  # ppp.pm:179
        if not lst:
            return (1, 0, 0, 0, p)
        universe = list(q.keys())
        if len(universe) > 8192:
            raise ValueError("query profile larger than 8192 taxa")
        index = packer.universe_index(universe) if universe else {}
        if not index:
            return (1, 0, 0, 0, p)
        P, Y, _ = packer.pack_queries([q], index)
        idx, raw, off = packer.pack_ranked_targets([lst], index, dup=self._dup)
        ctx = context()
        ctx.set_targets_ranked(idx, raw, off, len(index))
        ctx.set_option("dup", int(self._dup))  # ppp.pm:189-200, 222-224
        try:
            h = ctx.score(P, Y, np.array([p], dtype=np.float64))[0]
        finally:
            ctx.set_option("dup", 0)
        return (float(h["score"]), int(h["yes"]), int(h["number"]), int(h["depth"]), p)
