"""Python mirror of the reference's operator interface for the PPP path -- same names, argument meaning and error
behaviour as lib/PhyloProf/Profile.pm and lib/PhyloProf/Algorithm/ppp.pm -- scored on the B200 through the C ABI.

    prof1 = Profile(profile={"a": 1, "b": 0})            # Profile->new(-profile => {...})     Profile.pm:120-203
    prof2 = Profile(raw=["a", "zz", "a", "b"], yes=4, no=0)  # Profile->new(-raw=>..., -yes, -no)  Profile.pm:93-104
    score, yes, number, depth, p = ppp(prof1, prof2, prob=None).get_match_score()     # ppp.pm:130-245

There is no CPU scoring path here: get_match_score() needs libppp_b200.so and an sm_100 GPU.
"""
from __future__ import annotations

import re

import numpy as np

from . import capi, packer


_NUM = re.compile(r"^\s*[+-]?(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)")


def perl_num(v) -> float:
    """Perl's numeric value of a scalar: the leading number of a string, 0 when there is none (`$hash{$taxon} != 0`,
    Profile.pm:463, 477)."""
    if isinstance(v, (int, float)):
        return float(v)
    m = _NUM.match(str(v))
    return float(m.group(0)) if m else 0.0


def perl_true(v) -> bool:
    """Perl truth of a scalar: undef, '' and '0' are false (`next unless $taxon`, Profile.pm:460)."""
    return v is not None and v != "" and v != "0" and v != 0


def perl_split_tab(line: str):
    """split(/\t/, $line): trailing empty fields are dropped (Profile.pm:452)."""
    f = line.split("\t")
    while f and f[-1] == "":
        f.pop()
    return f


def collapse(taxonomy, rank, taxon, what="Taxonomy does not exist"):
    """_get_taxon of ppp.pm:328-370 / Profile.pm:516-549: identity unless a rank is set; collapse_taxon(taxon, rank), else
    collapse_taxon(taxon, "genus"), else the raw id.  The algorithm croaks without a taxonomy object (ppp.pm:345-349); the
    profile class would open the SeqToolBox SQLite database instead (Profile.pm:526-530), which this host does not have."""
    if not rank:
        return taxon
    if taxonomy is None:
        raise RuntimeError(what)
    return taxonomy.collapse_taxon(taxon, rank) or taxonomy.collapse_taxon(taxon, "genus") or taxon


class NodesTaxonomy:
    """collapse_taxon over NCBI `nodes` rows held in memory: {tax_id: (parent_tax_id, rank)} or a 3-column TSV file
    (tax_id, parent, rank).  Same walk as SeqToolBox::Taxonomy::collapse_taxon (lib/SeqToolBox/Taxonomy.pm:296-360) does
    against its SQLite copy of nodes.dmp: climb from the id until a node of the wanted rank (case-insensitive) is found;
    stop at the root (tax_id 1) or at an id with no row -> None."""

    def __init__(self, rows=None, file=None):
        self.rows = dict(rows or {})
        if file is not None:
            with open(file) as fh:
                for line in fh:
                    f = line.rstrip("\n").split("\t")
                    if len(f) >= 3:
                        self.rows[f[0]] = (f[1], f[2])
        self._cache = {}

    def collapse_taxon(self, taxon, rank):
        if not perl_true(taxon):
            raise ValueError("ERROR: ID not defined")
        if not rank:
            raise ValueError("ERROR: RANK not defined")
        key = (str(taxon), rank.lower())
        if key in self._cache:
            return self._cache[key]
        taxid, found = str(taxon), None
        while perl_true(taxid) and taxid != "1":
            row = self.rows.get(taxid)
            if row is None:
                break
            parent, node_rank = row
            if node_rank is not None and rank.lower() == str(node_rank).lower():
                found = taxid
                break
            taxid = parent
        self._cache[key] = found
        return found


class Profile:
    """PhyloProf::Profile: a query is a taxon -> 0/1 hash with yes/no counts; a target carrier holds an ordered list
    (ARRAY) or a taxon -> rank hash (HASH) under -raw.  `rank` / `taxonomy` collapse the query's taxa as
    Profile.pm:455-466, 516-549 does (a host-side pre-pass)."""

    def __init__(self, file=None, profile=None, raw=None, yes=None, no=None, header=False, rank=None, taxonomy=None):
        self._prof, self._yes, self._no = None, 0, 0
        self._rank, self._dbh = rank, taxonomy
        if raw is not None and yes is not None and no is not None:  # Profile.pm:93-104
            self._prof, self._yes, self._no = raw, yes, no
            return
        items = None
        if file is not None:
            items = self._parse_file(file, header)
        elif profile is not None:
            items = list(profile.items())
        if items is not None:
            h = {}
            for k, v in items:
                k = self._get_taxon(k)
                if not perl_true(k):  # `next unless $taxon` (Profile.pm:460, 139): '' and '0' are dropped
                    continue
                if k in h and perl_num(h[k]) != 0:  # a non-zero value is never overwritten (Profile.pm:463-468)
                    continue
                h[k] = v
            self._prof = h
            self._yes = sum(1 for v in h.values() if perl_num(v) != 0)
            self._no = len(h) - self._yes

    def _get_taxon(self, taxon):  # Profile.pm:516-549
        return collapse(self._dbh, self._rank, taxon, "Taxonomy object does not exist (no SeqToolBox database on this host)")

    @staticmethod
    def _parse_file(path, header):
        """2 tab-separated columns taxon, 0/1 (Profile.pm:433-514): with -header physical line 1 is skipped whatever it
        holds (:447-449); '#' lines are comments (:450); every other line -- an empty one included -- must split into
        exactly two fields (:452-456)."""
        out = []
        try:
            fh = open(path, newline="\n")  # chomp removes "\n" only
        except OSError:
            raise OSError(f"Can' open {path}")  # the reference's own spelling (Profile.pm:436)
        with fh:
            for lineno, line in enumerate(fh, 1):
                line = line[:-1] if line.endswith("\n") else line
                if header and lineno == 1:
                    continue
                if line.startswith("#"):
                    continue
                f = perl_split_tab(line)
                if len(f) != 2:
                    raise ValueError(f"{path} must contain only two fields")
                out.append((f[0], f[1]))
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
        return taxon in self._prof and perl_num(self._prof[taxon]) != 0


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

    def _taxon(self, t):  # ppp.pm:328-370; croaks "Taxonomy does not exist" when -rank is set without -taxonomy (:345-349)
        return collapse(self._dbh, self._rank, t)

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
        if any(not perl_true(x) for x in lst):
            raise ValueError("Taxon not found")  # ppp.pm:179
        if not lst:
            return (1, 0, 0, 0, p)
        universe = list(q.keys())
        if len(universe) > 8192:
            raise ValueError("query profile larger than 8192 taxa")
        index = packer.universe_index(universe) if universe else {}
        if not index:
            return (1, 0, 0, 0, p)
        ctx = context()
        distinct = list(dict.fromkeys(lst))
        if not self._dup and len(distinct) > 2048:
            # lists of more than 2048 distinct taxa (the ranked entry's limit): embed the pair exactly by taking the target's
            # own de-duplicated list as the genome order -- the target plane is a run of ones, the query planes are permuted
            # to match, and the raw list position is restored from the de-duplicated depth (as ppp_b200.pm does)
            if len(distinct) > 8192:
                raise ValueError("target list of more than 8192 distinct taxa")
            first_pos = {}
            for i, t_ in enumerate(lst):
                first_pos.setdefault(t_, i + 1)
            tindex = packer.universe_index(distinct)
            P, Y, _ = packer.pack_queries([{k: v for k, v in q.items() if k in tindex}], tindex)
            T = packer.pack_target_sets([distinct], tindex)
            ctx.set_targets_bits(T, len(distinct))
            h = ctx.score(P, Y, np.array([p], dtype=np.float64))[0]
            depth = first_pos[distinct[int(h["depth"]) - 1]] if int(h["depth"]) > 0 else 0
            return (float(h["score"]), int(h["yes"]), int(h["number"]), depth, p)
        P, Y, _ = packer.pack_queries([q], index)
        idx, raw, off = packer.pack_ranked_targets([lst], index, dup=self._dup)
        ctx.set_targets_ranked(idx, raw, off, len(index))
        ctx.set_option("dup", int(self._dup))  # ppp.pm:189-200, 222-224
        try:
            h = ctx.score(P, Y, np.array([p], dtype=np.float64))[0]
        finally:
            ctx.set_option("dup", 0)
        return (float(h["score"]), int(h["yes"]), int(h["number"]), int(h["depth"]), p)
