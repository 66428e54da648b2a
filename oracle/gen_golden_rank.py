#!/usr/bin/env python
"""oracle/gen_golden_rank.py -- TEST INFRASTRUCTURE ONLY.

`-rank` cases: the UNMODIFIED reference scorer (lib/PhyloProf/Algorithm/ppp.pm:130-245, its
_get_taxon :328-370) run through oracle/run_reference.pl with a mock -taxonomy object whose collapse_taxon(id, rank) is a
table, and -- in the same pass -- the oracle's restatement of that host pre-pass pinned against it.  Also the croak of
ppp.pm:345-349 when -rank is given without a taxonomy.  Writes tests/golden/ppp_rank_cases.json.
Needs /root/reference; the GPU box only reads the JSON."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from oracle.gen_golden import run_reference  # noqa: E402


def make_case(rng, i):
    nsp = int(rng.integers(10, 60))  # species ids
    species = [str(5000 + j) for j in range(nsp)]
    nfam = max(2, nsp // 4)
    fam = {s: str(900 + int(rng.integers(nfam))) for s in species if rng.random() < 0.8}       # rank table (some species missing)
    genus = {s: str(700 + int(rng.integers(2 * nfam))) for s in species if rng.random() < 0.6}  # fallback rank (ppp.pm:354-358)
    # the query is expressed in collapsed ids (the algorithm only collapses the TARGET's taxa, ppp.pm:175)
    coll = sorted({fam.get(s) or genus.get(s) or s for s in species})
    dens = float(rng.choice([0.2, 0.5, 0.8]))
    query = {t: int(rng.random() < dens) for t in coll if rng.random() < 0.9}
    if not query:
        query = {coll[0]: 1}
    L = int(rng.integers(1, 2 * nsp))
    pool = species + [str(8000 + j) for j in range(3)]  # species unknown to every table stay raw ids
    target = [pool[int(rng.integers(len(pool)))] for _ in range(L)]
    case = {"name": f"rank-{i}", "query": query, "target": target, "rank": "family", "collapse": {"family": fam, "genus": genus}}
    if i % 4 == 3:  # HASH target (HMMER style): distinct species with distinct ranks
        k = min(L, len(pool))
        pick = list(rng.permutation(len(pool))[:k])
        ranks = rng.permutation(5 * k + 1)[:k] + 1
        case["target"] = {pool[int(j)]: int(r) for j, r in zip(pick, ranks)}
    if i % 5 == 0:
        case["prob"] = 0.3
    if i % 7 == 0 and not isinstance(case["target"], dict):
        case["dup"] = 1
    return case


def main():
    rng = np.random.Generator(np.random.PCG64(4242))
    cases = [make_case(rng, i) for i in range(120)]
    nomap = {"name": "rank without taxonomy croaks", "query": {"a": 1, "b": 0}, "target": ["a", "b"], "rank": "family"}
    ref = run_reference(cases + [nomap])
    for c, r in zip(cases, ref):
        if isinstance(r, dict):
            raise SystemExit(f"reference died on {c['name']}: {r}")
        got = oracle.score_case(c["query"], c["target"], c.get("prob"), bool(c.get("dup")), rank=c["rank"], collapse=c["collapse"])
        exp = (float(r[0]), r[1], r[2], r[3], float(r[4]))
        if got != exp:
            raise SystemExit(f"ORACLE != REFERENCE on {c['name']}: {got} vs {exp}")
        c["expect"] = r[:5]
    assert isinstance(ref[-1], dict) and "Taxonomy does not exist" in ref[-1]["error"], ref[-1]
    nomap["error"] = "Taxonomy does not exist"  # the message of ppp.pm:348 (croak adds the caller's file and line)
    changed = sum(1 for c in cases if oracle.score_case(c["query"], c["target"], c.get("prob"), bool(c.get("dup")))[:4]
                  != (float(c["expect"][0]), c["expect"][1], c["expect"][2], c["expect"][3]))
    json.dump({"generator": "oracle/gen_golden_rank.py", "reference": "lib/PhyloProf/Algorithm/ppp.pm:130-245, 328-370 via oracle/run_reference.pl "
               "with a mock -taxonomy object", "cases": cases, "croak": nomap},
              open(os.path.join(ROOT, "tests", "golden", "ppp_rank_cases.json"), "w"), separators=(",", ":"))
    print(f"{len(cases)} -rank cases pinned (oracle == reference); the collapse changes the result in {changed} of them; "
          f"croak: {nomap['error']!r}", file=sys.stderr)


if __name__ == "__main__":
    main()
