#!/usr/bin/env python
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.json by running the UNMODIFIED reference scorer
(/root/reference/lib/PhyloProf/Algorithm/ppp.pm:130-245 through oracle/run_reference.pl, with
Math::Cephes::bdtrc bound to the reference's own bundled binary) and, in the same pass, pins the C
restatement (oracle/libppp_oracle.so) against it: the script aborts on the first disagreement.

Run in the build container only (needs /root/reference, perl, `make -C oracle ref`):
    python oracle/gen_golden.py
The GPU box never runs this; tests read the committed JSON.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from prophylo_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def run_reference(cases):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        subprocess.run(["perl", os.path.join(ROOT, "oracle", "run_reference.pl"), tf.name], input=json.dumps(cases).encode(),
                       stdout=subprocess.DEVNULL, check=True)
        return json.load(open(tf.name))


def parse_profile(path):
    """2-column taxon<TAB>0/1 file as Profile::_parse_file reads it (Profile.pm:433-514), order preserved."""
    q = {}
    for line in open(path):
        line = line.rstrip("\n")
        if not line or line.startswith("#"):
            continue
        f = line.split("\t")
        if len(f) != 2:
            continue
        if f[0] in q and float(q[f[0]]) != 0:
            continue
        q[f[0]] = int(f[1])
    return q


def check(case, ref, what):
    got = oracle.score_case(case["query"], case["target"], case.get("prob"), bool(case.get("dup")))
    exp = (float(ref[0]), ref[1], ref[2], ref[3], float(ref[4]))
    if got != exp:
        raise SystemExit(f"ORACLE != REFERENCE on {what}: oracle {got} reference {exp}\ncase: {json.dumps(case)[:2000]}")


def known_answers():
    sulf = parse_profile(os.path.join(REF, "t", "sulf_mod.profile"))
    tl = ["gac_atcc", "gec5", "Xca", "gec5", "ntbc02", "NOT_IN_PROFILE", "nthc01", "ntms05", "arg", "ntpf02",
          "ntph02", "ntsa06", "ntte02", "bcl"]
    q6 = {"a": 1, "b": 1, "e": 1, "c": 0, "d": 0, "f": 0}
    cases = [
        {"name": "KA-1 t/sulf_mod.profile vs t/test.profile (BASELINE config 1)", "query": sulf, "target": {"1140": 1}},
        {"name": "KA-2", "query": sulf, "target": tl},
        {"name": "KA-3 prob=0.25", "query": sulf, "target": tl, "prob": 0.25},
        {"name": "KA-4 dup", "query": sulf, "target": tl, "dup": 1},
        {"name": "KA-5 prob=1.2 (domain error)", "query": sulf, "target": tl[:3], "prob": 1.2},
        {"name": "HASH target", "query": q6, "target": {"a": 1, "zz": 2, "b": 5, "d": 9, "e": 11}},
        {"name": "ARRAY target with dups and foreign taxa", "query": q6, "target": ["a", "zz", "a", "b", "zz", "d", "e"]},
        {"name": "empty target", "query": q6, "target": []},
        {"name": "all-miss target", "query": q6, "target": ["x1", "x2", "c", "d"]},
        {"name": "zero-yes query", "query": {"a": 0, "b": 0}, "target": ["a", "b"]},
        {"name": "all-yes query (p=1)", "query": {"a": 1, "b": 1}, "target": ["a", "b"]},
        {"name": "negative prob", "query": q6, "target": ["a", "b"], "prob": -0.5},
        {"name": "prob=1", "query": q6, "target": ["a", "b", "e"], "prob": 1},
    ]
    return cases


def fuzz_cases(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    cases = []
    for i in range(n):
        U = int(rng.integers(2, 70))
        taxa = [f"t{j}" for j in range(U)]
        npres = int(rng.integers(1, U + 1))
        pres = list(rng.permutation(U)[:npres])
        dens = float(rng.choice([0.05, 0.2, 0.5, 0.9]))
        query = {taxa[j]: int(rng.random() < dens) for j in pres}
        L = int(rng.integers(0, 2 * U))
        pool = taxa + [f"f{j}" for j in range(int(rng.integers(0, 6)))]
        mode = i % 5
        case = {"name": f"fuzz-{seed}-{i}", "query": query}
        if mode == 4:  # HASH target: unique taxa, distinct ranks
            k = min(L, len(pool))
            pick = list(rng.permutation(len(pool))[:k])
            ranks = rng.permutation(5 * k + 1)[:k] + 1
            case["target"] = {pool[int(j)]: int(r) for j, r in zip(pick, ranks)}
        elif mode == 3:  # target enriched in the query's yes taxa -> small p-values
            ys = [t for t, v in query.items() if v] or taxa[:1]
            tl = [ys[int(rng.integers(len(ys)))] if rng.random() < 0.7 else pool[int(rng.integers(len(pool)))] for _ in range(L)]
            case["target"] = tl
        else:
            if rng.random() < 0.5:  # permutation without dups
                tl = [pool[int(j)] for j in rng.permutation(len(pool))[:L]]
            else:
                tl = [pool[int(rng.integers(len(pool)))] for _ in range(L)]
            case["target"] = tl
        r = rng.random()
        if r < 0.15:
            case["prob"] = 0.3
        elif r < 0.2:
            case["prob"] = 1.2
        elif r < 0.3:
            case["prob"] = float(10 ** rng.uniform(-6, -1))
        elif r < 0.35:
            case["prob"] = 0.999
        if mode != 4 and rng.random() < 0.2:
            case["dup"] = 1
        cases.append(case)
    return cases


def bit_cases():
    """Canonical-order bit-plane pairs (SURVEY 8d embedding) at small and BASELINE genome counts; inputs are
    regenerated from (G, kind, seeds) by prophylo_b200.synth, only the expected outputs are stored."""
    specs = [
        dict(G=64, kind="mixed", nq=6, nt=24, qseed=11, tseed=12),
        dict(G=200, kind="family", nq=6, nt=20, qseed=21, tseed=22),
        dict(G=1024, kind="single", nq=1, nt=96, qseed=synth.SEED_BASE + 2, tseed=synth.SEED_BASE + 102),
        dict(G=1024, kind="mixed", nq=4, nt=24, qseed=31, tseed=32),
        dict(G=2048, kind="prefix", nq=6, nt=12, qseed=synth.SEED_BASE + 4, tseed=synth.SEED_BASE + 104),
        dict(G=4096, kind="mixed", nq=6, nt=12, qseed=synth.SEED_BASE + 3, tseed=synth.SEED_BASE + 103),
        dict(G=8192, kind="family", nq=3, nt=6, qseed=synth.SEED_BASE + 5, tseed=synth.SEED_BASE + 105),
    ]
    out = []
    for sp in specs:
        G = sp["G"]
        P, Y, p = synth.make_queries(G, sp["nq"], sp["qseed"], sp["kind"])
        T = synth.make_targets(G, sp["nt"], sp["tseed"], plant_from=Y, plant_frac=0.25)
        Pb, Yb, Tb = synth.unpack_bits(P, G), synth.unpack_bits(Y, G), synth.unpack_bits(T, G)
        cases = []
        for q in range(sp["nq"]):
            query = {str(g + 1): int(Yb[q, g]) for g in range(G) if Pb[q, g]}
            for t in range(sp["nt"]):
                cases.append({"query": query, "target": [str(g + 1) for g in np.nonzero(Tb[t])[0]], "prob": float(p[q])})
        ref = run_reference(cases)
        res = []
        for q in range(sp["nq"]):
            for t in range(sp["nt"]):
                r = ref[q * sp["nt"] + t]
                got = oracle.score_bits(P[q], Y[q], T[t], G, float(p[q]))[:5]
                exp = (float(r[0]), r[1], r[2], r[3], float(r[4]))
                if got != exp:
                    raise SystemExit(f"ORACLE(bits) != REFERENCE at {sp} q={q} t={t}: {got} vs {exp}")
                res.append([r[0], r[1], r[2], r[3]])
        out.append({"spec": sp, "p": [repr(float(x)) for x in p], "results": res})
        print(f"bits G={G} kind={sp['kind']}: {len(res)} pairs pinned", file=sys.stderr)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    ka = known_answers()
    fz = fuzz_cases(600, 1234)
    ref = run_reference(ka + fz)
    for c, r in zip(ka + fz, ref):
        if isinstance(r, dict):
            raise SystemExit(f"reference died on {c['name']}: {r}")
        check(c, r, c["name"])
        c["expect"] = r[:5]
        c["query_yes_total"] = r[5:7]
    json.dump({"generator": "oracle/gen_golden.py", "reference": "lib/PhyloProf/Algorithm/ppp.pm:130-245 via oracle/run_reference.pl",
               "known_answers": ka, "fuzz": fz}, open(os.path.join(GOLD, "ppp_list_cases.json"), "w"), separators=(",", ":"))
    print(f"list form: {len(ka)} known-answer + {len(fz)} fuzz cases pinned (oracle == reference)", file=sys.stderr)
    bits = bit_cases()
    json.dump({"generator": "oracle/gen_golden.py", "cases": bits}, open(os.path.join(GOLD, "ppp_bit_cases.json"), "w"),
              separators=(",", ":"))
    # bdtrc points straight from the reference binary (dlopen), incl. SURVEY 8c's table
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "Cephes.so"), mode=os.RTLD_LAZY)
    lib.bdtrc.restype = ctypes.c_double
    lib.bdtrc.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
    rng = np.random.Generator(np.random.PCG64(99))
    pts = [(0, 1, .02), (0, 5, .0207), (1, 5, .0207), (8, 12, 9 / 434), (3, 100, .25), (49, 50, .5), (99, 4096, .01),
           (500, 4096, .1), (-1, 7, .3), (5, 5, .3), (150, 4096, 1e-4), (0, 5, .005), (0, 5, .01), (0, 8000, 0.0099),
           (3, 2, .5), (1, 3, 0.0), (1, 3, 1.0), (0, 3, 1.0), (2, 8192, 2 / 8192)]
    for _ in range(3000):
        n = int(rng.integers(1, 8193))
        p = float(rng.choice([rng.uniform(0, 1), 10 ** rng.uniform(-5, 0), min(1.0, rng.integers(1, n + 1) / n)]))
        k = int(rng.choice([rng.integers(0, n + 1), min(n, max(0, int(n * p + rng.normal() * (1 + 3 * np.sqrt(n * p * (1 - p)))))), 0]))
        pts.append((k, n, p))
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the binary's mtherr prints on stdout
    vals = [lib.bdtrc(k, n, p) for k, n, p in pts]
    os.dup2(saved, 1)
    for (k, n, p), v in zip(pts, vals):
        # k==0 and p<0.01 goes through the process' libm expm1 (see cephes_port.c); same libm here
        if oracle.bdtrc(k, n, p) != v:
            raise SystemExit(f"cp_bdtrc != binary at {(k, n, p)}")
    json.dump({"generator": "oracle/gen_golden.py", "source": "bdtrc symbol of the reference's bundled Math::Cephes 0.47 Cephes.so",
               "points": [[k, n, repr(p), repr(v)] for (k, n, p), v in zip(pts, vals)]},
              open(os.path.join(GOLD, "bdtrc_points.json"), "w"), separators=(",", ":"))
    print(f"bdtrc: {len(pts)} points pinned", file=sys.stderr)


if __name__ == "__main__":
    main()
