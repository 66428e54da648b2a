#!/usr/bin/env python
"""oracle/gen_golden_cross.py -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference driver bin/ppp_cross_score.pl (BLAST x BLAST mode, through the oracle's perl stubs and the
reference's bundled Cephes binary) on small synthetic BLAST -m 8 tables and stores inputs + stdout in
tests/golden/ppp_cross_score.json.  Needs /root/reference; the GPU box only reads the JSON."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def make_table(rng, query, taxa_pool, gi_base, nrows, shared):
    """BLAST -m 8 rows: 12 tab columns, subject 'gi|N|ref|X|', bit score in column 12; some subjects repeat (HSPs), some
    gis have no taxon (skipped by the reference), several gis share a taxon."""
    rows, gi2tax = [], {}
    for i in range(nrows):
        gi = gi_base + int(rng.integers(0, max(4, nrows * 2 // 3)))
        if gi not in gi2tax and rng.random() > 0.08:
            pool = shared if rng.random() < 0.5 else taxa_pool
            gi2tax[gi] = str(rng.choice(pool))
        bits = 900.0 - 7.5 * i + float(rng.integers(0, 5))
        rows.append("\t".join([query, f"gi|{gi}|ref|XP_{gi}.1|", "%.2f" % rng.uniform(25, 100), str(int(rng.integers(50, 400))), "3", "1", "1",
                               "200", "5", "210", "%.0e" % 10.0 ** -float(rng.integers(3, 150)), "%.1f" % bits]))
    return "# BLASTP\n" + "\n".join(rows) + "\n", gi2tax


def make_case(seed, n1, n2, ntaxa, start=1, end=None, skip=1, scale=None):
    rng = np.random.default_rng(seed)
    taxa = [str(100 + 3 * i) for i in range(ntaxa)]
    shared = list(rng.choice(taxa, max(3, ntaxa // 6), replace=False))
    t1, m1 = make_table(rng, "q1", taxa, 10_000, n1, shared)
    t2, m2 = make_table(rng, "q2", taxa, 20_000, n2, shared)
    dist = "# taxon distribution\n" + "".join(f"{t}\n" for t in taxa)
    return dict(seed=seed, file1=t1, file2=t2, gi2taxid={str(k): v for k, v in {**m1, **m2}.items()}, dist=dist,
                start=start, end=end, skip=skip, scale=scale)


def run_reference(case):
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(f"{d}/sdb/taxonomy")
        for f in ("gi_taxid_prot.db", "nodes.db"):
            open(f"{d}/sdb/taxonomy/{f}", "w").write("x\n")
        open(f"{d}/a.m8", "w").write(case["file1"])
        open(f"{d}/b.m8", "w").write(case["file2"])
        open(f"{d}/dist.txt", "w").write(case["dist"])
        open(f"{d}/gi2taxid.tsv", "w").write("".join(f"{g}\t{t}\n" for g, t in case["gi2taxid"].items()))
        env = dict(os.environ, SEQTOOLBOXDB=f"{d}/sdb", CEPHES_SHIM_SO=f"{ROOT}/oracle/_ref/Cephes_shim.so",
                   CEPHES_REAL_SO=f"{ROOT}/oracle/_ref/Cephes.so", PERL5LIB=f"{ROOT}/oracle/stubs:{REF}/lib",
                   PPP_STUB_GI2TAXID=f"{d}/gi2taxid.tsv")
        cmd = ["perl", f"{REF}/bin/ppp_cross_score.pl", "--file1", "a.m8", "--file2", "b.m8", "--method1", "blast", "--method2", "blast",
               "--dist-file", "dist.txt", "--start-depth", str(case["start"]), "--skip", str(case["skip"])]
        if case["end"]:
            cmd += ["--end-depth", str(case["end"])]
        if case["scale"]:
            cmd += ["--scale", str(case["scale"])]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=d)
        if r.returncode != 0:
            raise SystemExit(f"reference failed (seed {case['seed']}): {r.stderr[-2000:]}")
        case["stdout"] = r.stdout
    return case


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    cases = [run_reference(make_case(11, 40, 55, 90)), run_reference(make_case(12, 120, 80, 300, start=5)),
             run_reference(make_case(13, 60, 60, 150, start=3, end=25, skip=2)), run_reference(make_case(14, 200, 150, 1459, scale=0.5)),
             run_reference(make_case(15, 30, 25, 40, start=2))]
    json.dump({"generator": "oracle/gen_golden_cross.py", "reference": "bin/ppp_cross_score.pl --method1 blast --method2 blast", "cases": cases},
              open(os.path.join(ROOT, "tests", "golden", "ppp_cross_score.json"), "w"), separators=(",", ":"))
    for c in cases:
        print(c["stdout"], file=sys.stderr)


if __name__ == "__main__":
    main()
