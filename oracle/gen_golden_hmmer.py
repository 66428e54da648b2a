#!/usr/bin/env python
"""oracle/gen_golden_hmmer.py -- TEST INFRASTRUCTURE ONLY.

HMMER-side goldens from the UNMODIFIED reference drivers, run through the oracle's perl stubs (SeqToolBox.pm, DBI.pm serving
a gi -> tax_id table, an empty Bio::SearchIO) with the reference's own SeqToolBox::HMMER::Parser and PhyloProf::HMMERResult:

  * bin/ppp_cross_score.pl with --method hmmer on either or both sides (_evaluate_hmmer, bin/ppp_cross_score.pl:549-679;
    get_target_profile_hmmer :326-348), incl. --trusted_cutoff, --start-depth / --end-depth / --skip / --scale;
  * bin/ppp_hmmer.pl (one query profile against one hmmsearch result, :179-226), full-sequence and --domain scores,
    --trusted_cutoff, --prob.

Inputs are synthetic hmmsearch 3 reports (the per-sequence table the parser reads, lib/SeqToolBox/HMMER/Parser.pm:358-422)
and BLAST -m 8 tables; scores are distinct, because the reference orders equal scores by Perl's hash order.
Writes tests/golden/ppp_cross_score_hmmer.json and tests/golden/ppp_hmmer.json.  Needs /root/reference."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
from oracle.gen_golden_cross import make_table  # noqa: E402


def make_hmmsearch(rng, name, taxa_pool, gi_base, nrows, shared):
    """A hmmsearch 3 report with `nrows` per-sequence rows (descending, distinct full and best-domain scores), an inclusion
    threshold line, blank lines inside the table, descriptions with several words; some gis have no taxon, several share one."""
    full = np.sort(rng.choice(np.arange(50, 9000), nrows, replace=False))[::-1] / 10.0
    dom = np.sort(rng.choice(np.arange(40, 8900), nrows, replace=False))[::-1] / 10.0
    perm = rng.permutation(nrows)  # best-domain order differs from full-sequence order
    gi2tax, rows = {}, []
    for i in range(nrows):
        gi = gi_base + i
        if rng.random() > 0.08:
            pool = shared if rng.random() < 0.5 else taxa_pool
            gi2tax[gi] = str(rng.choice(pool))
        d = dom[perm[i]] if i % 3 else min(dom[perm[i]], full[i])
        seq = f"gi|{gi}|ref|YP_{gi}.1|"
        rows.append(f"  {10.0 ** -float(full[i] / 3):9.2g} {full[i]:6.1f} {rng.uniform(0, 2):5.1f}  {10.0 ** -float(d / 3):9.2g} {d:6.1f} "
                    f"{rng.uniform(0, 2):5.1f}  {rng.uniform(1, 2.5):4.1f} {int(rng.integers(1, 3)):2d}  {seq:<36s} "
                    f"hypothetical protein X{gi} [Genus species {i}]")
    cut = nrows * 2 // 3
    body = rows[:cut] + ["  ------ inclusion threshold ------"] + rows[cut:cut + 2] + [""] + rows[cut + 2:]
    text = (f"# hmmsearch :: search profile HMM(s) against a sequence database\n# HMMER 3.0 (synthetic)\n\nQuery:       {name}  [M=117]\n"
            f"Accession:   {name}\nDescription: synthetic family {name}\nScores for complete sequences (score includes all domains):\n"
            "   --- full sequence ---   --- best 1 domain ---    -#dom-\n"
            "    E-value  score  bias    E-value  score  bias    exp  N  Sequence                             Description\n"
            "    ------- ------ -----    ------- ------ -----   ---- --  --------                             -----------\n"
            + "\n".join(body) + "\n\n\nDomain annotation for each sequence (and alignments):\n>> not parsed\n")
    return text, gi2tax


def env_for(d, gi2taxid):
    os.makedirs(f"{d}/sdb/taxonomy", exist_ok=True)
    for f in ("gi_taxid_prot.db", "nodes.db"):
        open(f"{d}/sdb/taxonomy/{f}", "w").write("x\n")
    open(f"{d}/gi2taxid.tsv", "w").write("".join(f"{g}\t{t}\n" for g, t in gi2taxid.items()))
    return dict(os.environ, SEQTOOLBOXDB=f"{d}/sdb", CEPHES_SHIM_SO=f"{ROOT}/oracle/_ref/Cephes_shim.so",
                CEPHES_REAL_SO=f"{ROOT}/oracle/_ref/Cephes.so", PERL5LIB=f"{ROOT}/oracle/stubs:{REF}/lib",
                PPP_STUB_GI2TAXID=f"{d}/gi2taxid.tsv", PERL_HASH_SEED="3")


def cross_case(seed, m1, m2, n1, n2, ntaxa, start=1, end=None, skip=1, scale=None, cutoff=None):
    rng = np.random.default_rng(seed)
    taxa = [str(100 + 3 * i) for i in range(ntaxa)]
    shared = list(rng.choice(taxa, max(3, ntaxa // 6), replace=False))
    mk = lambda m, q, base, n: (make_hmmsearch(rng, q, taxa, base, n, shared) if m == "hmmer" else make_table(rng, q, taxa, base, n, shared))  # noqa: E731
    t1, g1 = mk(m1, "FAM1", 10_000, n1)
    t2, g2 = mk(m2, "FAM2", 20_000, n2)
    case = dict(seed=seed, method1=m1, method2=m2, file1=t1, file2=t2, gi2taxid={str(k): v for k, v in {**g1, **g2}.items()},
                dist="# taxon distribution\n" + "".join(f"{t}\n" for t in taxa), start=start, end=end, skip=skip, scale=scale, cutoff=cutoff)
    with tempfile.TemporaryDirectory() as d:
        open(f"{d}/a.out", "w").write(t1)
        open(f"{d}/b.out", "w").write(t2)
        open(f"{d}/dist.txt", "w").write(case["dist"])
        cmd = ["perl", f"{REF}/bin/ppp_cross_score.pl", "--file1", "a.out", "--file2", "b.out", "--method1", m1, "--method2", m2,
               "--dist-file", "dist.txt", "--start-depth", str(start), "--skip", str(skip)]
        if end:
            cmd += ["--end-depth", str(end)]
        if scale:
            cmd += ["--scale", str(scale)]
        if cutoff:
            cmd += ["--trusted_cutoff", str(cutoff)]
        r = subprocess.run(cmd, env=env_for(d, case["gi2taxid"]), capture_output=True, text=True, cwd=d)
        if r.returncode != 0:
            raise SystemExit(f"ppp_cross_score.pl failed (seed {seed}): {r.stderr[-2000:]}")
        case["stdout"] = r.stdout
    return case


def hmmer_case(seed, nrows, ntaxa, domain=False, cutoff=None, prob=None):
    rng = np.random.default_rng(seed)
    taxa = [str(100 + 3 * i) for i in range(ntaxa)]
    shared = list(rng.choice(taxa, max(3, ntaxa // 5), replace=False))
    text, g = make_hmmsearch(rng, "FAMQ", taxa, 30_000, nrows, shared)
    in_profile = taxa[: int(ntaxa * 0.9)]
    profile = "".join(f"{t}\t{1 if (t in shared and rng.random() < 0.8) or rng.random() < 0.05 else 0}\n" for t in in_profile)
    case = dict(seed=seed, hmm=text, profile=profile, gi2taxid={str(k): v for k, v in g.items()}, domain=bool(domain), cutoff=cutoff, prob=prob)
    with tempfile.TemporaryDirectory() as d:
        open(f"{d}/q.profile", "w").write(profile)
        open(f"{d}/r.out", "w").write(text)
        cmd = ["perl", f"{REF}/bin/ppp_hmmer.pl", "-p", "q.profile", "--hmm", "r.out"]
        if domain:
            cmd += ["--domain"]
        if cutoff:
            cmd += ["-t", str(cutoff)]
        if prob:
            cmd += ["--prob", str(prob)]
        r = subprocess.run(cmd, env=env_for(d, case["gi2taxid"]), capture_output=True, text=True, cwd=d)
        if r.returncode != 0:
            raise SystemExit(f"ppp_hmmer.pl failed (seed {seed}): {r.stderr[-2000:]}")
        case["stdout"] = r.stdout
    return case


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    cross = [cross_case(41, "hmmer", "hmmer", 60, 70, 120), cross_case(42, "blast", "hmmer", 80, 90, 300, start=3),
             cross_case(43, "hmmer", "blast", 100, 60, 200, start=2, end=30, skip=2), cross_case(44, "hmmer", "hmmer", 150, 120, 1459, scale=0.5, cutoff=25),
             cross_case(45, "hmmer", "blast", 40, 50, 60, cutoff=100)]
    json.dump({"generator": "oracle/gen_golden_hmmer.py", "reference": "bin/ppp_cross_score.pl with --method hmmer", "cases": cross},
              open(os.path.join(ROOT, "tests", "golden", "ppp_cross_score_hmmer.json"), "w"), separators=(",", ":"))
    for c in cross:
        print(c["seed"], c["method1"], c["method2"], repr(c["stdout"]), file=sys.stderr)
    hm = [hmmer_case(51, 60, 100), hmmer_case(52, 90, 200, domain=True), hmmer_case(53, 120, 300, cutoff=30), hmmer_case(54, 80, 150, domain=True, cutoff=50, prob=0.1),
          hmmer_case(55, 40, 60, prob=0.3)]
    json.dump({"generator": "oracle/gen_golden_hmmer.py", "reference": "bin/ppp_hmmer.pl", "cases": hm},
              open(os.path.join(ROOT, "tests", "golden", "ppp_hmmer.json"), "w"), separators=(",", ":"))
    for c in hm:
        print(c["seed"], repr(c["stdout"][:400]), file=sys.stderr)


if __name__ == "__main__":
    main()
