#!/usr/bin/env python
"""oracle/gen_golden_e2e.py -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference drivers bin/ppp.pl (server + forked clients, through the oracle's perl stubs and the
reference's bundled Cephes binary) and bin/ppp_cutoff.pl on small synthetic BLAST caches and stores inputs + stdout
in tests/golden/ppp_pl_e2e.json.  Needs /root/reference; the GPU box only reads the JSON."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def make_nodes(taxa, rng):
    """NCBI-style nodes rows (tax_id, parent, rank) for species ids: species -> genus -> family -> root 1; one species in
    ten has no row at all (collapse_taxon returns undef: the scorer falls back to genus, then to the raw id), one genus in
    eight hangs directly under the root (no family: the genus fallback of ppp.pm:354-358 applies)."""
    rows = {}
    for i, t in enumerate(taxa):
        if rng.random() < 0.1:
            continue
        genus, family = str(50000 + i // 3), str(70000 + i // 9)
        rows[t] = (genus, "species")
        rows[genus] = ("1", "genus") if (i // 3) % 8 == 0 else (family, "genus")
        if (i // 3) % 8:
            rows[family] = ("1", "family")
    return "".join(f"{t}\t{p}\t{r}\n" for t, (p, r) in rows.items())


def make_case(seed, ngenes, ntaxa, prob=None, slope=None, dup=False, level=None):
    rng = np.random.default_rng(seed)
    taxa = [str(1000 + i) for i in range(ntaxa)]
    in_profile = taxa[: int(ntaxa * 0.85)]
    yes = set(rng.choice(in_profile, max(2, ntaxa // 8), replace=False))
    profile = "".join(f"{t}\t{1 if t in yes else 0}\n" for t in in_profile)
    bla = {}
    for g in range(ngenes):
        gi = str(500000 + g)
        L = int(rng.integers(0, 120))
        rows = [f"{gi}\t{gi}\t0.0\t1140\n"] if g % 7 else []
        for _ in range(L):
            t = str(rng.choice(sorted(yes))) if (g % 4 == 0 and rng.random() < 0.6) else str(rng.choice(taxa + ["0"]))
            s = 700000 + int(rng.integers(0, 60))
            rows.append(f"{gi}\t{s}\t{10.0 ** -float(rng.integers(1, 100)):.3g}\t{t}\n")
        if not rows:
            rows = [f"{gi}\t{gi}\t0.0\t1140\n"]
        bla[gi] = "".join(rows)
    case = dict(seed=seed, profile=profile, bla=bla, prob=prob, slope=slope, dup=bool(dup))
    if level:
        case["level"], case["nodes"] = level, make_nodes(taxa, rng)
    return case


def run_reference(case):
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(f"{d}/db/blast/1140"); os.makedirs(f"{d}/db/desc"); os.makedirs(f"{d}/ws"); os.makedirs(f"{d}/sdb/taxonomy")
        for f in ("gi_taxid_prot.db", "nodes.db"):
            open(f"{d}/sdb/taxonomy/{f}", "w").write("x\n")
        open(f"{d}/db/desc/gi_desc.sqlite3", "w").close()
        open(f"{d}/q.profile", "w").write(case["profile"])
        for gi, txt in case["bla"].items():
            open(f"{d}/db/blast/1140/{gi}.bla", "w").write(txt)
        env = dict(os.environ, SEQTOOLBOXDB=f"{d}/sdb", CEPHES_SHIM_SO=f"{ROOT}/oracle/_ref/Cephes_shim.so",
                   CEPHES_REAL_SO=f"{ROOT}/oracle/_ref/Cephes.so", PERL5LIB=f"{ROOT}/oracle/stubs:{REF}/lib")
        cmd = ["perl", f"{REF}/bin/ppp.pl", "-d", f"{d}/db", "-t", "1140", "-p", f"{d}/q.profile", "-w", f"{d}/ws", "--threads", "2"]
        if case.get("level"):  # -l <rank>: SeqToolBox::Taxonomy::collapse_taxon answered by the DBI stub's nodes table
            open(f"{d}/nodes.tsv", "w").write(case["nodes"])
            env["PPP_STUB_NODES"] = f"{d}/nodes.tsv"
            cmd += ["-l", case["level"]]
        if case["prob"]:
            cmd += ["--prob", str(case["prob"])]
        if case["slope"]:
            cmd += ["-m", str(case["slope"])]
        if case.get("dup"):
            cmd += ["--dup"]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True)
        case["ppp_pl_stdout"] = r.stdout
        case["cutoff"] = {}
        open(f"{d}/table.txt", "w").write(r.stdout)
        for s in (10, 20, 50):
            c = subprocess.run(["perl", f"{REF}/bin/ppp_cutoff.pl", "-s", str(s), f"{d}/table.txt"], capture_output=True, text=True, check=True)
            case["cutoff"][str(s)] = {"stdout": c.stdout, "stderr": c.stderr}
    return case


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    if "--level" in sys.argv:  # -l family (bin/ppp.pl:532-547, 566-568 -> Profile.pm:516-549, ppp.pm:328-370), its own fixture file
        cases = [run_reference(make_case(31, 36, 90, slope=30, level="family")), run_reference(make_case(32, 40, 240, prob=0.06, level="family")),
                 run_reference(make_case(33, 30, 120, level="genus", dup=True))]
        name = "ppp_pl_e2e_level.json"
    elif "--dup" in sys.argv:  # the -dup flag of bin/ppp.pl (:282 -> ppp.pm:189-200), its own fixture file
        cases = [run_reference(make_case(11, 40, 80, slope=30, dup=True)), run_reference(make_case(12, 50, 250, prob=0.08, dup=True))]
        name = "ppp_pl_e2e_dup.json"
    else:
        cases = [run_reference(make_case(1, 30, 60, slope=30)), run_reference(make_case(2, 45, 200)),
                 run_reference(make_case(3, 25, 90, prob=0.05, slope=15)), run_reference(make_case(4, 60, 400, slope=40))]
        name = "ppp_pl_e2e.json"
    json.dump({"generator": "oracle/gen_golden_e2e.py" + (" --level" if "--level" in sys.argv else " --dup" if "--dup" in sys.argv else ""),
               "reference": "bin/ppp.pl (server+clients), bin/ppp_cutoff.pl", "cases": cases},
              open(os.path.join(ROOT, "tests", "golden", name), "w"), separators=(",", ":"))
    print(f"{len(cases)} end-to-end cases; rows: {[c['ppp_pl_stdout'].count(chr(10)) for c in cases]}", file=sys.stderr)


if __name__ == "__main__":
    main()
