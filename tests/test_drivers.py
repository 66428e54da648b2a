"""Either side of the hot path: the `.bla` reader, the ppp.pl stdout table and the ppp_cutoff.pl
slope filter, against golden outputs of the UNMODIFIED reference drivers (oracle/gen_golden_e2e.py ran bin/ppp.pl with
forked clients and bin/ppp_cutoff.pl in the build container)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from prophylo_b200 import drivers
from prophylo_b200.algorithm import Profile


@pytest.fixture(scope="module")
def e2e(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ppp_pl_e2e.json")))["cases"]


@pytest.fixture(scope="module")
def e2e_dup(golden_dir):
    """bin/ppp.pl --dup (:282 -> ppp.pm:189-200) run unmodified on the same kind of caches (gen_golden_e2e.py --dup)"""
    return json.load(open(os.path.join(golden_dir, "ppp_pl_e2e_dup.json")))["cases"]


@pytest.fixture(scope="module")
def e2e_level(golden_dir):
    """bin/ppp.pl -l <rank> run unmodified, SeqToolBox::Taxonomy::collapse_taxon answered from a synthetic NCBI nodes table
    (gen_golden_e2e.py --level)"""
    return json.load(open(os.path.join(golden_dir, "ppp_pl_e2e_level.json")))["cases"]


def nodes_taxonomy(case):
    from prophylo_b200.algorithm import NodesTaxonomy

    rows = {}
    for line in case["nodes"].split("\n"):
        f = line.split("\t")
        if len(f) == 3:
            rows[f[0]] = (f[1], f[2])
    return NodesTaxonomy(rows)


def materialise(case, tmp_path):
    d = tmp_path / f"case{case['seed']}"
    (d / "bla").mkdir(parents=True)
    (d / "q.profile").write_text(case["profile"])
    paths = {}
    for gi, txt in case["bla"].items():
        p = d / "bla" / f"{gi}.bla"
        p.write_text(txt)
        paths[gi] = str(p)
    return str(d / "q.profile"), paths


def normalise(table):
    """The reference orders rows with equal raw score by Perl hash order (bin/ppp.pl:473-474): compare tie groups as
    sets -- rows are grouped by their printed Score column and sorted inside a group; marker lines stay in place."""
    lines = table.rstrip("\n").split("\n")
    out, group, key = [lines[0]], [], None
    for ln in lines[1:]:
        k = ln.split("\t")[5] if "\t" in ln else ln
        if k != key and group:
            out += sorted(group)
            group = []
        key = k
        group.append(ln)
    return out + sorted(group)


def test_ppp_cutoff_matches_reference_script(e2e):
    n = 0
    for case in e2e:
        for slope, ref in case["cutoff"].items():
            text, cutoff = drivers.ppp_cutoff(case["ppp_pl_stdout"], float(slope))
            assert text == ref["stdout"], (case["seed"], slope)
            exp = [ln for ln in ref["stderr"].split("\n") if ln.startswith("CUTOFF=")][-1]
            if exp == "CUTOFF=":
                assert cutoff is None
            else:
                assert float(exp[7:]) == cutoff
            n += 1
    assert n == 12


def test_table_format_matches_reference_with_oracle_scores(e2e, tmp_path):
    """CPU-only check of the reader + formatter: rows scored by the oracle, table must equal ppp.pl's stdout."""
    for case in e2e:
        prof, paths = materialise(case, tmp_path)
        q = Profile(file=prof)
        qd = {k: int(v) for k, v in q.get_hash().items()}
        rows = []
        for gi, path in paths.items():
            lst = drivers.read_bla(path)
            s, y, n, d, p = oracle.score_case(qd, lst, case["prob"])
            rows.append((gi, s, y, n, d, p))
        got = drivers.format_ppp_table(rows, slope=case["slope"])
        assert normalise(got) == normalise(case["ppp_pl_stdout"]), case["seed"]


def test_dup_table_matches_reference_with_oracle_scores(e2e_dup, tmp_path):
    """CPU: the oracle's -dup loop reproduces the stdout of the reference's `ppp.pl --dup`, and the flag matters here."""
    for case in e2e_dup:
        prof, paths = materialise(case, tmp_path)
        qd = {k: int(v) for k, v in Profile(file=prof).get_hash().items()}
        rows, plain = [], []
        for gi, path in paths.items():
            lst = drivers.read_bla(path)
            rows.append((gi,) + oracle.score_case(qd, lst, case["prob"], True))
            plain.append((gi,) + oracle.score_case(qd, lst, case["prob"], False))
        assert normalise(drivers.format_ppp_table(rows, slope=case["slope"])) == normalise(case["ppp_pl_stdout"]), case["seed"]
        assert normalise(drivers.format_ppp_table(plain, slope=case["slope"])) != normalise(case["ppp_pl_stdout"])


def test_level_table_matches_reference_with_oracle_scores(e2e_level, tmp_path):
    """CPU: `ppp.pl -l <rank>` -- the query profile and every hit list collapsed through the taxonomy (Profile.pm:516-549,
    ppp.pm:328-370: the rank, else genus, else the raw id), scored by the oracle -- equals the unmodified script's stdout."""
    from prophylo_b200.algorithm import collapse

    for case in e2e_level:
        prof, paths = materialise(case, tmp_path)
        tax = nodes_taxonomy(case)
        q = Profile(file=prof, rank=case["level"], taxonomy=tax)
        qd = {k: int(v) for k, v in q.get_hash().items()}
        rows, raw_rows = [], []
        for gi, path in paths.items():
            lst = drivers.read_bla(path)
            rows.append((gi,) + oracle.score_case(qd, [collapse(tax, case["level"], t) for t in lst], case["prob"], case["dup"]))
            raw_rows.append((gi,) + oracle.score_case(qd, lst, case["prob"], case["dup"]))
        assert normalise(drivers.format_ppp_table(rows, slope=case["slope"])) == normalise(case["ppp_pl_stdout"]), case["seed"]
        assert normalise(drivers.format_ppp_table(raw_rows, slope=case["slope"])) != normalise(case["ppp_pl_stdout"])  # the collapse matters


@pytest.mark.gpu
def test_level_scan_on_gpu_and_perl_cli_reproduce_ppp_pl_stdout(e2e_level, tmp_path):
    """-l <rank> through both hosts: drivers.ppp_scan(level=, taxonomy=) and `perl/bin/ppp_b200.pl -l <rank>` (with a test stub
    of SeqToolBox::Taxonomy on PERL5LIB walking the same nodes table) against the unmodified reference script's stdout."""
    import subprocess

    from prophylo_b200 import capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "perl")], stdout=subprocess.DEVNULL)
    ctx = capi.Context(0)
    try:
        for case in e2e_level:
            prof, paths = materialise(case, tmp_path)
            tax = nodes_taxonomy(case)
            q = Profile(file=prof, rank=case["level"], taxonomy=tax)
            rows = drivers.ppp_scan(ctx, q, paths, prob=case["prob"], dup=case["dup"], level=case["level"], taxonomy=tax)
            assert normalise(drivers.format_ppp_table(rows, slope=case["slope"])) == normalise(case["ppp_pl_stdout"]), case["seed"]
            d = tmp_path / f"lvl{case['seed']}"
            (d / "db" / "blast" / "1140").mkdir(parents=True)
            (d / "ws").mkdir()
            (d / "q.profile").write_text(case["profile"])
            (d / "nodes.tsv").write_text(case["nodes"])
            for gi, txt in case["bla"].items():
                (d / "db" / "blast" / "1140" / f"{gi}.bla").write_text(txt)
            cmd = ["perl", os.path.join(root, "perl", "bin", "ppp_b200.pl"), "-d", str(d / "db"), "-t", "1140", "-p", str(d / "q.profile"),
                   "-w", str(d / "ws"), "-l", case["level"]]
            if case["prob"]:
                cmd += ["--prob", str(case["prob"])]
            if case["slope"]:
                cmd += ["-m", str(case["slope"])]
            if case["dup"]:
                cmd += ["--dup"]
            env = dict(os.environ, PERL5LIB=os.path.join(root, "perl", "t", "lib"), PPP_MOCK_NODES=str(d / "nodes.tsv"))
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
            assert r.returncode == 0, r.stderr
            assert normalise(r.stdout) == normalise(case["ppp_pl_stdout"]), case["seed"]
    finally:
        ctx.close()


@pytest.mark.gpu
def test_ppp_scan_on_gpu_reproduces_ppp_pl_stdout(e2e, tmp_path):
    """One query profile vs a genome's genes in ONE GPU call (ranked targets), formatted: byte-identical to the stdout of
    the reference's bin/ppp.pl server run (modulo Perl's hash order inside score ties)."""
    from prophylo_b200 import capi

    ctx = capi.Context(0)
    for case in e2e:
        prof, paths = materialise(case, tmp_path)
        rows = drivers.ppp_scan(ctx, Profile(file=prof), paths, prob=case["prob"])
        got = drivers.format_ppp_table(rows, slope=case["slope"])
        assert normalise(got) == normalise(case["ppp_pl_stdout"]), case["seed"]
    ctx.close()


@pytest.mark.gpu
def test_ppp_scan_dup_on_gpu_reproduces_ppp_pl_dup_stdout(e2e_dup, tmp_path):
    """`ppp.pl --dup`: lists packed for the flag by the native reader (PPP_BLA_DUP) and by the Python packer, scored on the
    device with the "dup" option; the table equals the unmodified reference script's stdout."""
    from prophylo_b200 import capi

    ctx = capi.Context(0)
    try:
        for case in e2e_dup:
            prof, paths = materialise(case, tmp_path)
            for targets in (paths, {gi: drivers.read_bla(p) for gi, p in paths.items()}):
                rows = drivers.ppp_scan(ctx, Profile(file=prof), targets, prob=case["prob"], dup=True)
                got = drivers.format_ppp_table(rows, slope=case["slope"])
                assert normalise(got) == normalise(case["ppp_pl_stdout"]), case["seed"]
    finally:
        ctx.close()


@pytest.fixture(scope="module")
def ppp2d_cases(golden_dir):
    """stdout of the UNMODIFIED bin/ppp2d.pl (forking the unmodified bin/ppp.pl once per depth) on synthetic caches
    (oracle/gen_golden_ppp2d.py); rows with tied printed scores are stored sorted by source line"""
    return json.load(open(os.path.join(golden_dir, "ppp2d.json")))["cases"]


def materialise_2d(case, tmp_path):
    d = tmp_path / f"ppp2d{case['seed']}"
    (d / "bla").mkdir(parents=True)
    paths = {}
    for gi, txt in case["bla"].items():
        p = d / "bla" / f"{gi}.bla"
        p.write_text(txt)
        paths[gi] = str(p)
    return paths


def test_ppp2d_host_logic_matches_reference_with_oracle_scores(ppp2d_cases, tmp_path):
    """CPU: depth-prefix profiles (bin/ppp2d.pl:253-315), p = taxa / 1459, best non-self locus per depth, printed-score
    filter and ordering (:322-364), with every pair scored by the oracle: equals the reference script's stdout."""
    for case in ppp2d_cases:
        paths = materialise_2d(case, tmp_path)
        gi = case["gi"]
        lists = {g: drivers.read_bla(p) for g, p in paths.items()}
        qs = drivers.ppp2d_prefix_queries(gi, paths[gi], case["taxdis"], start=case["start"], end=case["end"], skip=case["skip"])
        best = []
        for _, _, profq, pq in qs:
            rows = []
            for g in sorted(lists):
                s, y, n, d, _ = oracle.score_case(profq, lists[g], pq)
                rows.append((g, s, y, n, d))
            rows.sort(key=lambda r: (float("%.15g" % r[1]), r[0]))
            best.append(rows[:2])
        assert drivers.format_ppp2d(gi, qs, best) == case["stdout"], case["seed"]


@pytest.mark.gpu
def test_ppp2d_on_gpu_reproduces_reference_stdout(ppp2d_cases, tmp_path):
    """bin/ppp2d.pl in process: every depth prefix of the gene against every gene of the genome in ONE GPU call (top-2 per
    depth); byte-identical to the stdout of the unmodified reference script."""
    from prophylo_b200 import capi

    ctx = capi.Context(0)
    try:
        for case in ppp2d_cases:
            paths = materialise_2d(case, tmp_path)
            got = drivers.ppp2d_scan(ctx, case["gi"], paths[case["gi"]], dict(sorted(paths.items())), case["taxdis"], start=case["start"],
                                     end=case["end"], skip=case["skip"])
            assert got == case["stdout"], case["seed"]
    finally:
        ctx.close()


@pytest.fixture(scope="module")
def cross_cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ppp_cross_score.json")))["cases"]


def test_cross_score_host_logic_matches_reference_with_oracle_scores(cross_cases):
    """CPU: the depth-prefix construction, the target list, the %.3f ranking and the stdout format of
    bin/ppp_cross_score.pl (BLAST x BLAST) with the per-pair scores taken from the oracle instead of the GPU."""

    def oracle_scorer(queries, target, probs):
        return [oracle.score_case(q, target, prob=p)[:4] for q, p in zip(queries, probs)]

    for case in cross_cases:
        n_taxa = sum(1 for ln in case["dist"].split("\n") if ln and not ln.startswith("#"))
        got = drivers.format_cross_score(oracle_scorer, "a.m8", case["file1"], "b.m8", case["file2"], case["gi2taxid"], n_taxa,
                                         start=case["start"], end=case["end"], skip=case["skip"], scale=case["scale"] or 1.0)
        assert got == case["stdout"], (case["seed"], got, case["stdout"])


@pytest.mark.gpu
def test_cross_score_on_gpu_reproduces_reference_stdout(cross_cases):
    """bin/ppp_cross_score.pl in process: every depth prefix of one BLAST table against the other's hit list, one GPU
    call per direction; byte-identical to the stdout of the unmodified reference script (oracle/gen_golden_cross.py)."""
    from prophylo_b200 import capi

    ctx = capi.Context(0)
    try:
        for case in cross_cases:
            n_taxa = sum(1 for ln in case["dist"].split("\n") if ln and not ln.startswith("#"))
            got = drivers.cross_score_blast(ctx, "a.m8", case["file1"], "b.m8", case["file2"], case["gi2taxid"], n_taxa,
                                            start=case["start"], end=case["end"], skip=case["skip"], scale=case["scale"] or 1.0)
            assert got == case["stdout"], (case["seed"], got, case["stdout"])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_perl_cli_dropin_reproduces_ppp_pl_stdout(e2e, e2e_dup, tmp_path):
    """perl/bin/ppp_b200.pl -- same options, database layout and stdout table as the reference's bin/ppp.pl, with the
    chunk files / forked Perl clients replaced by one GPU call through the XS binding and the native `.bla` reader --
    against the stdout of the unmodified reference script (modulo Perl's hash order inside score ties)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "perl")], stdout=subprocess.DEVNULL)
    for case in e2e + e2e_dup:
        d = tmp_path / f"cli{case['seed']}"
        (d / "db" / "blast" / "1140").mkdir(parents=True)
        (d / "ws").mkdir()
        (d / "q.profile").write_text(case["profile"])
        for gi, txt in case["bla"].items():
            (d / "db" / "blast" / "1140" / f"{gi}.bla").write_text(txt)
        cmd = ["perl", os.path.join(root, "perl", "bin", "ppp_b200.pl"), "-d", str(d / "db"), "-t", "1140", "-p", str(d / "q.profile"),
               "-w", str(d / "ws"), "--threads", "2"]
        if case["prob"]:
            cmd += ["--prob", str(case["prob"])]
        if case["slope"]:
            cmd += ["-m", str(case["slope"])]
        if case.get("dup"):
            cmd += ["--dup"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert normalise(r.stdout) == normalise(case["ppp_pl_stdout"]), case["seed"]


def _write_family(tmp_path, n=14, ntaxa=260, seed=21):
    rng = np.random.default_rng(seed)
    taxa = [str(x) for x in rng.choice(np.arange(100, 90000), ntaxa, replace=False)]
    paths, profs = [], []
    for i in range(n):
        listed = [t for t in taxa if rng.random() < 0.9]
        dens = float(np.exp(rng.uniform(np.log(0.03), np.log(0.4))))
        prof = {t: int(rng.random() < dens) for t in listed}
        if i and i % 5 == 0:  # a near copy of an earlier profile: strong matches
            prof = {t: (v if rng.random() > 0.05 else 1 - v) for t, v in profs[i - 1].items()}
        rows = sorted(prof.items(), key=lambda kv: float(kv[0]))
        p = tmp_path / f"fam{i}.profile"
        p.write_text("# taxon\tvalue\n" + "".join(f"{t}\t{v}\n" for t, v in rows))
        paths.append(str(p))
        profs.append(prof)
    return paths, profs


def test_profile_family_universe_and_planes(tmp_path):
    """Profile files -> shared universe in ascending numeric taxon order (data/taxdis.txt / hmm3profile.pl order), query
    planes and target planes; p = yes / total over the file's own taxa (ppp.pm:138-142)."""
    from prophylo_b200 import synth

    paths, profs = _write_family(tmp_path)
    universe, P, Y, p, T = drivers.read_profile_family(paths)
    assert universe == sorted({t for q in profs for t in q}, key=float)
    G = len(universe)
    Pb, Yb = synth.unpack_bits(P, G), synth.unpack_bits(Y, G)
    for i, q in enumerate(profs):
        assert {universe[g] for g in np.nonzero(Pb[i])[0]} == set(q)
        assert {universe[g] for g in np.nonzero(Yb[i])[0]} == {t for t, v in q.items() if v}
        assert p[i] == sum(q.values()) / len(q)
    assert np.array_equal(T, Y)


@pytest.mark.gpu
def test_profile_family_all_pairs_vs_oracle(tmp_path):
    """BASELINE config 5 through real profile files: every profile against every profile's yes set (canonical order),
    dense, top-k and ppp_cutoff threshold -- each pair against the oracle's get_match_score on the same lists."""
    from prophylo_b200 import capi

    paths, profs = _write_family(tmp_path)
    ctx = capi.Context(0)
    try:
        universe, hits = drivers.profile_family_scan(ctx, paths)
        n = len(paths)
        assert hits.size == n * n
        for h in hits:
            q, t = profs[int(h["q"])], profs[int(h["t"])]
            target = [x for x in universe if t.get(x)]
            s, y, num, d, _ = oracle.score_case({k: v for k, v in q.items()}, target)
            assert (int(h["yes"]), int(h["number"]), int(h["depth"])) == (y, num, d), (h, y, num, d)
            assert h["score"] == s or abs(np.log(h["score"]) - np.log(s)) <= 1e-9 * abs(np.log(s))
        dense = np.sort(hits, order=["q", "score", "t"])
        _, top = drivers.profile_family_scan(ctx, paths, k=3)
        exp = np.concatenate([dense[dense["q"] == q][:3] for q in range(n)])
        assert np.array_equal(top[["q", "t", "yes", "number", "depth"]], exp[["q", "t", "yes", "number", "depth"]])
        _, th = drivers.profile_family_scan(ctx, paths, thresh_log10=-0.0005)
        keep = dense[np.log10(dense["score"]) < -0.0005]
        assert np.array_equal(np.sort(th, order=["q", "score", "t"])[["q", "t", "depth"]], keep[["q", "t", "depth"]])
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------------ the Perl command-line drop-ins
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERL_BIN = os.path.join(ROOT_DIR, "perl", "bin")


def perl_env(tmp, gi2taxid=None, nodes=None):
    """PERL5LIB with the test stubs of the reference's host modules (SeqToolBox::Taxonomy answering from TSV tables,
    PhyloProf::Profile) -- on the GPU box the reference tree does not exist; in production its own lib/ is on PERL5LIB."""
    env = dict(os.environ, PERL5LIB=os.path.join(ROOT_DIR, "perl", "t", "lib"))
    if gi2taxid is not None:
        p = os.path.join(str(tmp), "gi2taxid.tsv")
        open(p, "w").write("".join(f"{g}\t{t}\n" for g, t in gi2taxid.items()))
        env["PPP_MOCK_GI2TAXID"] = p
    if nodes is not None:
        p = os.path.join(str(tmp), "nodes.tsv")
        open(p, "w").write(nodes)
        env["PPP_MOCK_NODES"] = p
    return env


def test_perl_cutoff_cli_matches_reference_script(e2e, tmp_path):
    """perl/bin/ppp_cutoff_b200.pl (pure text filter, no GPU): stdout and the CUTOFF= line identical to the unmodified
    bin/ppp_cutoff.pl for every golden table and slope, from a file and from stdin."""
    import subprocess

    n = 0
    for case in e2e:
        table = tmp_path / f"table{case['seed']}.txt"
        table.write_text(case["ppp_pl_stdout"])
        for slope, ref in case["cutoff"].items():
            for kw in (dict(args=[str(table)]), dict(args=[], input=case["ppp_pl_stdout"])):
                r = subprocess.run(["perl", os.path.join(PERL_BIN, "ppp_cutoff_b200.pl"), "-s", slope] + kw["args"], input=kw.get("input"),
                                   capture_output=True, text=True, timeout=60)
                assert r.returncode == 0, r.stderr
                assert r.stdout == ref["stdout"], (case["seed"], slope)
                assert [ln for ln in r.stderr.split("\n") if ln.startswith("CUTOFF=")] == [ln for ln in ref["stderr"].split("\n") if ln.startswith("CUTOFF=")]
                n += 1
    assert n == 24


@pytest.mark.gpu
def test_perl_ppp2d_cli_reproduces_reference_stdout(ppp2d_cases, tmp_path):
    """perl/bin/ppp2d_b200.pl: same options and stdout as the unmodified bin/ppp2d.pl (which forks bin/ppp.pl per depth); all
    depths x all genes in one GPU call through XS."""
    import subprocess

    subprocess.check_call(["make", "-C", os.path.join(ROOT_DIR, "perl")], stdout=subprocess.DEVNULL)
    for case in ppp2d_cases:
        d = tmp_path / f"cli2d{case['seed']}"
        (d / "db" / "blast" / case["genome"]).mkdir(parents=True)
        (d / "ws").mkdir()
        for gi, txt in case["bla"].items():
            (d / "db" / "blast" / case["genome"] / f"{gi}.bla").write_text(txt)
        (d / "taxdis.txt").write_text("".join(t + "\n" for t in case["taxdis"]))
        cmd = ["perl", os.path.join(PERL_BIN, "ppp2d_b200.pl"), "-d", str(d / "db"), "-w", str(d / "ws"), "-g", case["gi"], "--start-depth", str(case["start"]),
               "--skip", str(case["skip"]), "--threads", "2", "--taxdis", str(d / "taxdis.txt")]
        if case["end"]:
            cmd += ["--end-depth", str(case["end"])]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=perl_env(d, {case["gi"]: case["genome"]}))
        assert r.returncode == 0, r.stderr
        assert r.stdout == case["stdout"], case["seed"]


@pytest.fixture(scope="module")
def cross_hmmer_cases(golden_dir):
    """stdout of the UNMODIFIED bin/ppp_cross_score.pl with --method hmmer on either or both sides (oracle/gen_golden_hmmer.py)"""
    return json.load(open(os.path.join(golden_dir, "ppp_cross_score_hmmer.json")))["cases"]


@pytest.mark.gpu
def test_perl_cross_score_cli_reproduces_reference_stdout(cross_cases, cross_hmmer_cases, tmp_path):
    """perl/bin/ppp_cross_score_b200.pl, BLAST x BLAST, HMMER x HMMER and mixed (bin/ppp_cross_score.pl:415-546, 549-679;
    target profiles :326-413), with --trusted_cutoff / --start-depth / --end-depth / --skip / --scale: byte-identical to the
    unmodified reference script's stdout, every direction one GPU call."""
    import subprocess

    subprocess.check_call(["make", "-C", os.path.join(ROOT_DIR, "perl")], stdout=subprocess.DEVNULL)
    for case in cross_cases + cross_hmmer_cases:
        m1, m2 = case.get("method1", "blast"), case.get("method2", "blast")
        d = tmp_path / f"cross{case['seed']}"
        d.mkdir()
        n1, n2 = ("a.m8", "b.m8") if "method1" not in case else ("a.out", "b.out")
        (d / n1).write_text(case["file1"])
        (d / n2).write_text(case["file2"])
        (d / "dist.txt").write_text(case["dist"])
        cmd = ["perl", os.path.join(PERL_BIN, "ppp_cross_score_b200.pl"), "--file1", n1, "--file2", n2, "--method1", m1, "--method2", m2,
               "--dist-file", "dist.txt", "--start-depth", str(case["start"]), "--skip", str(case["skip"])]
        if case["end"]:
            cmd += ["--end-depth", str(case["end"])]
        if case["scale"]:
            cmd += ["--scale", str(case["scale"])]
        if case.get("cutoff"):
            cmd += ["--trusted_cutoff", str(case["cutoff"])]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(d), env=perl_env(d, case["gi2taxid"]))
        assert r.returncode == 0, r.stderr
        assert r.stdout == case["stdout"], (case["seed"], m1, m2, r.stdout, case["stdout"])


@pytest.mark.gpu
def test_perl_hmmer_cli_reproduces_reference_stdout(golden_dir, tmp_path):
    """perl/bin/ppp_hmmer_b200.pl: one query profile against one hmmsearch 3 report (bin/ppp_hmmer.pl:179-251; full-sequence
    and --domain scores, -t, --prob): byte-identical to the unmodified reference script's stdout."""
    import subprocess

    subprocess.check_call(["make", "-C", os.path.join(ROOT_DIR, "perl")], stdout=subprocess.DEVNULL)
    cases = json.load(open(os.path.join(golden_dir, "ppp_hmmer.json")))["cases"]
    for case in cases:
        d = tmp_path / f"hmm{case['seed']}"
        d.mkdir()
        (d / "q.profile").write_text(case["profile"])
        (d / "r.out").write_text(case["hmm"])
        cmd = ["perl", os.path.join(PERL_BIN, "ppp_hmmer_b200.pl"), "-p", "q.profile", "--hmm", "r.out"]
        if case["domain"]:
            cmd += ["--domain"]
        if case["cutoff"]:
            cmd += ["-t", str(case["cutoff"])]
        if case["prob"]:
            cmd += ["--prob", str(case["prob"])]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(d), env=perl_env(d, case["gi2taxid"]))
        assert r.returncode == 0, r.stderr
        assert r.stdout == case["stdout"], (case["seed"], r.stdout[:500], case["stdout"][:500])
