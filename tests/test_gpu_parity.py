"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI (ctypes), against
the oracle on the same seeded inputs and against the committed golden vectors produced by the reference itself.
Bar: yes / number / depth bit-exact; score within 1e-9 relative on -log(score) (bit-identical where the exact
tier ran).  /root/reference is never read here."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from prophylo_b200 import capi, synth

pytestmark = pytest.mark.gpu

REL_TOL_NEGLOG = 1e-9  # north_star: relative tolerance on -log p


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def assert_hits_match(hits, ref, what=""):
    """hits: HIT_DTYPE [nq*nt] dense; ref: oracle structured array [nq, nt]."""
    nq, nt = ref.shape
    h = hits.reshape(nq, nt)
    assert np.array_equal(h["q"], np.repeat(np.arange(nq, dtype=np.uint32)[:, None], nt, axis=1))
    assert np.array_equal(h["t"], np.repeat(np.arange(nt, dtype=np.uint32)[None, :], nq, axis=0))
    for f in ("yes", "number", "depth"):
        bad = np.nonzero(h[f] != ref[f])
        assert bad[0].size == 0, f"{what}: {f} differs at {list(zip(*bad))[:5]}: gpu {h[f][bad][:5]} ref {ref[f][bad][:5]}"
    gs, rs = h["score"], ref["score"]
    same = gs == rs
    with np.errstate(divide="ignore", invalid="ignore"):
        lg, lr = -np.log(gs), -np.log(rs)
        rel = np.abs(lg - lr) / np.abs(lr)
    ok = same | (rel <= REL_TOL_NEGLOG)
    bad = np.nonzero(~ok)
    assert bad[0].size == 0, f"{what}: score differs at {list(zip(*bad))[:5]}: gpu {gs[bad][:5]} ref {rs[bad][:5]} rel {rel[bad][:5]}"


def test_exact_tier_bdtrc_is_bit_identical(ctx, golden_dir):
    pts = json.load(open(os.path.join(golden_dir, "bdtrc_points.json")))["points"]
    k = np.array([x[0] for x in pts], dtype=np.int32)
    n = np.array([x[1] for x in pts], dtype=np.int32)
    p = np.array([float(x[2]) for x in pts])
    exp = np.array([float(x[3]) for x in pts])
    got = ctx.debug_bdtrc(k, n, p)
    libm = (k == 0) & (p < 0.01)  # host-libm expm1 branch of the reference: fdlibm restatement, <= 1 ulp
    assert np.array_equal(got[~libm].view(np.uint64), exp[~libm].view(np.uint64))
    assert np.all(np.abs(got[libm] - exp[libm]) <= np.spacing(exp[libm]))
    assert np.mean(got[libm] == exp[libm]) > 0.99


def test_exact_tier_bdtrc_random_vs_oracle(ctx):
    rng = np.random.default_rng(17)
    m = 200000
    n = rng.integers(1, 8193, size=m).astype(np.int32)
    p = np.where(rng.random(m) < 0.5, rng.random(m), 10 ** rng.uniform(-5, 0, size=m))
    k = np.minimum(n, np.maximum(0, (n * p + rng.normal(size=m) * (1 + 4 * np.sqrt(n * p * (1 - p)))).astype(np.int64))).astype(np.int32)
    k[::7] = rng.integers(1, 6, size=k[::7].size)
    k = np.minimum(k, n).astype(np.int32)
    got = ctx.debug_bdtrc(k, n, p)
    exp = oracle.bdtrc_batch(k, n, p)
    libm = (k == 0) & (p < 0.01)
    assert np.array_equal(got[~libm].view(np.uint64), exp[~libm].view(np.uint64))


@pytest.mark.parametrize("force_exact", [0, 1])
def test_golden_bit_cases(ctx, golden_dir, force_exact):
    g = json.load(open(os.path.join(golden_dir, "ppp_bit_cases.json")))
    ctx.set_option("force_exact", force_exact)
    try:
        for case in g["cases"]:
            sp = case["spec"]
            G = sp["G"]
            P, Y, p = synth.make_queries(G, sp["nq"], sp["qseed"], sp["kind"])
            T = synth.make_targets(G, sp["nt"], sp["tseed"], plant_from=Y, plant_frac=0.25)
            ctx.set_targets_bits(T, G)
            hits = ctx.score(P, Y, p)
            ref = np.zeros((sp["nq"], sp["nt"]), dtype=oracle.RESULT_DTYPE)
            for i, r in enumerate(case["results"]):
                ref.flat[i] = (float(r[0]), r[1], r[2], r[3], 0, 0.0)
            assert_hits_match(hits, ref, f"golden {sp} force_exact={force_exact}")
            if force_exact:
                # the exact tier is the reference's arithmetic: bit-identical scores; 1 ulp only where the winning step is
                # bdtrc's k == 0, p < 0.01 branch, which the reference evaluates with the host libm's expm1 (cephes_exact.cuh)
                gs, rs = hits["score"], ref["score"].reshape(-1)
                same = gs.view(np.uint64) == rs.view(np.uint64)
                libm = (hits["yes"] == 1) & (np.repeat(p, sp["nt"]) < 0.01)
                assert np.all(same | libm), f"golden {sp}: {int((~(same | libm)).sum())} scores differ from the reference bit pattern"
                assert np.all(np.abs(gs[~same] - rs[~same]) <= np.spacing(rs[~same]))
    finally:
        ctx.set_option("force_exact", 0)


@pytest.mark.parametrize("G,kind,nq,nt", [(1024, "single", 1, 3000), (1024, "mixed", 40, 300), (4096, "mixed", 70, 200),
                                          (2048, "prefix", 64, 150), (8192, "family", 33, 120), (1459, "mixed", 20, 100),
                                          (100, "mixed", 35, 64)])
def test_random_pairs_vs_oracle(ctx, G, kind, nq, nt):
    P, Y, p = synth.make_queries(G, nq, 1000 + G, kind)
    T = synth.make_targets(G, nt, 2000 + G, plant_from=Y, plant_frac=0.05)
    ctx.set_targets_bits(T, G)
    hits = ctx.score(P, Y, p)
    ref, _ = oracle.score_bits_batch(P, Y, p, T, G, nthreads=os.cpu_count() or 1)
    assert_hits_match(hits, ref, f"G={G} {kind}")
    st = ctx.stats()
    assert st["pairs"] == nq * nt and st["kernel_launches"] >= 2


def test_force_exact_is_bit_identical_to_oracle(ctx):
    G, nq, nt = 1024, 8, 80
    P, Y, p = synth.make_queries(G, nq, 5, "mixed")
    p[:] = np.maximum(p, 0.011)  # keep away from the host-libm expm1 branch (k == 0, p < 0.01)
    T = synth.make_targets(G, nt, 6, plant_from=Y, plant_frac=0.2)
    ctx.set_targets_bits(T, G)
    ctx.set_option("force_exact", 1)
    try:
        hits = ctx.score(P, Y, p)
    finally:
        ctx.set_option("force_exact", 0)
    ref, _ = oracle.score_bits_batch(P, Y, p, T, G, nthreads=os.cpu_count() or 1)
    h = hits.reshape(nq, nt)
    for f in ("yes", "number", "depth"):
        assert np.array_equal(h[f], ref[f])
    assert np.array_equal(h["score"].view(np.uint64), ref["score"].view(np.uint64))


@pytest.mark.parametrize("G,kind,nq,nt", [(1024, "mixed", 48, 96), (4096, "mixed", 48, 96), (2048, "prefix", 2000, 24),
                                          (8192, "family", 40, 64), (1024, "single", 1, 2000)])
def test_enclosures_contain_accurate_value(ctx, G, kind, nq, nt):
    """Every tier-A enclosure [lo, hi] must contain the FP64 series value (checked on the device for every candidate of
    every pair), on every query kind of the BASELINE configs -- the dense prefix queries of config 4 (p up to 0.98) and
    G = 8192 included."""
    P, Y, p = synth.make_queries(G, nq, 77, kind)
    T = synth.make_targets(G, nt, 78, plant_from=Y, plant_frac=0.1)
    ctx.set_targets_bits(T, G)
    ctx.set_option("check_bounds", 1)
    ctx.set_option("table_mode", 0)
    try:
        ctx.score(P, Y, p)
    finally:
        ctx.set_option("check_bounds", 0)
        ctx.set_option("table_mode", -1)
    assert ctx.counter("bound_checks") > 10000
    assert ctx.counter("bound_violations") == 0


def _enclosure_points(rng, m, p_choices):
    n = rng.integers(1, 8193, size=m)
    p = rng.choice(p_choices, size=m)
    mean = n * p
    sd = np.sqrt(n * p * (1 - p)) + 1
    y = np.where(rng.random(m) < 0.5, mean + rng.normal(size=m) * 3 * sd, mean + rng.exponential(size=m) * 12 * sd)
    y = np.clip(np.rint(y), 1, n).astype(np.int32)
    return y, n.astype(np.int32), p.astype(np.float64)


def test_enclosures_against_exact_tier_and_oracle(ctx):
    """The staircase prune is sound only if tier A's lower bound never exceeds the TRUE ln P(X >= yes | number, p): a `lo`
    above the truth would silently drop a true top-k hit.  400,000 points (yes, number, p) -- p from 1e-6 to 0.98, yes from
    far below the mean to deep in the upper tail -- checked against the reference's own arithmetic: the device's exact
    tier (bit-identical to Math::Cephes bdtrc) and the oracle.  Where Cephes' value is subnormal or exact 0 (its logarithm
    is then quantised or gone) `lo` is only required to lie below the smallest normal double's logarithm; where it
    saturates at 1 only `hi` is checked."""
    rng = np.random.default_rng(2026)
    ps = [1e-6, 1e-4, 0.003, 0.02, 0.1, 0.25, 0.5, 0.75, 0.9, 0.98]
    y, n, p = _enclosure_points(rng, 400_000, ps)
    lo, hi, lnf = ctx.debug_enclosure(y, n, p)
    exact = ctx.debug_bdtrc(y - 1, n, p)  # P(X >= y) = bdtrc(y - 1, n, p)   (ppp.pm:225)
    ora = oracle.bdtrc_batch(y - 1, n, p)
    libm = (y == 1) & (p < 0.01)
    assert np.array_equal(exact[~libm].view(np.uint64), ora[~libm].view(np.uint64))
    pos = ora >= np.finfo(np.float64).tiny  # normal range
    with np.errstate(divide="ignore"):
        truth = np.where(pos, np.log(np.where(pos, ora, 1.0)), -np.inf)
    tol = 1e-12 * np.maximum(1.0, np.abs(np.where(pos, truth, 0.0)))  # Cephes' own relative error on the tail
    bad_lo = pos & (lo.astype(np.float64) > truth + tol)
    assert not bad_lo.any(), (y[bad_lo][:5], n[bad_lo][:5], p[bad_lo][:5], lo[bad_lo][:5], truth[bad_lo][:5])
    assert np.all(lo[~pos].astype(np.float64) <= np.log(np.finfo(np.float64).tiny))  # below the normal range: so is lo
    unsat = pos & (ora < 1.0)
    bad_hi = unsat & (hi.astype(np.float64) < truth - tol)
    assert not bad_hi.any(), (y[bad_hi][:5], n[bad_hi][:5], p[bad_hi][:5], hi[bad_hi][:5], truth[bad_hi][:5])
    # the FP64 series tier: within 1e-9 relative on -ln p wherever it is used (ln p in [-700, -1e-6])
    use = pos & (truth > -700.0) & (truth < -1e-6)
    rel = np.abs(lnf[use] - truth[use]) / np.abs(truth[use])
    assert rel.max() <= REL_TOL_NEGLOG, rel.max()
    assert use.sum() > 200_000


def test_rank_cases_on_gpu(ctx, golden_dir):
    """-rank (ppp.pm:328-370): the host classes collapse the target list through a mock taxonomy
    object exposing collapse_taxon, then score on the GPU; expected values are the unmodified reference's (oracle/
    gen_golden_rank.py).  ARRAY and HASH targets, -prob, -dup."""
    from prophylo_b200 import algorithm

    class MockTaxonomy:
        def __init__(self, table):
            self.table = table

        def collapse_taxon(self, taxon, rank):
            return self.table.get(rank, {}).get(taxon) or None

    g = json.load(open(os.path.join(golden_dir, "ppp_rank_cases.json")))
    algorithm._ctx = ctx
    try:
        for c in g["cases"]:
            t = c["target"]
            prof2 = algorithm.Profile(raw=t, yes=len(t), no=0)
            a = algorithm.ppp(algorithm.Profile(profile=c["query"]), prof2, prob=c.get("prob"), rank=c["rank"],
                              taxonomy=MockTaxonomy(c["collapse"]), dup=c.get("dup"))
            score, yes, number, depth, p = a.get_match_score()
            e = c["expect"]
            assert (yes, number, depth, p) == (e[1], e[2], e[3], float(e[4])), c["name"]
            es = float(e[0])
            assert score == es or abs(np.log(score) - np.log(es)) <= REL_TOL_NEGLOG * abs(np.log(es)), c["name"]
    finally:
        algorithm._ctx = None


def test_edge_cases(ctx):
    G = 256
    wpr = synth.words_per_row(G)
    rng = np.random.default_rng(3)
    Tb = (rng.random((12, G)) < 0.5).astype(np.uint8)
    Tb[0] = 0            # empty target
    Tb[1] = 1            # full target
    Pb = np.ones((7, G), dtype=np.uint8)
    Yb = (rng.random((7, G)) < 0.2).astype(np.uint8)
    Yb[1] = 0            # zero-yes query -> p = 0
    Yb[2] = 1            # all-yes query  -> p = 1
    Pb[3] = Yb[3]        # present == yes (cross-score degenerate form)
    p = Yb.sum(1) / Pb.sum(1)
    p[3] = Yb[3].sum() / G
    p[4] = 1.2           # domain error (KA-5)
    p[5] = -0.5
    p[6] = 1e-9
    Tb[2] = Yb[0]        # target == query yes set: extreme significance
    P, Y, T = synth.pack_bits(Pb), synth.pack_bits(Yb), synth.pack_bits(Tb)
    assert P.shape[1] == wpr
    ctx.set_targets_bits(T, G)
    hits = ctx.score(P, Y, p)
    ref, _ = oracle.score_bits_batch(P, Y, p, T, G)
    assert_hits_match(hits, ref, "edge cases")
    h = hits.reshape(7, 12)
    assert tuple(h[0, 0][["score", "yes", "number", "depth"]]) == (1.0, 0, 0, 0)
    assert tuple(h[4, 1][["score", "depth"]]) == (0.0, 1)


def test_underflow_to_zero_and_self_pairs(ctx):
    """Config 5's diagonal: a family scored against itself underflows Cephes to exactly 0.0 and freezes there."""
    G = 8192
    P, Y, p = synth.make_queries(G, 6, 9, "family")
    ctx.set_targets_bits(Y, G)  # targets = the queries' own yes planes
    hits = ctx.score(P, Y, p)
    ref, _ = oracle.score_bits_batch(P, Y, p, Y, G, nthreads=os.cpu_count() or 1)
    assert_hits_match(hits, ref, "self pairs")


def test_api_errors(ctx):
    c2 = capi.Context(0)
    with pytest.raises(capi.PPPError):
        c2.run()
    with pytest.raises(capi.PPPError):
        c2.set_targets_bits(np.zeros((1, synth.words_per_row(9000)), dtype=np.uint32), 9000)
    with pytest.raises(capi.PPPError):
        c2.set_option("nope", 1)
    c2.close()


def test_filters_match_dense(ctx):
    G, nq, nt = 1024, 37, 700
    P, Y, p = synth.make_queries(G, nq, 41, "mixed")
    T = synth.make_targets(G, nt, 42, plant_from=Y, plant_frac=0.03)
    ctx.set_targets_bits(T, G)
    dense = ctx.score(P, Y, p).reshape(nq, nt)
    th = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-2.0))
    with np.errstate(divide="ignore"):
        keep = np.log10(dense["score"]) < -2.0
    assert th.size == keep.sum()
    exp = np.sort(dense[keep], order=["q", "score", "t"])
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(th[f], exp[f]), f
    k = 10
    tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    assert tk.size == nq * k
    for q in range(nq):
        exp = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], exp[f]), (q, f)


@pytest.mark.parametrize("G,kind,nq,nt,k", [(4096, "mixed", 70, 9000, 100), (1024, "single", 1, 60000, 25), (2048, "prefix", 40, 5000, 1),
                                            (8192, "family", 33, 3000, 10)])
def test_topk_multichunk_matches_dense(ctx, G, kind, nq, nt, k):
    """Several target chunks -> staircase rebuilds, thread-per-pair pre-filter, refine, device merge: the lists must be
    exactly the first k rows of the dense result sorted by (score, target)."""
    P, Y, p = synth.make_queries(G, nq, 300 + G, kind)
    T = synth.make_targets(G, nt, 400 + G, plant_from=Y, plant_frac=0.02)
    ctx.set_targets_bits(T, G)
    ctx.set_option("table_mode", 0)  # compare like with like: the general sweep on both sides (bit-identical scores)
    try:
        dense = ctx.score(P, Y, p).reshape(nq, nt)
    finally:
        ctx.set_option("table_mode", -1)
    tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    assert ctx.counter("survivors") > 0 and ctx.counter("full_pairs") < nq * nt
    assert tk.size == nq * k
    for q in range(nq):
        exp = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("q", "t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], exp[f]), (q, f)


def test_threshold_filter_uses_staircase_and_matches_dense(ctx):
    G, nq, nt = 4096, 48, 4000
    P, Y, p = synth.make_queries(G, nq, 51, "mixed")
    T = synth.make_targets(G, nt, 52, plant_from=Y, plant_frac=0.02)
    ctx.set_targets_bits(T, G)
    dense = ctx.score(P, Y, p).reshape(nq, nt)
    for thresh in (-3.0, -0.0005):
        th = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=thresh))
        with np.errstate(divide="ignore"):
            keep = np.log10(dense["score"]) < thresh
        exp = np.sort(dense[keep], order=["q", "score", "t"])
        assert th.size == exp.size
        for f in ("q", "t", "score", "yes", "number", "depth"):
            assert np.array_equal(th[f], exp[f]), (thresh, f)


def test_threshold_results_ranked_on_device(ctx):
    """Threshold runs rank their hits by (q, score, t) on the device (ppp_rank.cu: the order bin/ppp.pl:473 prints in).
    All-pairs family shape with the diagonal (exact 0.0 scores -> ties broken by target index), enough hits for
    several sort tiles and all 64 score bits: identical, record for record, to the host-ordered result."""
    G, n = 8192, 700
    P, Y, p = synth.make_queries(G, n, 77, "family")
    ctx.set_targets_bits(Y, G)
    flt = capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-0.0005)
    dev = ctx.score(P, Y, p, flt).copy()
    assert ctx.counter("us_merge") > 0
    ctx.set_option("device_rank", 0)
    try:
        host = ctx.score(P, Y, p, flt).copy()
    finally:
        ctx.set_option("device_rank", 1)
    assert dev.size == host.size and dev.size > 3 * 4096
    assert dev.tobytes() == host.tobytes()
    # the same run cut into several launches: each launch's hits are gathered on the device and ranked once at the end
    ctx.set_option("launch_pairs_log2", 17)
    try:
        multi = ctx.score(P, Y, p, flt).copy()
        assert ctx.stats()["sweep_launches"] > 2
        assert ctx.printed_cutoffs(20.0)["row_off"][-1] == multi.size  # still resident and ranked on the device
    finally:
        ctx.set_option("launch_pairs_log2", 0)
    assert multi.tobytes() == host.tobytes()
    key = np.lexsort((dev["t"], dev["score"], dev["q"]))
    assert np.array_equal(key, np.arange(dev.size))
    assert (dev["score"] == 0.0).sum() >= 10  # diagonal pairs of the larger families underflow


def test_printed_scores_and_slope_cutoffs_on_device(ctx):
    """ppp_printed_cutoffs: the per-query tail of the drivers (sort, "%.3f" log10 score, ppp.pl -m marker, ppp_cutoff.pl
    CUTOFF; bin/ppp.pl:473-506, bin/ppp_cutoff.pl:130-148) on the device, against the host restatements in
    prophylo_b200.drivers -- which reproduce the unmodified reference scripts byte for byte (tests/test_drivers.py)."""
    from prophylo_b200 import drivers

    G, nq, nt = 1024, 60, 3000
    P, Y, p = synth.make_queries(G, nq, 91, "mixed")
    T = synth.make_targets(G, nt, 92, plant_from=Y, plant_frac=0.05)
    ctx.set_targets_bits(T, G)
    hits = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-0.0005)).copy()
    for slope in (10.0, 30.0):
        r = ctx.printed_cutoffs(slope)
        off = r["row_off"].astype(np.int64)
        assert off[0] == 0 and off[-1] == hits.size and np.all(np.diff(off) >= 0)
        checked = 0
        for q in range(nq):
            rows = hits[off[q]:off[q + 1]]
            assert np.all(rows["q"] == q)
            if r["status"][q] == capi.CUT_LOG0:
                assert (rows["score"] == 0).any()
                continue
            assert not (rows["score"] == 0).any()
            if r["status"][q] == capi.CUT_DOUBT or rows.size == 0:
                continue
            # printed scores: the server reads the client's %.15g string back before taking the logarithm (bin/ppp.pl:582)
            strs = [drivers._log10_3f(drivers._perl_num(float(s))) for s in rows["score"]]
            assert [int(round(float(x) * 1000)) for x in strs] == r["milli"][off[q]:off[q + 1]].tolist(), q
            table_rows = [(int(h["t"]), float(h["score"]), int(h["yes"]), int(h["number"]), int(h["depth"]), float(p[q])) for h in rows]
            if r["status"][q] == capi.CUT_DIV0:
                with pytest.raises(ZeroDivisionError):
                    drivers.format_ppp_table(table_rows, slope=slope)
                text = drivers.format_ppp_table(table_rows)
            else:
                text = drivers.format_ppp_table(table_rows, slope=slope)
                lines = text.split("\n")[1:]
                exp_marker = next((i for i, ln in enumerate(lines) if ln and set(ln) == {"-"}), -1)
                assert r["marker"][q] == exp_marker, (q, slope)
            _, cutoff = drivers.ppp_cutoff(text, slope)
            got = r["cutoff_milli"][q]
            assert (cutoff is None and got == capi.CUT_NONE) or (cutoff is not None and got == int(round(cutoff * 1000))), (q, slope, cutoff, got)
            checked += 1
        assert checked > nq // 2
    assert (r["marker"] >= 0).sum() > 5 and (r["cutoff_milli"] != capi.CUT_NONE).sum() > 5
    ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=5))
    with pytest.raises(capi.PPPError):
        ctx.printed_cutoffs(10.0)


def test_full_size_properties(ctx):
    """Size-independent properties at BASELINE genome counts: every top-k list is sorted, within range, idempotent
    under a second run, a sub-list of the k+5 list, and each reported hit re-scores to the same record alone."""
    G, nq, nt, k = 4096, 256, 20000, 20
    P, Y, p = synth.make_queries(G, nq, 61, "mixed")
    T = synth.make_targets(G, nt, 62, plant_from=Y, plant_frac=0.01)
    ctx.set_targets_bits(T, G)
    a = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k)).copy()
    b = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k)).copy()
    assert np.array_equal(a, b)
    c = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k + 5)).copy()
    for q in range(0, nq, 7):
        la, lc = a[a["q"] == q], c[c["q"] == q]
        assert la.size == k and np.all(np.diff(la["score"]) >= 0) and la["t"].max() < nt
        assert np.array_equal(la[["t", "score", "depth"]], lc[:k][["t", "score", "depth"]])
    # re-score a sample of reported hits as single dense pairs
    pick = a[:: max(1, a.size // 40)]
    for h in pick:
        ctx.set_targets_bits(T[h["t"]:h["t"] + 1], G)
        one = ctx.score(P[h["q"]:h["q"] + 1], Y[h["q"]:h["q"] + 1], p[h["q"]:h["q"] + 1])[0]
        assert (one["score"], one["yes"], one["number"], one["depth"]) == (h["score"], h["yes"], h["number"], h["depth"])


def test_perl_dropin_on_golden_cases():
    """The reference-facing binding: perl XS -> libppp_b200.so, PhyloProf::Algorithm::ppp_b200 (same constructor and
    return list as ppp.pm) on the golden list cases generated by the reference."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "perl")], stdout=subprocess.DEVNULL)
    r = subprocess.run(["perl", os.path.join(root, "perl", "t", "golden.pl")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


def test_device_global_topk_merge_of_two_shards(ctx):
    """The multi-GPU exchange step on one device: per-shard top-k lists, merged by ppp_merge_topk_device with shard
    offsets, must equal the top-k of the unsharded run."""
    import torch

    from prophylo_b200 import parallel

    G, nq, nt, k = 1024, 50, 3000, 16
    P, Y, p = synth.make_queries(G, nq, 71, "mixed")
    T = synth.make_targets(G, nt, 72, plant_from=Y, plant_frac=0.02)
    flt = capi.make_filter(capi.FILTER_TOPK, k=k)
    ctx.set_targets_bits(T, G)
    full = ctx.score(P, Y, p, flt).copy()
    parts, cnts, offs = [], [], []
    for r in range(2):
        t0, t1 = parallel.shard_range(nt, r, 2)
        ctx.set_targets_bits(T[t0:t1], G)
        h = ctx.score(P, Y, p, flt)
        lp, cp, kk, nqq = ctx.topk_device()
        assert (kk, nqq) == (k, nq)
        dev = torch.as_tensor(parallel._DevBuf(lp, nq * k * 32), device="cuda").clone()
        cnt = torch.as_tensor(parallel._DevBuf(cp, nq * 4), device="cuda").clone()
        assert np.array_equal(dev.cpu().numpy().view(capi.HIT_DTYPE), h)
        parts.append(dev)
        cnts.append(cnt)
        offs.append(t0)
    g_lists, g_counts = torch.cat(parts), torch.cat(cnts)
    torch.cuda.synchronize()
    out, cnt = ctx.merge_topk_device(g_lists.data_ptr(), g_counts.data_ptr(), 2, offs, nq, k)
    assert np.all(cnt == k)
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(out[f], full[f]), f


def test_python_operator_mirror_on_golden_list_cases(golden_dir):
    """prophylo_b200.algorithm mirrors PhyloProf::Profile / PhyloProf::Algorithm::ppp; every golden list case from the
    reference (ARRAY targets with duplicates and foreign taxa, HASH targets, -prob, -dup) through the ranked-target path."""
    from prophylo_b200.algorithm import Profile, ppp

    g = json.load(open(os.path.join(golden_dir, "ppp_list_cases.json")))
    n = 0
    for c in g["known_answers"] + g["fuzz"]:
        t = c["target"]
        prof1 = Profile(profile=c["query"])
        prof2 = Profile(raw=t, yes=len(t), no=0)
        got = ppp(prof1, prof2, prob=c.get("prob"), dup=c.get("dup")).get_match_score()
        e = c["expect"]
        assert got[1:4] == (e[1], e[2], e[3]), (c["name"], got, e)
        assert got[4] == float(e[4])
        es = float(e[0])
        assert got[0] == es or abs(np.log(got[0]) - np.log(es)) <= REL_TOL_NEGLOG * abs(np.log(es)), (c["name"], got, e)
        n += 1
    assert n > 600


def test_ranked_targets_batch_vs_oracle(ctx):
    """Many ranked targets (own order, duplicates, foreign taxa) x many queries in one call, dense / top-k / threshold."""
    from prophylo_b200 import packer

    rng = np.random.default_rng(5)
    G, nq, nt = 600, 37, 300
    universe = [f"g{i}" for i in range(G)]
    index = packer.universe_index(universe)
    queries = []
    for _ in range(nq):
        pres = rng.random(G) < rng.choice([1.0, 0.6])
        dens = float(np.exp(rng.uniform(np.log(0.01), np.log(0.3))))
        queries.append({universe[g]: int(rng.random() < dens) for g in np.nonzero(pres)[0]})
    lists = []
    for t in range(nt):
        L = int(rng.integers(0, 1500))
        pool = rng.permutation(G)[: max(1, int(G * rng.uniform(0.05, 0.9)))]
        lst = [universe[int(pool[int(rng.integers(len(pool)))])] if rng.random() > 0.03 else f"foreign{int(rng.integers(40))}"
               for _ in range(L)]
        if t % 17 == 0 and queries[t % nq]:  # planted: enriched in some query's yes taxa
            ys = [k for k, v in queries[t % nq].items() if v]
            lst = [ys[int(rng.integers(len(ys)))] for _ in range(min(L, 3 * len(ys) + 1))] + lst
        lists.append(lst)
    P, Y, p = packer.pack_queries(queries, index)
    idx, raw, off = packer.pack_ranked_targets(lists, index)
    ctx.set_targets_ranked(idx, raw, off, G)
    dense = ctx.score(P, Y, p).reshape(nq, nt)
    # oracle: list form with integer ids; foreign taxa get ids >= G
    fid = {}
    for qi, q in enumerate(queries):
        present = np.zeros(G, dtype=np.uint8)
        isyes = np.zeros(G, dtype=np.uint8)
        for k, v in q.items():
            present[index[k]] = 1
            isyes[index[k]] = int(v != 0)
        for ti, lst in enumerate(lists):
            ids = [index[x] if x in index else G + fid.setdefault(x, len(fid)) for x in lst]
            exp = oracle.score_list(present, isyes, int(isyes.sum()), ids, float(p[qi]))
            h = dense[qi, ti]
            assert (h["yes"], h["number"], h["depth"]) == exp[1:4], (qi, ti, h, exp)
            assert h["score"] == exp[0] or abs(np.log(h["score"]) - np.log(exp[0])) <= REL_TOL_NEGLOG * abs(np.log(exp[0]))
    k = 7
    tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    for q in range(nq):
        e = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], e[f]), (q, f)
    th = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-2.5))
    with np.errstate(divide="ignore"):
        keep = np.log10(dense["score"]) < -2.5
    e = np.sort(dense[keep], order=["q", "score", "t"])
    assert th.size == e.size and np.array_equal(th["t"], e["t"]) and np.array_equal(th["depth"], e["depth"])


def test_ranked_targets_dup_flag_batch_vs_oracle(ctx):
    """-dup (ppp.pm:189-200: first repeated entry of a list re-counted, `yes` frozen from the second, `last` once yes
    exceeds |Y|, ppp.pm:222-224) on the device: lists packed with dup=True, sweep run with the "dup" option; every pair
    against the oracle's own -dup loop; lists that show every yes taxon and then repeat one (yes > max_yes) included."""
    from prophylo_b200 import packer

    rng = np.random.default_rng(11)
    G, nq, nt = 300, 33, 200
    universe = [f"g{i}" for i in range(G)]
    index = packer.universe_index(universe)
    queries = []
    for _ in range(nq):
        dens = float(np.exp(rng.uniform(np.log(0.01), np.log(0.2))))
        q = {u: int(rng.random() < dens) for u in universe if rng.random() < 0.8}
        queries.append(q)
    lists = []
    for t in range(nt):
        L = int(rng.integers(1, 400))
        lst = [universe[int(g)] for g in rng.permutation(G)[:L]]
        if t % 3 == 0:  # plant the query's yes taxa first, then repeat one of them: yes would exceed max_yes
            ys = [k for k, v in queries[t % nq].items() if v]
            if ys:
                rest = [x for x in lst if x not in set(ys)]
                lst = ys + [ys[int(rng.integers(len(ys)))]] + rest
        for _ in range(int(rng.integers(0, 4))):  # 0..3 repeated entries at random places
            lst.insert(int(rng.integers(1, len(lst) + 1)), lst[int(rng.integers(len(lst)))])
        lists.append(lst)
    P, Y, p = packer.pack_queries(queries, index)
    idx, raw, off = packer.pack_ranked_targets(lists, index, dup=True)
    ctx.set_targets_ranked(idx, raw, off, G)
    ctx.set_option("dup", 1)
    try:
        dense = ctx.score(P, Y, p).reshape(nq, nt)
        k = 5
        tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    finally:
        ctx.set_option("dup", 0)
    differs = 0
    for qi, q in enumerate(queries):
        present = np.zeros(G, dtype=np.uint8)
        isyes = np.zeros(G, dtype=np.uint8)
        for kx, v in q.items():
            present[index[kx]] = 1
            isyes[index[kx]] = int(v != 0)
        for ti, lst in enumerate(lists):
            ids = [index[x] for x in lst]
            exp = oracle.score_list(present, isyes, int(isyes.sum()), ids, float(p[qi]), True)
            differs += exp[:4] != oracle.score_list(present, isyes, int(isyes.sum()), ids, float(p[qi]), False)[:4]
            h = dense[qi, ti]
            assert (h["yes"], h["number"], h["depth"]) == exp[1:4], (qi, ti, h, exp)
            assert h["score"] == exp[0] or abs(np.log(h["score"]) - np.log(exp[0])) <= REL_TOL_NEGLOG * abs(np.log(exp[0]))
    assert differs > 20  # the flag changes results on this input
    for q in range(nq):
        e = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], e[f]), (q, f)


def test_topk_ties_at_zero_and_one(ctx):
    """Top-k lists whose k-th score is exactly 0.0 (Cephes underflow / domain error) or exactly 1.0 (no common
    yes-genome): ties are broken by target index and nothing may be lost to the strict `score < kth` acceptance."""
    G, nt, k = 2048, 2500, 40
    rng = np.random.default_rng(8)
    Pb = np.ones((6, G), dtype=np.uint8)
    Yb = np.zeros((6, G), dtype=np.uint8)
    Yb[0, :900] = 1                      # dense yes set: hundreds of targets underflow to exactly 0.0
    Yb[1, rng.choice(G, 5, replace=False)] = 1   # 5 yes genomes: most targets score exactly 1.0 or near it
    Yb[2] = rng.random(G) < 0.1          # domain error p = 1.2 -> every non-empty target scores 0.0 at depth 1
    Yb[3] = 1                            # p = 1 -> every target scores 1.0
    Yb[4] = rng.random(G) < 0.05
    Yb[5] = 0                            # p = 0
    p = Yb.sum(1) / G
    p[2] = 1.2
    p[0] = 0.01                          # -prob: 800 common yes genomes at p = 0.01 underflow Cephes to exactly 0.0
    Tb = (rng.random((nt, G)) < rng.uniform(0.0, 0.5, size=(nt, 1))).astype(np.uint8)
    Tb[::3, :900] |= (rng.random((len(Tb[::3]), 900)) < 0.9).astype(np.uint8)   # many targets rich in query 0's yes set
    Tb[5] = 0
    Tb[:, np.nonzero(Yb[1])[0]] &= (rng.random((nt, 5)) < 0.02).astype(np.uint8)  # query 1's yes genomes are rare
    P, Y, T = synth.pack_bits(Pb), synth.pack_bits(Yb), synth.pack_bits(Tb)
    ctx.set_targets_bits(T, G)
    dense = ctx.score(P, Y, p).reshape(6, nt)
    nz, no = int((dense[0]["score"] == 0).sum()), int((dense[1]["score"] == 1).sum())
    assert nz > k and no > nt // 2, (nz, no)
    tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    assert tk.size == 6 * k
    for q in range(6):
        exp = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], exp[f]), (q, f)


@pytest.mark.parametrize("G,kind,nq,nt", [(1024, "single", 1, 20000), (1024, "mixed", 24, 9000), (4096, "mixed", 6, 8500), (200, "family", 8, 300)])
def test_exact_table_mode_is_bit_identical(ctx, G, kind, nq, nt):
    """Dense runs of few queries x many targets use one bit-exact (yes, number) table per query: counts, cut-offs AND
    scores must equal the oracle's exactly (1 ulp only on the host-libm expm1 branch: yes == 1 and p < 0.01)."""
    P, Y, p = synth.make_queries(G, nq, 900 + G, kind)
    T = synth.make_targets(G, nt, 901 + G, plant_from=Y, plant_frac=0.05)
    ctx.set_targets_bits(T, G)
    ctx.set_option("table_mode", 1)
    try:
        hits = ctx.score(P, Y, p).reshape(nq, nt)
        assert ctx.counter("used_table") == 1
    finally:
        ctx.set_option("table_mode", -1)
    sub = slice(0, min(nt, 1500))
    ref, _ = oracle.score_bits_batch(P, Y, p, T[sub], G, nthreads=os.cpu_count() or 1)
    h = hits[:, sub]
    for f in ("yes", "number", "depth"):
        assert np.array_equal(h[f], ref[f]), f
    libm = (ref["yes"] == 1) & (p[:, None] < 0.01)
    assert np.array_equal(h["score"][~libm].view(np.uint64), ref["score"][~libm].view(np.uint64))
    assert np.all(np.abs(h["score"][libm] - ref["score"][libm]) <= np.spacing(ref["score"][libm]))
    # and the general sweep agrees with the table on every pair (cut-offs exactly, scores within the stated tolerance)
    ctx.set_option("table_mode", 0)
    try:
        gen = ctx.score(P, Y, p).reshape(nq, nt)
        assert ctx.counter("used_table") == 0
    finally:
        ctx.set_option("table_mode", -1)
    for f in ("yes", "number", "depth"):
        assert np.array_equal(gen[f], hits[f]), f
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(np.log(gen["score"]) - np.log(hits["score"])) / np.abs(np.log(hits["score"]))
    assert np.all((gen["score"] == hits["score"]) | (rel <= REL_TOL_NEGLOG))


@pytest.mark.parametrize("front_loaded", [False, True])
def test_topk_speculative_threshold_is_verified(ctx, front_loaded):
    """Top-k runs screen the bulk of the targets against a speculative threshold (rank r of the list after the first
    chunks) and verify the lists afterwards; queries that fail are re-scanned with the guaranteed bound.  Exchangeable
    targets: (almost) no failures.  Front-loaded targets (every strong match among the first targets, so the
    speculative threshold is far too tight): the re-scan must run and the lists must still be exactly the dense
    result's first k rows."""
    G, nq, nt, k = 2048, 24, 30000, 20
    P, Y, p = synth.make_queries(G, nq, 77, "mixed")
    T = synth.make_targets(G, nt, 78, plant_from=Y, plant_frac=0.0)
    rng = np.random.default_rng(79)
    Yb = synth.unpack_bits(Y, G)
    rows = np.arange(0, 600) if front_loaded else rng.choice(nt, 600, replace=False)
    planted = (Yb[rng.integers(0, nq, size=rows.size)] != 0) ^ (rng.random((rows.size, G)) < 0.03)
    T[rows] = synth.pack_bits(planted)
    ctx.set_targets_bits(T, G)
    ctx.set_option("table_mode", 0)
    try:
        dense = ctx.score(P, Y, p).reshape(nq, nt)
    finally:
        ctx.set_option("table_mode", -1)
    tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    assert ctx.counter("spec_rank") >= 1, "the speculative phase did not run"
    if front_loaded:
        assert ctx.counter("spec_failed") > 0, "front-loaded targets must make the verification fail for some query"
    assert tk.size == nq * k
    for q in range(nq):
        exp = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("q", "t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], exp[f]), (front_loaded, q, f)
    ctx.set_option("speculate", 0)
    try:
        tk0 = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    finally:
        ctx.set_option("speculate", 1)
    assert ctx.counter("spec_rank") == 0
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(tk0[f], tk[f]), f
    # the first chunk is speculative too (thresholds = the seeding pass's rs-th smallest bounds): same lists without it;
    # and with one / many speculation levels instead of the default schedule
    ctx.set_option("respec_min_pairs_log2", 0)  # further speculation levels are reserved for large runs by default
    for opt, val, restore in (("spec_first", 0, 1), ("respec_mult", 0, 8), ("respec_mult", 2, 8), ("respec_mult", 3, 8)):
        ctx.set_option(opt, val)
        try:
            tk1 = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
        finally:
            ctx.set_option(opt, restore)
            ctx.set_option("respec_min_pairs_log2", 0 if opt == "spec_first" or val != 3 else 27)
        assert ctx.counter("spec_rank") >= 1
        for f in ("q", "t", "score", "yes", "number", "depth"):
            assert np.array_equal(tk1[f], tk[f]), (opt, val, f)


@pytest.mark.parametrize("warps", [16, 8])
def test_g8192_sweep_warps(ctx, warps):
    """G = 8192 with 16 warps per CTA (the library's default): the query tile + log-factorial table leave shared memory for 15
    warps only -- the launchers must clamp instead of failing (regression), and top-k must still equal the dense result; the
    same with 8 warps (option sweep_warps)."""
    G, nq, nt, k = 8192, 12, 2600, 7
    P, Y, p = synth.make_queries(G, nq, 91, "family")
    T = synth.make_targets(G, nt, 92, plant_from=Y, plant_frac=0.02)
    ctx.set_option("sweep_warps", warps)
    ctx.set_option("table_mode", 0)
    try:
        ctx.set_targets_bits(T, G)
        dense = ctx.score(P, Y, p).reshape(nq, nt)
        tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    finally:
        ctx.set_option("sweep_warps", 16)
        ctx.set_option("table_mode", -1)
    ref, _ = oracle.score_bits_batch(P, Y, p, T[:64], G, nthreads=os.cpu_count() or 1)
    assert_hits_match(np.ascontiguousarray(dense[:, :64]).reshape(-1).copy(), ref, f"G=8192, {warps} warps")
    for q in range(nq):
        exp = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], exp[f]), (q, f)


def test_config2_full_size_is_bit_exact_against_oracle(ctx):
    """BASELINE config 2 at its full size (1 query x 1,000,000 targets x 1024 genomes, dense): EVERY pair against the
    oracle -- counts and cut-off indices bit-exact, scores bit-identical (exact table mode; 1 ulp only on the
    host-libm expm1 branch)."""
    G, nt = 1024, 1_000_000
    P, Y, p = synth.make_queries(G, 1, synth.SEED_BASE + 2, "single")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 102, plant_from=Y, plant_frac=0.01)
    ctx.set_targets_bits(T, G)
    hits = ctx.score(P, Y, p)
    assert ctx.counter("used_table") == 1
    ref, _ = oracle.score_bits_batch(P, Y, p, T, G, nthreads=os.cpu_count() or 1)
    ref = ref.reshape(-1)
    for f in ("yes", "number", "depth"):
        assert np.array_equal(hits[f], ref[f]), f
    same = hits["score"] == ref["score"]
    ulp = np.abs(hits["score"].view(np.int64) - ref["score"].view(np.int64))
    assert np.all(same | ((ulp <= 1) & (hits["yes"] == 1))), int((~same).sum())
    assert same.mean() > 0.999


def test_config3_per_gpu_share_properties_and_sampled_oracle(ctx):
    """BASELINE config 3's per-GPU share at full size (10,000 queries x 125,000 targets x 4096 genomes, top-100; the
    bench workload): lists complete and sorted; identical with the speculative threshold switched off; a sample of
    reported hits equals the oracle's record for that pair; and no target of a dense-scored sample that is missing from
    a list beats the list's last entry."""
    G, nq, nt, k = 4096, 10000, 125000, 100
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
    ctx.set_targets_bits(T, G)
    flt = capi.make_filter(capi.FILTER_TOPK, k=k)
    a = ctx.score(P, Y, p, flt).copy()
    assert ctx.counter("spec_rank") >= 1
    assert a.size == nq * k and np.array_equal(a["q"], np.repeat(np.arange(nq, dtype=np.uint32), k))
    sc = a["score"].reshape(nq, k)
    tt = a["t"].reshape(nq, k)
    assert np.all(np.diff(sc, axis=1) >= 0) and tt.max() < nt
    tie = np.diff(sc, axis=1) == 0
    assert np.all(np.diff(tt.astype(np.int64), axis=1)[tie] > 0)  # ties by ascending target index
    ctx.set_option("speculate", 0)
    try:
        b = ctx.score(P, Y, p, flt).copy()
    finally:
        ctx.set_option("speculate", 1)
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(a[f], b[f]), f
    rng = np.random.default_rng(3)
    for i in rng.choice(a.size, 48, replace=False):  # sampled oracle check of reported hits
        h = a[i]
        r, _ = oracle.score_bits_batch(P[h["q"]:h["q"] + 1], Y[h["q"]:h["q"] + 1], p[h["q"]:h["q"] + 1], T[h["t"]:h["t"] + 1], G)
        r = r.reshape(-1)[0]
        assert (h["yes"], h["number"], h["depth"]) == (r["yes"], r["number"], r["depth"]), i
        assert h["score"] == r["score"] or abs(np.log(h["score"]) - np.log(r["score"])) <= REL_TOL_NEGLOG * abs(np.log(r["score"])), i
    qs = rng.choice(nq, 16, replace=False)  # nothing outside the lists beats them (dense sample of targets)
    ts = np.sort(rng.choice(nt, 3000, replace=False))
    ctx.set_targets_bits(np.ascontiguousarray(T[ts]), G)
    ctx.set_option("table_mode", 0)
    try:
        d = ctx.score(np.ascontiguousarray(P[qs]), np.ascontiguousarray(Y[qs]), np.ascontiguousarray(p[qs])).reshape(len(qs), len(ts))
    finally:
        ctx.set_option("table_mode", -1)
    for j, q in enumerate(qs):
        last_s, last_t = sc[q, -1], tt[q, -1]
        inlist = set(tt[q].tolist())
        for s, t in zip(d[j]["score"], ts):
            if int(t) not in inlist:
                assert (s, t) > (last_s, last_t), (q, t, s, last_s)


def test_config5_full_size_threshold_run_ranked_on_device(ctx):
    """BASELINE config 5 at full size (4,500 x 4,500 profile families x 8192 genomes, the ppp_cutoff.pl filter: every pair
    whose printed log10 score is < 0): the device-ranked hit list is exactly the dense result's passing pairs -- same
    records, same count -- in (query, score, target) order, and the device's printed scores / CUTOFFs are consistent."""
    G, n = 8192, 4500
    P, Y, p = synth.make_queries(G, n, synth.SEED_BASE + 5, "family")
    ctx.set_targets_bits(Y, G)
    dense = ctx.score(P, Y, p).reshape(n, n).copy()
    hits = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-0.0005))
    with np.errstate(divide="ignore"):
        keep = np.log10(dense["score"]) < -0.0005
    assert hits.size == int(keep.sum()) and hits.size > 10_000_000
    d = dense[hits["q"], hits["t"]]
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(d[f], hits[f]), f
    q, s, t = hits["q"].astype(np.int64), hits["score"], hits["t"].astype(np.int64)
    same_q = q[1:] == q[:-1]
    assert np.all(q[1:] >= q[:-1])
    assert np.all(s[1:][same_q] >= s[:-1][same_q])
    tie = same_q & (s[1:] == s[:-1])
    assert np.all(t[1:][tie] > t[:-1][tie])
    r = ctx.printed_cutoffs(20.0)
    off = r["row_off"].astype(np.int64)
    assert np.array_equal(off[1:] - off[:-1], np.bincount(q, minlength=n))
    ok = r["status"] == capi.CUT_OK
    assert ok.sum() > n // 4  # the diagonal's exact zeros put the other queries into CUT_LOG0, as the reference would die
    zero_q = np.unique(q[s == 0.0])
    assert np.all(r["status"][zero_q] == capi.CUT_LOG0)
    m = r["milli"]
    rows_ok = ok[q]
    with np.errstate(divide="ignore"):
        exp = np.rint(np.log10(s[rows_ok]) * 1000.0)
    assert np.max(np.abs(m[rows_ok] - exp)) <= 1  # numpy's log10 is not the reference's log/log(10): one unit at a boundary
    assert np.all(np.diff(m.astype(np.int64))[same_q & rows_ok[1:] & rows_ok[:-1]] >= 0)
    cm = r["cutoff_milli"][ok]
    assert np.all((cm == capi.CUT_NONE) | (cm <= -1000))


def _devices_for_group(n):
    """n device ordinals for a multi-device context: distinct GPUs when the box has them, else the same GPU n times (every
    entry of the list owns its own shard, stream and buffers, so the group code paths are the same)"""
    import torch

    have = torch.cuda.device_count()
    return [i % max(have, 1) for i in range(n)]


@pytest.mark.parametrize("ndev", [2, 3])
def test_multi_device_context_equals_single_device(ctx, ndev):
    """ppp_init(devices, ndev > 1): ONE process drives several GPUs -- targets sharded over the devices, every filter mode's
    result assembled by the library (dense rows by strided copies, threshold hits by a per-query merge, top-k lists merged on
    the devices by query slices over peer copies).  Must equal the single-device result record for record: same kernels per
    pair, so counts, cut-off indices AND scores are identical."""
    G, nq, nt, k = 1024, 37, 1203, 7   # uneven shards (1203 / 3) and uneven query slices (37 / 2, 37 / 3)
    P, Y, p = synth.make_queries(G, nq, 501, "mixed")
    T = synth.make_targets(G, nt, 502, plant_from=Y, plant_frac=0.05)
    grp = capi.Context(_devices_for_group(ndev))
    try:
        for c in (ctx, grp):
            c.set_option("table_mode", 0)
        ctx.set_targets_bits(T, G)
        grp.set_targets_bits(T, G)
        filters = [None, capi.make_filter(capi.FILTER_TOPK, k=k), capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-0.3),
                   capi.make_filter(capi.FILTER_TOPK, k=2000)]  # k > targets per device and > nt: short lists
        for flt in filters:
            a = ctx.score(P, Y, p, flt).copy()
            b = grp.score(P, Y, p, flt).copy()
            assert a.size == b.size and a.size > 0, (flt.mode if flt else 0, a.size, b.size)
            for f in ("q", "t", "yes", "number", "depth"):
                assert np.array_equal(a[f], b[f]), (ndev, flt.mode if flt else 0, f)
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = np.abs(np.log(a["score"]) - np.log(b["score"])) / np.abs(np.log(a["score"]))
            assert np.all((a["score"] == b["score"]) | (rel <= REL_TOL_NEGLOG))
        assert grp.stats()["pairs"] == nq * nt and grp.counter("full_pairs") > 0
        # ranked targets (own list order, duplicates, foreign taxa) over the group
        rng = np.random.default_rng(9)
        lists = [list(rng.integers(0, G + 40, size=int(rng.integers(0, 300)))) for _ in range(101)]
        from prophylo_b200 import packer
        index = {g: g for g in range(G)}
        idx, raw, off = packer.pack_ranked_targets(lists, index)
        ctx.set_targets_ranked(idx, raw, off, G)
        grp.set_targets_ranked(idx, raw, off, G)
        a, b = ctx.score(P, Y, p).copy(), grp.score(P, Y, p).copy()
        for f in ("q", "t", "score", "yes", "number", "depth"):
            assert np.array_equal(a[f], b[f]), (ndev, "ranked", f)
        # fewer targets than devices
        grp.set_targets_bits(T[:1], G)
        ctx.set_targets_bits(T[:1], G)
        for flt in (None, capi.make_filter(capi.FILTER_TOPK, k=3)):
            a, b = ctx.score(P, Y, p, flt).copy(), grp.score(P, Y, p, flt).copy()
            for f in ("q", "t", "score", "yes", "number", "depth"):
                assert np.array_equal(a[f], b[f]), (ndev, "one target", f)
    finally:
        ctx.set_option("table_mode", -1)
        grp.close()


def test_multi_device_context_at_config3_share_equals_single_device(ctx):
    """the group's top-k merge at the bench shape: 10,000 queries x 125,000 targets x 4096 genomes over two shards equals one
    device's lists (targets, counts, cut-off indices exactly; scores within 1e-9 on -ln p: the speculative schedule differs
    per shard, so a pair may take the series tier in one run and the exact tier in the other)."""
    G, nq, nt, k = 4096, 10000, 125000, 100
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
    flt = capi.make_filter(capi.FILTER_TOPK, k=k)
    ctx.set_targets_bits(T, G)
    a = ctx.score(P, Y, p, flt).copy()
    grp = capi.Context(_devices_for_group(2))
    try:
        grp.set_targets_bits(T, G)
        b = grp.score(P, Y, p, flt).copy()
    finally:
        grp.close()
    assert a.size == b.size == nq * k
    for f in ("q", "t", "yes", "number", "depth"):
        assert np.array_equal(a[f], b[f]), f
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(np.log(a["score"]) - np.log(b["score"])) / np.abs(np.log(a["score"]))
    assert np.all((a["score"] == b["score"]) | (rel <= REL_TOL_NEGLOG))


@pytest.mark.parametrize("G,kind,nq,nt,k", [(4096, "mixed", 600, 20000, 20), (1024, "mixed", 300, 9000, 10), (8192, "family", 200, 6000, 12),
                                            (2048, "prefix", 300, 12000, 2)])
def test_tensor_core_chunk_screen_equals_popc_screen(ctx, G, kind, nq, nt, k):
    """The tcgen05 chunk screen (option prefilter_tc, ppp_prefilter_tc.cu: T.Y counts per 128-genome chunk accumulated in tensor
    memory by tcgen05.mma kind::i8) against the POPC screen: identical top-k lists record for record, identical SET of fully
    evaluated pairs (counter full_pairs), for both launches (all-ones and general presence planes), with the in-kernel exact
    re-screen and with the pair-level (superset) survivor list."""
    P, Y, p = synth.make_queries(G, nq, 900 + G, kind)
    T = synth.make_targets(G, nt, 901 + G, plant_from=Y, plant_frac=0.01)
    ctx.set_targets_bits(T, G)
    flt = capi.make_filter(capi.FILTER_TOPK, k=k)
    try:
        ctx.set_option("prefilter_tc", 0)
        base = ctx.score(P, Y, p, flt).copy()
        full = ctx.counter("full_pairs")
        surv = ctx.counter("survivors")
        assert ctx.counter("tc_launches") == 0
        for tc, exact in ((1, 1), (3, 1), (3, 0)):
            ctx.set_option("prefilter_tc", tc)
            ctx.set_option("tc_exact", exact)
            got = ctx.score(P, Y, p, flt).copy()
            # G = 8192 family tiles hold more staircase entries than the kernel's shared memory: the launch falls back to POPC
            assert ctx.counter("tc_launches") > 0 or G == 8192, (tc, exact)
            for f in ("q", "t", "score", "yes", "number", "depth"):
                assert np.array_equal(base[f], got[f]), (tc, exact, f)
            assert ctx.counter("full_pairs") == full, (tc, exact)
            if exact:
                assert ctx.counter("survivors") <= surv * 1.01 + 64  # exact but for queue overflows, which accept unscreened
            else:
                assert ctx.counter("survivors") >= surv
    finally:
        ctx.set_option("prefilter_tc", 0)
        ctx.set_option("tc_exact", 1)


@pytest.mark.parametrize("G,nq,nt,ndev", [(4096, 300, 150000, 1), (1000, 64, 70001, 1), (4096, 200, 150000, 2)])
def test_streamed_target_upload_equals_plain_upload(ctx, G, nq, nt, ndev):
    """ppp_set_targets_bits_streamed: the planes cross PCIe in pieces while the scan of the first pieces is already running
    (every launch waits for -- and transposes -- only the pieces it reads).  Same kernels on the same rows as the plain upload,
    so every filter mode must return identical records; the host buffer is page-locked once and pageable once; the planes are
    replaced between runs (the buffers are reused, and an upload still in flight is drained first)."""
    import torch
    P, Y, p = synth.make_queries(G, nq, 700 + G, "mixed")
    T = synth.make_targets(G, nt, 701 + G, plant_from=Y, plant_frac=0.01)
    T2 = synth.make_targets(G, nt // 3, 702 + G, plant_from=Y, plant_frac=0.01)
    Tpin = torch.from_numpy(T).pin_memory().numpy()
    dut = ctx if ndev == 1 else capi.Context(_devices_for_group(ndev))
    try:
        filters = [capi.make_filter(capi.FILTER_TOPK, k=50), capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-6.0)]
        if nq * nt <= 1 << 23:
            filters.append(None)
        for flt in filters:
            ctx.set_targets_bits(T, G)
            want = ctx.score(P, Y, p, flt).copy()
            for host in (Tpin, T):
                dut.set_targets_bits(host, G, streamed=True)
                got = dut.score(P, Y, p, flt).copy()
                assert got.size == want.size and want.size > 0
                if flt is None and ndev == 1:
                    assert dut.counter("used_table") == 1  # dense run in exact-table mode: one sweep launch per arriving piece
                for f in ("q", "t", "yes", "number", "depth") + (("score",) if ndev == 1 else ()):
                    assert np.array_equal(got[f], want[f]), (flt.mode if flt else 0, f)
            # replaced while nothing has consumed the previous streamed upload; then a plain upload over a streamed one
            dut.set_targets_bits(Tpin, G, streamed=True)
            dut.set_targets_bits(T2, G, streamed=True)
            a = dut.score(P, Y, p, flt).copy()
            dut.set_targets_bits(Tpin, G, streamed=True)
            dut.wait_targets()
            dut.set_targets_bits(T2, G)
            b = dut.score(P, Y, p, flt).copy()
            ctx.set_targets_bits(T2, G)
            c = ctx.score(P, Y, p, flt).copy()
            for f in ("q", "t", "yes", "number", "depth"):
                assert np.array_equal(a[f], c[f]) and np.array_equal(b[f], c[f]), (flt.mode if flt else 0, f)
    finally:
        if dut is not ctx:
            dut.close()


@pytest.mark.parametrize("G,nq,nt,k", [(4096, 90, 9000, 50), (1000, 40, 20000, 10), (8192, 20, 3000, 5)])
def test_p_equals_y_class_matches_dense(ctx, G, nq, nt, k):
    """Queries with P == Y (the cross-score form, p = |Y| / G) have their own pre-filter launch: number == yes at every
    candidate, so the pair is screened with ONE staircase lookup at the row's total popc(T & Y) (ppp_prefilter.cu, MODE 2).  Top-k
    lists and threshold hits must equal the dense result; a query with P == Y == all ones stays in the all-ones class; queries
    of the three classes are mixed in one run."""
    rng = np.random.default_rng(5 + G)
    P, Y, p = synth.make_queries(G, nq, 600 + G, "family")
    Yb = synth.unpack_bits(Y, G)
    Pb = Yb.copy()                        # P == Y ...
    Pb[::7] = 1                           # ... except every 7th query: all-ones presence plane
    Pb[3::7] |= rng.random((Pb[3::7].shape[0], G)) < 0.5   # ... and some general planes (Y subset of P, P neither all ones nor Y)
    Yb[5] = 1
    Pb[5] = 1                             # P == Y == all ones
    P, Y = synth.pack_bits(Pb), synth.pack_bits(Yb)
    p = Yb.sum(axis=1) / np.where((Pb == Yb).all(axis=1), G, Pb.sum(axis=1))
    p = np.clip(p, 1e-6, 0.999)
    T = synth.make_targets(G, nt, 601 + G, plant_from=Y, plant_frac=0.03)
    ctx.set_targets_bits(T, G)
    ctx.set_option("table_mode", 0)
    try:
        dense = ctx.score(P, Y, p).reshape(nq, nt)
    finally:
        ctx.set_option("table_mode", -1)
    tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    assert tk.size == nq * k and ctx.counter("full_pairs") < nq * nt
    for q in range(nq):
        exp = np.sort(dense[q], order=["score", "t"])[:k]
        got = tk[q * k:(q + 1) * k]
        for f in ("q", "t", "score", "yes", "number", "depth"):
            assert np.array_equal(got[f], exp[f]), (q, f)
    thr = -3.0
    th = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=thr)).copy()
    with np.errstate(divide="ignore"):
        keep = np.log10(dense["score"]) < thr
    exp = np.sort(dense[keep], order=["q", "score", "t"])
    assert th.size == exp.size and th.size > 0
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(th[f], exp[f]), f


def test_dense_yes_planes_run_with_smaller_prefilter_tiles(ctx):
    """Very dense yes planes (|Y| ~ 4,100 of 8,192 genomes): the staircases of a full 32-query (16-query) tile do not fit the
    pre-filter's shared memory, so ppp_set_queries gives the class smaller tiles -- one code path for every query set.  All
    three classes (all-ones presence planes, P == Y, general), top-k (speculative schedule, re-scan tiles) and threshold runs
    against the dense result."""
    G, nq, nt, k = 8192, 44, 2600, 8
    rng = np.random.default_rng(77)
    Yb = (rng.random((nq, G)) < 0.5).astype(np.uint8)
    Pb = np.ones((nq, G), dtype=np.uint8)
    Pb[1::5] = Yb[1::5]                                             # P == Y
    Pb[2::5] = Yb[2::5] | (rng.random((Yb[2::5].shape[0], G)) < 0.5)  # general
    P, Y = synth.pack_bits(Pb), synth.pack_bits(Yb)
    p = Yb.sum(axis=1) / np.where((Pb == Yb).all(axis=1), G, Pb.sum(axis=1))
    T = synth.make_targets(G, nt, 78, plant_from=Y, plant_frac=0.05)
    ctx.set_targets_bits(T, G)
    ctx.set_option("table_mode", 0)
    try:
        dense = ctx.score(P, Y, p).reshape(nq, nt)
    finally:
        ctx.set_option("table_mode", -1)
    old = None
    for lp in (0, 16):  # one launch / several launches per run
        ctx.set_option("launch_pairs_log2", lp)
        try:
            tk = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k)).copy()
        finally:
            ctx.set_option("launch_pairs_log2", 0)
        assert 1 <= ctx.counter("pf_tile_q_all") < 32 and 1 <= ctx.counter("pf_tile_q_pyy") < 32 and 1 <= ctx.counter("pf_tile_q_gen") < 16
        assert tk.size == nq * k and ctx.counter("full_pairs") < nq * nt
        for q in range(nq):
            exp = np.sort(dense[q], order=["score", "t"])[:k]
            got = tk[q * k:(q + 1) * k]
            for f in ("q", "t", "score", "yes", "number", "depth"):
                assert np.array_equal(got[f], exp[f]), (lp, q, f)
        assert old is None or old.tobytes() == tk.tobytes()
        old = tk
    thr = -6.0
    th = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=thr)).copy()
    with np.errstate(divide="ignore"):
        keep = np.log10(dense["score"]) < thr
    exp = np.sort(dense[keep], order=["q", "score", "t"])
    assert th.size == exp.size and th.size > 0
    for f in ("q", "t", "score", "yes", "number", "depth"):
        assert np.array_equal(th[f], exp[f]), f


def test_targets_from_device_memory_equal_host_upload(ctx):
    """ppp_set_targets_bits_device: planes that already sit in device memory (what an NCCL all-gather of per-rank slices leaves
    behind) give the same records as the host upload -- dense and top-k, G with and without row padding, single context and
    device group; a host pointer is refused."""
    import torch

    for G, nq, nt in ((4096, 40, 3000), (1000, 24, 2500)):
        P, Y, p = synth.make_queries(G, nq, 31 + G, "mixed")
        T = synth.make_targets(G, nt, 32 + G, plant_from=Y, plant_frac=0.03)
        Td = torch.from_numpy(T.view(np.int32)).cuda()
        torch.cuda.synchronize()
        grp = capi.Context(_devices_for_group(2))
        try:
            for c in (ctx, grp):
                for flt in (None, capi.make_filter(capi.FILTER_TOPK, k=7)):
                    c.set_targets_bits(T, G)
                    a = c.score(P, Y, p, flt).copy()
                    c.set_targets_bits_device(Td.data_ptr(), nt, G)
                    b = c.score(P, Y, p, flt).copy()
                    assert a.size > 0 and a.tobytes() == b.tobytes()
        finally:
            grp.close()
    with pytest.raises(capi.PPPError):
        ctx.set_targets_bits_device(T.ctypes.data, nt, G)
