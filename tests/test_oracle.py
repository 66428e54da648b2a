"""CPU tests: the oracle (C restatement of ppp.pm:130-245 + Math::Cephes 0.47 bdtrc) against the golden vectors
that oracle/gen_golden.py produced by running the reference itself (perl + the reference's bundled Cephes binary)."""
import json
import os

import numpy as np

from oracle import oracle
from prophylo_b200 import synth


def test_bdtrc_points_match_reference_binary(golden_dir):
    pts = json.load(open(os.path.join(golden_dir, "bdtrc_points.json")))["points"]
    assert len(pts) > 3000
    k = np.array([x[0] for x in pts], dtype=np.int32)
    n = np.array([x[1] for x in pts], dtype=np.int32)
    p = np.array([float(x[2]) for x in pts])
    exp = np.array([float(x[3]) for x in pts])
    got = oracle.bdtrc_batch(k, n, p)
    # k == 0 and p < 0.01 binds to the host libm's expm1 in the reference (see oracle/cephes_port.c): allow 1 ulp there
    libm = (k == 0) & (p < 0.01)
    assert np.array_equal(got[~libm].view(np.uint64), exp[~libm].view(np.uint64))
    assert np.all(np.abs(got[libm] - exp[libm]) <= np.spacing(exp[libm]))


def test_survey_table_values():
    # SURVEY.md section 8c `bdtrc` points (binary)
    assert oracle.bdtrc(0, 1, .02) == 0.020000000000000018
    assert oracle.bdtrc(8, 12, 9 / 434) == 1.4745387719095076e-13
    assert oracle.bdtrc(3, 100, .25) == 0.9999999978918347
    assert oracle.bdtrc(49, 50, .5) == 8.881784197001258e-16
    assert oracle.bdtrc(99, 4096, .01) == 3.271936483402975e-15
    assert oracle.bdtrc(500, 4096, .1) == 2.125572110772742e-06
    assert oracle.bdtrc(-1, 5, .3) == 1.0
    assert oracle.bdtrc(5, 5, .3) == 0.0
    assert oracle.bdtrc(150, 4096, 1e-4) == 0.0
    assert oracle.bdtrc(0, 5, .01) == 0.04900995009999998
    assert oracle.bdtrc(2, 5, 1.2) == 0.0  # domain error -> 0.0


def test_list_cases_match_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ppp_list_cases.json")))
    cases = g["known_answers"] + g["fuzz"]
    assert len(cases) >= 600
    for c in cases:
        got = oracle.score_case(c["query"], c["target"], c.get("prob"), bool(c.get("dup")))
        e = c["expect"]
        assert got == (float(e[0]), e[1], e[2], e[3], float(e[4])), c["name"]


def test_known_answer_1_is_baseline_config_1(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ppp_list_cases.json")))
    ka1 = g["known_answers"][0]
    assert ka1["query_yes_total"] == [9, 434]  # t/00profile.t:10-14
    assert ka1["expect"] == ["1", 0, 0, 0, "0.020737327188940093"]


def test_bit_cases_match_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ppp_bit_cases.json")))
    for case in g["cases"]:
        sp = case["spec"]
        G = sp["G"]
        P, Y, p = synth.make_queries(G, sp["nq"], sp["qseed"], sp["kind"])
        T = synth.make_targets(G, sp["nt"], sp["tseed"], plant_from=Y, plant_frac=0.25)
        assert [repr(float(x)) for x in p] == case["p"]
        out, _ = oracle.score_bits_batch(P, Y, p, T, G, nthreads=4)
        res = case["results"]
        for q in range(sp["nq"]):
            for t in range(sp["nt"]):
                r = res[q * sp["nt"] + t]
                o = out[q, t]
                assert (o["score"], o["yes"], o["number"], o["depth"]) == (float(r[0]), r[1], r[2], r[3]), (sp, q, t)


def _dup_by_packed_list(query, tlist, prob):
    """The -dup result computed the way the GPU path does it: the packer's list (first repeated entry kept, list cut
    at the second, packer.pack_ranked_targets(dup=True)) scored WITHOUT the flag, the kept repeat standing in as a
    clone of its taxon; the reference's `last if $yes > $max_yes` (ppp.pm:222-224) is part of the oracle loop."""
    from prophylo_b200 import packer
    universe = list(query)
    index = packer.universe_index(universe)
    idx, raw, off = packer.pack_ranked_targets([tlist], index, dup=True)
    isyes = [1 if float(query[t]) != 0 else 0 for t in universe]
    present = [1] * len(universe)
    ids, seen = [], set()
    for g in idx.tolist():
        if g == packer.FOREIGN:
            ids.append(len(present) + 100000 + len(ids))  # foreign to the query: never present
            continue
        if g in seen:  # the re-counted entry: a clone with the same membership
            present.append(1)
            isyes.append(isyes[g])
            g = len(present) - 1
        seen.add(g)
        ids.append(g)
    yes = sum(isyes[:len(universe)])
    p = float(prob) if prob else yes / len(universe)
    r = oracle.score_list(present, isyes, yes, ids, p, False)
    depth = int(raw[r[3] - 1]) if r[3] else 0
    return (r[0], r[1], r[2], depth, p)


def test_dup_packing_rule_matches_reference(golden_dir):
    """-dup (ppp.pm:189-200 literal-key quirk) as the packer embeds it == the reference, on every golden -dup case and
    on random lists with many repeats (oracle with the flag, itself pinned above)."""
    g = json.load(open(os.path.join(golden_dir, "ppp_list_cases.json")))
    n = 0
    for c in g["known_answers"] + g["fuzz"]:
        if not c.get("dup") or isinstance(c["target"], dict):
            continue
        e = c["expect"]
        got = _dup_by_packed_list(c["query"], c["target"], c.get("prob"))
        assert got == (float(e[0]), e[1], e[2], e[3], float(e[4])), c["name"]
        n += 1
    assert n >= 90
    rng = np.random.default_rng(5)
    for _ in range(400):
        U = int(rng.integers(3, 40))
        taxa = [f"t{i}" for i in range(U)]
        query = {t: int(rng.random() < 0.4) for t in taxa}
        if not any(query.values()):
            query[taxa[0]] = 1
        pool = taxa + [f"x{i}" for i in range(4)]
        tlist = [pool[int(i)] for i in rng.integers(0, len(pool), int(rng.integers(1, 60)))]
        if rng.random() < 0.3:  # every yes taxon first, then a repeated yes taxon: yes exceeds max_yes
            ys = [t for t in taxa if query[t]]
            tlist = ys + [ys[0]] + tlist
        prob = [None, 0.3, 0.05][int(rng.integers(0, 3))]
        want = oracle.score_case(query, tlist, prob, True)
        assert _dup_by_packed_list(query, tlist, prob) == want, (query, tlist, prob)


def test_rank_cases_match_reference(golden_dir):
    """-rank (ppp.pm:328-370): the reference run with a mock -taxonomy object (oracle/gen_golden_rank.py) against the
    oracle's restatement of the collapse pre-pass; and the croak of ppp.pm:345-349 without a taxonomy."""
    import pytest

    g = json.load(open(os.path.join(golden_dir, "ppp_rank_cases.json")))
    assert len(g["cases"]) >= 100
    differs = 0
    for c in g["cases"]:
        got = oracle.score_case(c["query"], c["target"], c.get("prob"), bool(c.get("dup")), rank=c["rank"], collapse=c["collapse"])
        e = c["expect"]
        assert got == (float(e[0]), e[1], e[2], e[3], float(e[4])), c["name"]
        differs += oracle.score_case(c["query"], c["target"], c.get("prob"), bool(c.get("dup")))[:4] != got[:4]
    assert differs > 50  # the collapse matters in these cases
    c = g["croak"]
    with pytest.raises(RuntimeError, match=c["error"]):
        oracle.score_case(c["query"], c["target"], rank=c["rank"])


def test_memoised_batch_equals_plain_batch():
    """The per-query bdtrc cache of the full-size parity tests changes no result (same loop, pure function)."""
    for G, kind in ((1024, "mixed"), (4096, "mixed"), (2048, "prefix"), (8192, "family")):
        P, Y, p = synth.make_queries(G, 5, 300 + G, kind)
        T = synth.make_targets(G, 150, 400 + G, plant_from=Y, plant_frac=0.1)
        a, sa = oracle.score_bits_batch(P, Y, p, T, G, nthreads=4)
        b, sb = oracle.score_bits_batch(P, Y, p, T, G, nthreads=4, memo=True)
        assert sa == sb and a.tobytes() == b.tobytes(), (G, kind)
