"""Host-side drivers around the GPU scan: the parts of bin/ppp.pl and bin/ppp_cutoff.pl that sit either side of the hot
path -- `.bla` hit-list reader, batch scoring of one query profile against a genome's genes,
the stdout table of ppp.pl and the slope filter of ppp_cutoff.pl -- with the reference's formats kept byte for byte.

    rows = ppp_scan(ctx, Profile(file="q.profile"), {"500010": "db/blast/1140/500010.bla", ...})
    print(format_ppp_table(rows, slope=30), end="")          # == stdout of  ppp.pl -t 1140 -p q.profile -m 30
    print(ppp_cutoff(table_text, slope=20)[0], end="")       # == stdout of  ppp_cutoff.pl -s 20
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import capi, packer
from .algorithm import Profile, perl_num, perl_split_tab


def read_bla(path):
    """One gene's cached BLAST hits -> the ordered taxon list the reference scores (bin/ppp.pl:606-673 +
    lib/PhyloProf/BlastResult.pm:192-237): 4 tab columns query, subject, e-value, subject taxon; subjects keep their
    first-occurrence order, a subject's LAST taxon wins, subjects with a false taxon ('' or '0') are dropped,
    different subjects of the same taxon stay as duplicates (they count in the depth)."""
    order, taxon = [], {}
    with open(path, newline="\n") as fh:  # Perl's chomp removes "\n" only
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) < 4:
                f += [""] * (4 - len(f))
            if f[1] not in taxon:
                order.append(f[1])
            taxon[f[1]] = f[3]
    return [taxon[s] for s in order if taxon[s] not in ("", "0")]


def ppp_scan(ctx: capi.Context, query: Profile, targets: dict, prob=None, dup=False, level=None, taxonomy=None):
    """Score one query profile against every target hit list in ONE GPU call (replaces the chunk files, forked
    `ppp.pl --client` workers and per-gene Perl sweeps of bin/ppp.pl:354-462, 519-593).
    targets: {locus: path of a .bla file | ordered taxon list}.  dup: the --dup flag of bin/ppp.pl:282 (ppp.pm:189-200,
    see packer.pack_ranked_targets).  level / taxonomy: the -l <rank> option (bin/ppp.pl:532-547, 566-568): every taxon of
    every hit list is collapsed to that rank on the host before packing (ppp.pm:328-370); the caller builds `query` with the
    same rank and taxonomy, as the reference's client does.  Returns [(locus, score, yes, number, depth, p)]."""
    from .algorithm import collapse

    q = query.get_hash()
    loci = list(targets)
    p = float(prob) if prob else query.get_yes() / query.get_total()  # bin/ppp.pl:559-572, ppp.pm:138-142
    if not loci:
        return []
    index = packer.universe_index(list(q.keys()))
    P, Y, _ = packer.pack_queries([q], index)
    if level:
        if taxonomy is None:
            raise RuntimeError("Taxonomy does not exist")  # ppp.pm:345-349
        lists = [[collapse(taxonomy, level, t) for t in (read_bla(x) if isinstance(x, (str, os.PathLike)) else list(x))] for x in targets.values()]
        idx, raw, off = packer.pack_ranked_targets(lists, index, dup=bool(dup))
    elif all(isinstance(t, (str, os.PathLike)) for t in targets.values()):
        # a genome's worth of cache files: the native multi-threaded reader (csrc/ppp_bla.cpp) parses, de-duplicates and
        # maps them straight into the arrays of ppp_set_targets_ranked
        idx, raw, off = capi.read_bla_files(list(targets.values()), list(index), dup=bool(dup))
    else:
        lists = [read_bla(t) if isinstance(t, (str, os.PathLike)) else list(t) for t in targets.values()]
        idx, raw, off = packer.pack_ranked_targets(lists, index, dup=bool(dup))
    ctx.set_targets_ranked(idx, raw, off, len(index))
    ctx.set_option("dup", int(bool(dup)))
    try:
        hits = ctx.score(P, Y, np.array([p], dtype=np.float64))
    finally:
        ctx.set_option("dup", 0)
    return [(loci[int(h["t"])], float(h["score"]), int(h["yes"]), int(h["number"]), int(h["depth"]), p) for h in hits]


def _perl_num(x: float) -> float:
    """The client writes $score / $p with Perl's default %.15g stringification before the server reads them back
    (bin/ppp.pl:582, 443-462)."""
    return float("%.15g" % x)


def _log10_3f(score: float) -> str:
    if score == 0:
        raise ValueError("Can't take log of 0")  # bin/ppp.pl:476 dies the same way
    return "%.3f" % (math.log(score) / math.log(10))


def _handle_negative_zero(s: str) -> str:
    # bin/ppp.pl:595-604: "-0.000" -> "0.000"
    if s.startswith("-") and set(s[1:]) <= set("0."):
        return s[1:]
    return s


def format_ppp_table(rows, slope=None, desc=None) -> str:
    """stdout of the ppp.pl server (bin/ppp.pl:468-506): header, rows sorted ascending by raw score, `%.3f` Prob and
    log10 Score, optional slope marker line.  Ties in the raw score are ordered by locus here (hash order in Perl)."""
    out = ["Locus\t#Yes\t#Total\tDepth\tProb\tScore\tDesc\n"]
    rows = sorted(((loc, _perl_num(s), y, n, d, _perl_num(p)) for loc, s, y, n, d, p in rows), key=lambda r: (r[1], str(r[0])))
    last_score = None
    cutoff_found = False
    for loc, s, y, n, d, p in rows:
        log_score = _log10_3f(s)
        if last_score is None:
            last_score = log_score  # set once to the best row, never updated (bin/ppp.pl:477)
        if slope and not cutoff_found:
            pl, ps = abs(float(last_score)), abs(float(log_score))
            if pl == 0:
                raise ZeroDivisionError("Illegal division by zero")
            if (pl - ps) / pl * 100 > float(slope):
                out.append("-------------------------------------------\n")
                cutoff_found = True
        out.append(f"{loc}\t{y}\t{n}\t{d}\t{'%.3f' % p}\t{_handle_negative_zero(log_score)}\t{(desc or {}).get(loc, '')}\n")
    return "".join(out)


def _field5(line: str) -> float:
    """$f[5] of a table row in numeric context: a missing field is undef = 0 (bin/ppp_cutoff.pl:122, 153)."""
    f = perl_split_tab(line)
    return perl_num(f[5]) if len(f) > 5 else 0.0


def ppp_cutoff(table_text: str, slope: float):
    """bin/ppp_cutoff.pl:111-162 on the text of a ppp.pl table.  Returns (stdout text, cutoff or None)."""
    lines = table_text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    header = lines[0] + "\n" if lines else ""
    kept, scores = [], []
    for line in lines[1:]:
        if line and set(line) == {"-"}:
            continue
        kept.append(line)
        if not line.startswith("#"):
            v = _field5(line)
            if v > -0.001:
                continue
            scores.append(v)
    cutoff = None
    last = None
    for s in sorted(scores, reverse=True):
        if s > -1:
            continue
        if not last:
            last = s
            continue
        if abs((last - s) / last * 100) > slope:
            cutoff = last
            break
        last = s
    out = [header, "Significant hits:\n"]
    marker = False
    for line in kept:
        v = _field5(line)
        if v >= 0:
            continue
        if v > (cutoff if cutoff is not None else 0.0) and not marker:
            out.append("Insignificant hits:\n")
            marker = True
        out.append(line + "\n")
    return "".join(out), cutoff


def ppp2d_prefix_queries(gi: str, bla_path, taxdis, start=5, end=None, skip=1, total_taxa=1459):
    """The depth-prefix query profiles of bin/ppp2d.pl:253-315 for one gene: walk its `.bla` lines; the self hit only
    marks its taxon; a subject whose taxon was already seen is skipped; at every remaining line (subject to --skip,
    --start-depth, --end-depth) the profile is {seen taxa: 1, every other taxon of data/taxdis.txt: 0} and
    p = taxon_count / 1459 (the reference's hard-coded $total_taxa_count, passed through `--prob` as a %.15g string).
    Returns [(line_count, taxon_count, profile dict, p)]."""
    out, seen, line_count = [], {}, 0
    with open(bla_path) as fh:
        for line in fh:
            line_count += 1
            f = line.rstrip("\n").split("\t")
            if f[0] == f[1]:
                seen[f[3]] = 1
                continue
            if f[3] in seen:
                continue
            seen[f[3]] = 1
            if line_count % skip:
                continue
            tc = len(seen)
            if end and tc > end:
                break
            if tc >= start:
                prof = {t: 1 for t in seen}
                for t in taxdis:
                    if t not in seen:
                        prof[t] = 0
                out.append((line_count, tc, prof, float("%.15g" % (tc / total_taxa))))
    return out


def ppp2d_scan(ctx: capi.Context, gi: str, bla_path, genome_targets: dict, taxdis, start=5, end=None, skip=1, total_taxa=1459):
    """bin/ppp2d.pl in process: every depth prefix of gene `gi`'s hit list is one query, all of them are scored against
    every gene of the genome in ONE GPU call (the reference forks `perl ppp.pl` once per depth, bin/ppp2d.pl:317-320),
    the best non-self locus per depth is kept unless its printed score is 0.000, rows are sorted by printed score.
    Returns the stdout text of ppp2d.pl (bin/ppp2d.pl:354-364)."""
    qs = ppp2d_prefix_queries(gi, bla_path, taxdis, start, end, skip, total_taxa)
    header = "#Subject_best_depth\tLocus\t#Yes\t#Total\tDepth\tProb\tScore\tDesc\n"
    if not qs:
        return header
    loci = list(genome_targets)
    lists = [read_bla(t) if isinstance(t, (str, os.PathLike)) else list(t) for t in genome_targets.values()]
    universe = list(qs[-1][2].keys())  # the last prefix profile holds every taxon any prefix uses
    index = packer.universe_index(universe)
    P, Y, _ = packer.pack_queries([q[2] for q in qs], index)
    p = np.array([q[3] for q in qs], dtype=np.float64)
    idx, raw, off = packer.pack_ranked_targets(lists, index)
    ctx.set_targets_ranked(idx, raw, off, len(index))
    k = min(2, len(loci))
    hits = ctx.score(P, Y, p, capi.make_filter(capi.FILTER_TOPK, k=k))
    best = [[(loci[int(h["t"])], float(h["score"]), int(h["yes"]), int(h["number"]), int(h["depth"])) for h in hits[hits["q"] == qi]]
            for qi in range(len(qs))]
    return format_ppp2d(gi, qs, best)


def format_ppp2d(gi: str, qs, best) -> str:
    """stdout of bin/ppp2d.pl (:322-364) from the depth-prefix queries `qs` (ppp2d_prefix_queries) and, per query, its best
    rows [(locus, score, yes, number, depth)] in ppp.pl's order (ascending raw score): the first row that is not the gene
    itself is kept unless its printed score is 0.000; rows are sorted by printed score."""
    header = "#Subject_best_depth\tLocus\t#Yes\t#Total\tDepth\tProb\tScore\tDesc\n"
    rows = []
    for (line_count, tc, _, pq), mine in zip(qs, best):
        mine = [h for h in mine if h[0] != gi]  # bin/ppp2d.pl:336-338
        if not mine:
            continue
        locus, score, yes, number, depth = mine[0]
        score_str = _handle_negative_zero(_log10_3f(_perl_num(score)))
        if float(score_str) == 0.0:  # bin/ppp2d.pl:340 compares the PRINTED score
            continue
        # bin/ppp2d.pl:333-345 splits the ppp.pl row on tabs and joins it again: an empty Desc field (no description
        # database) is a trailing empty field, which split drops
        fields = [locus, str(yes), str(number), str(depth), "%.3f" % _perl_num(pq), score_str]
        rows.append((float(score_str), line_count, fields))
    rows.sort(key=lambda r: (r[0], r[1]))
    return header + "".join(f"{lc}\t" + "\t".join(f) + "\n" for _, lc, f in rows)


# ------------------------------------------------------------------------------------------------ ppp_cross_score.pl
def _m8_rows(text: str):
    """Non-comment, non-empty lines of a BLAST -m 8/9 table as 12-field lists (bin/ppp_cross_score.pl:361-370, 431-442)."""
    for line in text.split("\n"):
        if line.startswith("#") or line in ("", "0"):
            continue
        f = line.split("\t")
        if len(f) != 12:
            raise ValueError("Wrong BLAST result format. You should run BLAST with -m 8/9")
        yield f


def _m8_gi(subject: str) -> str:
    import re

    m = re.search(r"gi\|(\d+)", subject)
    if not m:
        raise ValueError("Could not parse gi from file")
    return m.group(1)


def cross_target_list(m8_text: str, gi2taxon: dict):
    """bin/ppp_cross_score.pl:350-413 (get_target_profile_blast): one list entry per table row whose subject gi has a
    taxon -- repeated subjects (HSPs) are NOT merged (the reference's `exists $data{$f[11]}` tests the bit score as a
    key, so its de-duplication never fires) -- holding the subject's LAST taxon; the scorer then skips repeated taxa but
    counts their positions in the depth (ppp.pm:189-192)."""
    order, taxon = [], {}
    for f in _m8_rows(m8_text):
        t = gi2taxon.get(_m8_gi(f[1]))
        if not t:
            continue
        order.append(f[1])
        taxon[f[1]] = t
    return [taxon[s] for s in order]


def cross_prefix_queries(m8_text: str, gi2taxon: dict, n_taxa: int, start=1, end=None, skip=1, scale=1.0):
    """The depth-prefix queries of bin/ppp_cross_score.pl:415-546 (_evaluate, BLAST mode): at every row that brings a new
    taxon (subject to --skip / --start-depth / --end-depth) the query profile is {seen taxa: 1} (no zeros) and
    p = line_count / n_taxa * scale.  Returns [(line_count, gi, seen-taxa tuple, p)]."""
    out, seen, line_count = [], {}, 0
    for f in _m8_rows(m8_text):
        line_count += 1
        gi = _m8_gi(f[1])
        t = gi2taxon.get(gi)
        if not t or t in seen:
            continue
        seen[t] = 1
        if line_count % skip:
            continue
        tc = len(seen)
        if end and tc > end:
            break
        if tc >= start:
            pr = line_count / n_taxa
            if scale:
                pr = pr * scale
            if not pr > 0:
                raise ValueError("Error in calculating probability")
            out.append((line_count, gi, tuple(seen), pr))
    return out


def _gpu_list_scorer(ctx: capi.Context):
    """queries (list of {taxon: 0|1}), ONE ordered target taxon list, p per query -> [(score, yes, number, depth)] in one GPU call."""

    def score(queries, target, probs):
        universe = {}
        for q in queries:
            for t in q:
                universe.setdefault(t, None)
        index = packer.universe_index(list(universe))
        P, Y, _ = packer.pack_queries(queries, index)
        idx, raw, off = packer.pack_ranked_targets([target], index)
        ctx.set_targets_ranked(idx, raw, off, len(index))
        hits = ctx.score(P, Y, np.asarray(probs, dtype=np.float64))
        return [(float(h["score"]), int(h["yes"]), int(h["number"]), int(h["depth"])) for h in hits]

    return score


def _cross_direction(scorer, src_text, dst_text, gi2taxon, n_taxa, start, end, skip, scale):
    """Every depth prefix of `src` scored against the hit list of `dst` in ONE scorer call; the best depth is the first
    one whose printed %.3f log-score is strictly lower than the running best (bin/ppp_cross_score.pl:511-538)."""
    qs = cross_prefix_queries(src_text, gi2taxon, n_taxa, start, end, skip, scale)
    if not qs:
        return None
    target = cross_target_list(dst_text, gi2taxon)
    res = scorer([{t: 1 for t in q[2]} for q in qs], target, [q[3] for q in qs])
    best = None
    for (score, yes, number, depth), (line_count, gi, _, pr) in zip(res, qs):
        if score <= 0.0:
            raise ZeroDivisionError("Can't take log of 0")  # bin/ppp_cross_score.pl:511 dies on log(0)
        row = (_log10_3f(score), yes, number, depth, "%.2f" % pr)
        if best is None or float(best[2][0]) > float(row[0]):
            best = (line_count, gi, row)
    return best


def format_cross_score(score_lists, file1: str, text1: str, file2: str, text2: str, gi2taxon: dict, n_taxa: int, start=1, end=None,
                       skip=1, scale=1.0) -> str:
    """Host logic of bin/ppp_cross_score.pl --method1 blast --method2 blast (:283-292, 415-546) around a list scorer
    `score_lists(queries, target, probs) -> [(score, yes, number, depth)]`: both directions, depth-prefix queries, the
    %.3f ranking and the stdout text.  cross_score_blast() below binds it to the GPU."""
    out = []
    for (which, fa, ta, fb, tb, total) in (("File1", file1, text1, file2, text2, "#Total"), ("File2", file2, text2, file1, text1, "#Toal")):
        best = _cross_direction(score_lists, ta, tb, gi2taxon, n_taxa, start, end, skip, scale)
        if best is None:
            raise ValueError(f"no depth of {fa} was evaluated")
        out.append(f"{which}: {fa} Best result in {fb}:\nSource_Line:gi\tScore\t#yes\t{total}\t#Depth\t#Prob\n"
                   f"{best[0]}:{best[1]}\t" + "\t".join(str(x) for x in best[2]) + "\n")
    return out[0] + "\n" + out[1]


def cross_score_blast(ctx, file1: str, text1: str, file2: str, text2: str, gi2taxon: dict, n_taxa: int, start=1, end=None,
                      skip=1, scale=1.0) -> str:
    """bin/ppp_cross_score.pl --method1 blast --method2 blast in process: both directions, each one GPU call over all
    depth prefixes (the reference writes a temporary profile file and runs the Perl sweep per depth).  Returns the
    script's stdout text (bin/ppp_cross_score.pl:283-292).  n_taxa = non-comment lines of --dist-file (:299-311)."""
    return format_cross_score(_gpu_list_scorer(ctx), file1, text1, file2, text2, gi2taxon, n_taxa, start, end, skip, scale)


# ------------------------------------------------------------------------------------------------ profile families
def _taxon_sort_key(t: str):
    """ascending numeric taxon id (bin/hmm3profile.pl:175-181 prints `sort { $a <=> $b }`); non-numeric names after, by name"""
    try:
        return (0, float(t), t)
    except ValueError:
        return (1, 0.0, t)


def read_profile_family(paths, header=False, native=True):
    """A set of profile files as the profile builders write them (bin/hmm3profile.pl:175-181, bin/blast2profile.pl,
    format of lib/PhyloProf/Profile.pm:433-514: `taxon<TAB>0|1`, '#' comments) -> one shared universe (the union of
    their taxa in canonical ascending-taxon-id order, as data/taxdis.txt lists it) and the packed planes:
      P, Y, p   every profile as a QUERY (present = taxa listed in the file, yes = value != 0, p = yes / total)
      T         every profile as a TARGET: its yes taxa as a bit plane, i.e. the list [g : Y[g] = 1] in canonical order
                (the embedding of SURVEY.md section 8d)
    native=True: the multi-threaded C++ reader and packer (ppp_read_profiles, csrc/ppp_pack.cpp); native=False: the Python
    restatement of the same rules (kept as its cross-check).  Returns (universe list, P, Y, p, T)."""
    if native:
        universe, P, Y, p, _, _ = capi.read_profile_files([str(x) for x in paths], header=header)
        return universe, P, Y, p, Y.copy()
    profs = [Profile(file=str(p), header=header).get_hash() for p in paths]
    taxa = set()
    for q in profs:
        taxa.update(q.keys())
    universe = sorted(taxa, key=_taxon_sort_key)
    index = packer.universe_index(universe)
    P, Y, p = packer.pack_queries(profs, index)
    return universe, P, Y, p, Y.copy()


def profile_family_scan(ctx: capi.Context, paths, k=None, thresh_log10=None, header=False):
    """All-pairs PPP scan of a family of profile files (BASELINE config 5: every profile against every profile's yes set)
    in one GPU call; `k` = per-query top-k, `thresh_log10` = keep log10(score) < thresh (e.g. -0.0005: every row whose
    printed %.3f log-score is negative, which is all bin/ppp_cutoff.pl ever reads), neither = every pair.
    Returns (universe, hits)."""
    universe, P, Y, p, T = read_profile_family(paths, header)
    ctx.set_targets_bits(T, len(universe))
    if k:
        flt = capi.make_filter(capi.FILTER_TOPK, k=min(int(k), len(paths)))
    elif thresh_log10 is not None:
        flt = capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=float(thresh_log10))
    else:
        flt = None
    return universe, ctx.score(P, Y, p, flt)
