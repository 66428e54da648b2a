/* perl/PPPB200.xs -- XS binding of the C ABI in include/ppp_b200.h for the Perl host (PhyloProf::B200).
 * Replaces the per-pair Perl sweep of lib/PhyloProf/Algorithm/ppp.pm:130-245 by one call into libppp_b200.so.
 * FFI::Platypus / Inline::C are not installed in this image; xsubpp + the perl headers are, hence XS.
 * Buffers are Perl strings of packed little-endian uint32 words (pack "V*" / vec), copied by the library before
 * the call returns, so Perl is free to move its SVs. Errors croak with ppp_last_error(). */
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"
#include "ppp_b200.h"

typedef ppp_ctx *PhyloProf__B200__Ctx;

MODULE = PhyloProf::B200  PACKAGE = PhyloProf::B200

IV
_init(devices)
    AV *devices
  CODE:
    /* one CUDA ordinal = one GPU; several = ONE perl process drives them all (targets sharded by the library) */
    ppp_ctx *ctx = NULL;
    int dev[16];
    const int n = (int)(av_len(devices) + 1);
    if (n < 1 || n > 16) croak("PhyloProf::B200: 1 .. 16 devices");
    for (int i = 0; i < n; ++i) {
        SV **e = av_fetch(devices, i, 0);
        dev[i] = e ? (int)SvIV(*e) : 0;
    }
    int rc = ppp_init(&ctx, dev, n);
    if (rc != PPP_OK) croak("PhyloProf::B200: %s", ppp_last_error(NULL));
    RETVAL = PTR2IV(ctx);
  OUTPUT:
    RETVAL

void
_destroy(h)
    IV h
  CODE:
    ppp_destroy(INT2PTR(ppp_ctx *, h));

int
words_per_row(G)
    unsigned int G
  CODE:
    RETVAL = (int)ppp_words_per_row(G);
  OUTPUT:
    RETVAL

void
_set_targets_bits(h, planes, nT, G)
    IV h
    SV *planes
    UV nT
    unsigned int G
  CODE:
    ppp_ctx *ctx = INT2PTR(ppp_ctx *, h);
    STRLEN len;
    const char *buf = SvPVbyte(planes, len);
    if (len < (STRLEN)nT * ppp_words_per_row(G) * 4) croak("PhyloProf::B200: target buffer too short");
    if (ppp_set_targets_bits(ctx, (const uint32_t *)buf, (size_t)nT, G) != PPP_OK)
        croak("PhyloProf::B200: %s", ppp_last_error(ctx));

void
_set_targets_ranked(h, idx, raw, off, nT, G)
    IV h
    SV *idx
    SV *raw
    SV *off
    UV nT
    unsigned int G
  CODE:
    ppp_ctx *ctx = INT2PTR(ppp_ctx *, h);
    STRLEN li, lr, lo;
    const char *ib = SvPVbyte(idx, li);   /* pack "S*": uint16 genome indices, 0xFFFF = foreign taxon */
    const char *rb = SvPVbyte(raw, lr);   /* pack "S*": 1-based raw list positions                      */
    const char *ob = SvPVbyte(off, lo);   /* pack "Q*": nT + 1 entry offsets                            */
    if (lo < (STRLEN)(nT + 1) * 8 || li != lr) croak("PhyloProf::B200: ranked target buffers inconsistent");
    if (nT && li < (STRLEN)(((const uint64_t *)ob)[nT] * 2)) croak("PhyloProf::B200: ranked target lists shorter than row_off[nT]");
    if (ppp_set_targets_ranked(ctx, (const uint16_t *)ib, (const uint16_t *)rb, (const uint64_t *)ob, (size_t)nT, G) != PPP_OK)
        croak("PhyloProf::B200: %s", ppp_last_error(ctx));

void
_set_option(h, name, value)
    IV h
    const char *name
    IV value
  CODE:
    ppp_ctx *ctx = INT2PTR(ppp_ctx *, h);
    if (ppp_set_option(ctx, name, (int64_t)value) != PPP_OK) croak("PhyloProf::B200: %s", ppp_last_error(ctx));

UV
_set_targets_bla(h, paths, universe, nthreads, flags)
    IV h
    AV *paths
    AV *universe
    int nthreads
    unsigned int flags
  CODE:
    /* a genome's worth of `.bla` caches (bin/ppp.pl:606-673) -> ranked targets, parsed natively by ppp_read_bla */
    ppp_ctx *ctx = INT2PTR(ppp_ctx *, h);
    const size_t np = (size_t)(av_len(paths) + 1), G = (size_t)(av_len(universe) + 1);
    const char **pv = NULL, **uv = NULL;
    Newx(pv, np ? np : 1, const char *);
    Newx(uv, G ? G : 1, const char *);
    SAVEFREEPV(pv); /* released when the XSUB is left, also through a croak of SvPVbyte_nolen */
    SAVEFREEPV(uv);
    for (size_t i = 0; i < np; ++i) {
        SV **e = av_fetch(paths, (SSize_t)i, 0);
        if (!e) croak("PhyloProf::B200: undefined path at index %lu", (unsigned long)i);
        pv[i] = SvPVbyte_nolen(*e);
    }
    for (size_t i = 0; i < G; ++i) {
        SV **e = av_fetch(universe, (SSize_t)i, 0);
        if (!e) croak("PhyloProf::B200: undefined taxon at index %lu", (unsigned long)i);
        uv[i] = SvPVbyte_nolen(*e);
    }
    ppp_ranked_lists l;
    char err[512];
    int rc = ppp_read_bla_flags(pv, np, uv, G, nthreads, flags, &l, err, sizeof(err)); /* flags: PPP_BLA_DUP = -dup packing */
    if (rc != PPP_OK) croak("PhyloProf::B200: %s", err);
    rc = ppp_set_targets_ranked(ctx, l.idx, l.rawdepth, l.row_off, l.n_targets, (uint32_t)G);
    RETVAL = (UV)l.n_entries;
    ppp_free_ranked_lists(&l);
    if (rc != PPP_OK) croak("PhyloProf::B200: %s", ppp_last_error(ctx));
  OUTPUT:
    RETVAL

void
_printed_cutoffs(h, nQ, slope)
    IV h
    UV nQ
    double slope
  PPCODE:
    /* per query of the last threshold run: [first row, rows, marker row or -1, CUTOFF in thousandths or undef, status]
     * (ppp_printed_cutoffs: bin/ppp.pl:473-506 -m marker, bin/ppp_cutoff.pl:130-148 CUTOFF, computed on the device) */
    ppp_ctx *ctx = INT2PTR(ppp_ctx *, h);
    uint64_t *off = NULL;
    int32_t *marker = NULL, *cut = NULL;
    uint8_t *status = NULL;
    size_t ctx_nq = 0;
    ppp_query_count(ctx, &ctx_nq); /* the library writes one entry per RESIDENT query: the buffers must be that long */
    if ((size_t)nQ != ctx_nq) croak("PhyloProf::B200: printed_cutoffs asked for %lu queries, the last run had %lu", (unsigned long)nQ, (unsigned long)ctx_nq);
    Newx(off, nQ + 1, uint64_t);
    Newx(marker, nQ ? nQ : 1, int32_t);
    Newx(cut, nQ ? nQ : 1, int32_t);
    Newx(status, nQ ? nQ : 1, uint8_t);
    int rc = ppp_printed_cutoffs(ctx, slope, off, NULL, marker, cut, status);
    if (rc != PPP_OK) {
        Safefree(off); Safefree(marker); Safefree(cut); Safefree(status);
        croak("PhyloProf::B200: %s", ppp_last_error(ctx));
    }
    EXTEND(SP, (SSize_t)nQ);
    for (size_t q = 0; q < (size_t)nQ; ++q) {
        AV *av = newAV();
        av_push(av, newSVuv((UV)off[q]));
        av_push(av, newSVuv((UV)(off[q + 1] - off[q])));
        av_push(av, newSViv(marker[q]));
        av_push(av, cut[q] == INT32_MIN ? newSV(0) : newSViv(cut[q]));
        av_push(av, newSVuv(status[q]));
        PUSHs(sv_2mortal(newRV_noinc((SV *)av)));
    }
    Safefree(off); Safefree(marker); Safefree(cut); Safefree(status);

void
_score(h, pplanes, yplanes, probs, nQ, mode, k, thresh)
    IV h
    SV *pplanes
    SV *yplanes
    SV *probs
    UV nQ
    unsigned int mode
    unsigned int k
    double thresh
  PPCODE:
    ppp_ctx *ctx = INT2PTR(ppp_ctx *, h);
    STRLEN lp, ly, lq;
    const char *pb = SvPVbyte(pplanes, lp);
    const char *yb = SvPVbyte(yplanes, ly);
    const char *qb = SvPVbyte(probs, lq); /* packed native doubles: pack "d*" */
    if (lq < (STRLEN)nQ * sizeof(double) || lp != ly) croak("PhyloProf::B200: query buffers inconsistent");
    {
        size_t wpr = 0;
        ppp_row_words(ctx, &wpr); /* words per plane row of the resident targets' universe */
        if (lp < (STRLEN)nQ * wpr * 4) croak("PhyloProf::B200: query planes shorter than nQ rows of %lu words", (unsigned long)wpr);
    }
    ppp_filter f;
    f.mode = mode; f.k = k; f.thresh = thresh;
    if (ppp_set_queries(ctx, (const uint32_t *)pb, (const uint32_t *)yb, (const double *)qb, (size_t)nQ) != PPP_OK)
        croak("PhyloProf::B200: %s", ppp_last_error(ctx));
    if (ppp_run(ctx, &f) != PPP_OK) croak("PhyloProf::B200: %s", ppp_last_error(ctx));
    size_t n = 0;
    ppp_result_count(ctx, &n);
    ppp_hit *hits = NULL;
    Newx(hits, n ? n : 1, ppp_hit);
    if (ppp_fetch(ctx, hits, n, &n) != PPP_OK) {
        Safefree(hits);
        croak("PhyloProf::B200: %s", ppp_last_error(ctx));
    }
    /* one array ref [q, t, score, yes, number, depth] per hit */
    EXTEND(SP, (SSize_t)n);
    for (size_t i = 0; i < n; ++i) {
        AV *av = newAV();
        av_push(av, newSVuv(hits[i].q));
        av_push(av, newSVuv(hits[i].t));
        av_push(av, newSVnv(hits[i].score));
        av_push(av, newSVuv(hits[i].yes));
        av_push(av, newSVuv(hits[i].number));
        av_push(av, newSVuv(hits[i].depth));
        PUSHs(sv_2mortal(newRV_noinc((SV *)av)));
    }
    Safefree(hits);

