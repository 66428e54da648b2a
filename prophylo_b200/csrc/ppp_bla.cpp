/*
 * prophylo_b200/csrc/ppp_bla.cpp -- native host-side reader of the reference's per-gene BLAST caches (`.bla`).
 *
 * Format (written by /root/reference/util/cache_blast_result.pl:86 and bin/ppp.pl:1143): one hit per line, four tab
 * separated columns  query_gi, subject_gi, e-value, subject_taxon.  What the reference builds from one file
 * (bin/ppp.pl:606-673 get_profile + lib/PhyloProf/BlastResult.pm:192-237 get_profile) and what this reader reproduces:
 *   - subjects keep their FIRST-occurrence order, a subject's LAST taxon wins (%subject_taxon_class is overwritten);
 *   - subjects with a false taxon (missing column, "" or "0") are dropped (BlastResult.pm:206-209);
 *   - different subjects of the same taxon stay as duplicates: they are skipped by the scorer (ppp.pm:189-192) but
 *     count in the depth (ppp.pm:169,231), so every kept entry carries its raw 1-based list position.
 * Output = the arrays ppp_set_targets_ranked() takes: for every file the de-duplicated taxa mapped to genome indices of
 * the caller's universe (0xFFFF = foreign taxon) with their raw positions.  Files are parsed by a pool of host threads
 * (the reference parses them in forked `perl ppp.pl --client` workers, 50 files per worker: bin/ppp.pl:955-1017).
 * Lines are split on '\n' only and fields on '\t' only, as Perl's chomp/split do.
 */
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ppp_b200.h"

namespace {

struct FileLists {
    std::vector<uint16_t> idx, raw;
    std::string err;
};

/* Strings are hashed 8 bytes at a time (taxon ids and gi numbers are 4 - 10 characters: one or two rounds). */
inline uint64_t hash_bytes(const char *p, size_t n)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    while (n >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        h = (h ^ v) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
        p += 8;
        n -= 8;
    }
    if (n) {
        uint64_t v = 0;
        memcpy(&v, p, n);
        h = (h ^ v) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    return h;
}
inline bool same(const char *a, size_t la, const char *b, size_t lb) { return la == lb && (la == 0 || memcmp(a, b, la) == 0); }
inline size_t pow2_at_least(size_t n)
{
    size_t c = 16;
    while (c < n) c <<= 1;
    return c;
}

/* taxon string -> genome index: open addressing, built once per call, read by every parser thread */
struct Universe {
    struct Slot {
        const char *s = nullptr;
        uint32_t len = 0;
        uint16_t g = 0;
    };
    std::vector<Slot> tab;
    size_t mask = 0;
    void build(const char *const *names, size_t G)
    {
        tab.assign(pow2_at_least(2 * G + 1), Slot());
        mask = tab.size() - 1;
        for (size_t g = 0; g < G; ++g) {
            const char *s = names[g];
            const size_t len = strlen(s);
            size_t i = hash_bytes(s, len) & mask;
            while (tab[i].s && !same(tab[i].s, tab[i].len, s, len)) i = (i + 1) & mask;
            if (!tab[i].s) { /* first occurrence wins */
                tab[i].s = s;
                tab[i].len = (uint32_t)len;
                tab[i].g = (uint16_t)g;
            }
        }
    }
    uint16_t find(const char *s, size_t len) const
    {
        for (size_t i = hash_bytes(s, len) & mask; tab[i].s; i = (i + 1) & mask)
            if (same(tab[i].s, tab[i].len, s, len)) return tab[i].g;
        return 0xFFFF;
    }
};

/* Per-thread scratch, reused from file to file: the file's bytes, its subjects in first-occurrence order, and two open-addressing
 * tables over them (subject -> entry; taxon -> seen).  The first version of this reader kept std::unordered_map / _set of
 * string_views per file -- a heap node per subject -- and parsed 94 MB/s per thread; this one allocates nothing per file. */
struct Scratch {
    struct Subject {
        const char *key, *tax;
        uint32_t klen, tlen;
    };
    std::vector<char> buf;
    std::vector<Subject> subj;
    std::vector<uint32_t> stab, ttab; /* entry index + 1, 0 = empty */
};

/* whole file into s.buf with one open / fstat / read (the caches are a few KB to a few hundred KB each: mapping them would
 * cost a page fault per 4 KB where read() copies them in one call) */
bool slurp(const char *path, std::vector<char> &buf)
{
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    struct stat st;
    size_t cap = (fstat(fd, &st) == 0 && st.st_size > 0) ? (size_t)st.st_size + 1 : (size_t)1 << 16, n = 0;
    buf.resize(cap);
    for (;;) {
        if (n == buf.size()) buf.resize(2 * buf.size());
        const ssize_t r = read(fd, buf.data() + n, buf.size() - n);
        if (r < 0) {
            if (errno == EINTR) continue;
            close(fd);
            return false;
        }
        if (r == 0) break;
        n += (size_t)r;
    }
    close(fd);
    buf.resize(n);
    return true;
}

void parse_one(const char *path, const Universe &universe, bool dup, Scratch &s, FileLists &out)
{
    if (!slurp(path, s.buf)) {
        out.err = std::string("Can't open ") + path;
        return;
    }
    const char *p = s.buf.data(), *const end = p + s.buf.size();
    size_t lines = 1;
    for (const char *q = p; (q = (const char *)memchr(q, '\n', (size_t)(end - q))) != nullptr; ++q) ++lines;
    s.subj.clear();
    s.stab.assign(pow2_at_least(2 * lines), 0u);
    const size_t smask = s.stab.size() - 1;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        /* fields 1 (subject) and 3 (its taxon) of the tab-separated line; missing fields are empty (Perl: undef -> false) */
        const char *f1 = eol, *f3 = eol;
        size_t l1 = 0, l3 = 0;
        const char *c = p;
        for (int i = 0; i < 4 && c <= eol; ++i) {
            const char *tab = (const char *)memchr(c, '\t', (size_t)(eol - c));
            if (!tab) tab = eol;
            if (i == 1) { f1 = c; l1 = (size_t)(tab - c); }
            if (i == 3) { f3 = c; l3 = (size_t)(tab - c); }
            c = tab + 1;
        }
        p = eol + 1;
        size_t i = hash_bytes(f1, l1) & smask;
        while (s.stab[i] && !same(s.subj[s.stab[i] - 1].key, s.subj[s.stab[i] - 1].klen, f1, l1)) i = (i + 1) & smask;
        if (!s.stab[i]) { /* subjects keep their first-occurrence order ... */
            s.subj.push_back({f1, f3, (uint32_t)l1, (uint32_t)l3});
            s.stab[i] = (uint32_t)s.subj.size();
        } else { /* ... and a subject's LAST taxon wins */
            s.subj[s.stab[i] - 1].tax = f3;
            s.subj[s.stab[i] - 1].tlen = (uint32_t)l3;
        }
    }
    s.ttab.assign(pow2_at_least(2 * s.subj.size() + 1), 0u);
    const size_t tmask = s.ttab.size() - 1;
    size_t position = 0, repeats = 0;
    for (size_t e = 0; e < s.subj.size(); ++e) {
        const char *t = s.subj[e].tax;
        const size_t tl = s.subj[e].tlen;
        if (tl == 0 || (tl == 1 && t[0] == '0')) continue; /* Perl-false taxon: the hit is not part of the list */
        ++position;
        if (position > 65535) {
            out.err = std::string(path) + ": target list longer than 65535 positions";
            return;
        }
        size_t i = hash_bytes(t, tl) & tmask;
        while (s.ttab[i] && !same(s.subj[s.ttab[i] - 1].tax, s.subj[s.ttab[i] - 1].tlen, t, tl)) i = (i + 1) & tmask;
        if (s.ttab[i]) {
            /* -dup (ppp.pm:189-200): `$seen{key}` is one counter for the whole list -- the first repeated entry is counted
             * again like a new taxon, from the second on `yes` is frozen and no later step can improve: the list ends */
            if (!dup) continue;
            if (++repeats >= 2) break;
        } else {
            s.ttab[i] = (uint32_t)e + 1;
        }
        out.idx.push_back(universe.find(t, tl));
        out.raw.push_back((uint16_t)position);
    }
}

} // namespace

extern "C" int ppp_read_bla(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, int nthreads,
                            ppp_ranked_lists *out, char *err, size_t errlen)
{
    return ppp_read_bla_flags(paths, n_paths, universe, G, nthreads, 0u, out, err, errlen);
}

extern "C" int ppp_read_bla_flags(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, int nthreads,
                                  uint32_t flags, ppp_ranked_lists *out, char *err, size_t errlen)
{
    const bool dup = (flags & PPP_BLA_DUP) != 0;
    if (err && errlen) err[0] = 0;
    if (!out || (n_paths && !paths) || (G && !universe) || G > 0xFFFF) {
        if (err && errlen) snprintf(err, errlen, "ppp_read_bla: invalid argument");
        return PPP_ERR_INVALID;
    }
    memset(out, 0, sizeof(*out));
    Universe uni;
    uni.build(universe, G); /* first occurrence wins */
    std::vector<FileLists> per(n_paths);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Scratch scratch;
        for (size_t i; (i = next.fetch_add(1)) < n_paths;) parse_one(paths[i], uni, dup, scratch, per[i]);
    };
    size_t nt = nthreads > 0 ? (size_t)nthreads : std::max(1u, std::thread::hardware_concurrency());
    if (nt > n_paths) nt = n_paths ? n_paths : 1;
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
    size_t total = 0;
    for (size_t i = 0; i < n_paths; ++i) {
        if (!per[i].err.empty()) {
            if (err && errlen) snprintf(err, errlen, "%s", per[i].err.c_str());
            return PPP_ERR_INVALID;
        }
        total += per[i].idx.size();
    }
    out->idx = (uint16_t *)malloc((total ? total : 1) * sizeof(uint16_t));
    out->rawdepth = (uint16_t *)malloc((total ? total : 1) * sizeof(uint16_t));
    out->row_off = (uint64_t *)malloc((n_paths + 1) * sizeof(uint64_t));
    if (!out->idx || !out->rawdepth || !out->row_off) {
        ppp_free_ranked_lists(out);
        if (err && errlen) snprintf(err, errlen, "ppp_read_bla: out of memory");
        return PPP_ERR_NOMEM;
    }
    size_t o = 0;
    for (size_t i = 0; i < n_paths; ++i) {
        out->row_off[i] = o;
        if (!per[i].idx.empty()) {
            memcpy(out->idx + o, per[i].idx.data(), per[i].idx.size() * sizeof(uint16_t));
            memcpy(out->rawdepth + o, per[i].raw.data(), per[i].raw.size() * sizeof(uint16_t));
        }
        o += per[i].idx.size();
    }
    out->row_off[n_paths] = o;
    out->n_targets = n_paths;
    out->n_entries = total;
    return PPP_OK;
}

extern "C" void ppp_free_ranked_lists(ppp_ranked_lists *l)
{
    if (!l) return;
    free(l->idx);
    free(l->rawdepth);
    free(l->row_off);
    memset(l, 0, sizeof(*l));
}

/* ---------------------------------------------------------------- host popcounts of the query planes (ppp_set_queries)
 * |Y| and |P| of every query: Profile::get_yes / get_total (lib/PhyloProf/Profile.pm:342-375).  nvcc's host pass compiles
 * __builtin_popcount without -mpopcnt into a library call (8 ms for 10,000 queries x 4096 genomes); here the POPCNT
 * instruction is used when the CPU has it. */
__attribute__((target("popcnt"))) static void count_rows_hw(const uint32_t *P, const uint32_t *Y, size_t nQ, size_t wpr, uint32_t *ny, uint32_t *np)
{
    for (size_t q = 0; q < nQ; ++q) {
        const uint32_t *pr = P + q * wpr, *yr = Y + q * wpr;
        uint32_t a = 0, b = 0;
        for (size_t w = 0; w < wpr; ++w) {
            a += (uint32_t)__builtin_popcount(pr[w] & yr[w]);
            b += (uint32_t)__builtin_popcount(pr[w]);
        }
        ny[q] = a;
        np[q] = b;
    }
}
static void count_rows_sw(const uint32_t *P, const uint32_t *Y, size_t nQ, size_t wpr, uint32_t *ny, uint32_t *np)
{
    for (size_t q = 0; q < nQ; ++q) {
        const uint32_t *pr = P + q * wpr, *yr = Y + q * wpr;
        uint32_t a = 0, b = 0;
        for (size_t w = 0; w < wpr; ++w) {
            a += (uint32_t)__builtin_popcount(pr[w] & yr[w]);
            b += (uint32_t)__builtin_popcount(pr[w]);
        }
        ny[q] = a;
        np[q] = b;
    }
}
extern "C" void ppp_host_count_rows(const uint32_t *P, const uint32_t *Y, size_t nQ, size_t wpr, uint32_t *ny, uint32_t *np)
{
    if (__builtin_cpu_supports("popcnt")) count_rows_hw(P, Y, nQ, wpr, ny, np);
    else count_rows_sw(P, Y, nQ, wpr, ny, np);
}
