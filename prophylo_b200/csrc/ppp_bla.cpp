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
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/ppp_b200.h"

namespace {

struct FileLists {
    std::vector<uint16_t> idx, raw;
    std::string err;
};

bool slurp(const char *path, std::string &buf)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char tmp[1 << 16];
    size_t n;
    buf.clear();
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, n);
    const bool ok = !ferror(f);
    fclose(f);
    return ok;
}

void parse_one(const char *path, const std::unordered_map<std::string_view, uint16_t> &universe, bool dup, FileLists &out)
{
    std::string buf;
    if (!slurp(path, buf)) {
        out.err = std::string("Can't open ") + path;
        return;
    }
    std::vector<std::string_view> order;                             /* subjects, first-occurrence order */
    std::unordered_map<std::string_view, std::string_view> taxon_of; /* subject -> its last taxon */
    size_t pos = 0;
    const size_t n = buf.size();
    while (pos < n) {
        size_t eol = buf.find('\n', pos);
        if (eol == std::string::npos) eol = n;
        std::string_view line(buf.data() + pos, eol - pos);
        pos = eol + 1;
        std::string_view f[4];
        size_t start = 0;
        for (int i = 0; i < 4; ++i) {
            if (start > line.size()) break;
            size_t tab = line.find('\t', start);
            if (tab == std::string_view::npos) tab = line.size();
            f[i] = line.substr(start, tab - start);
            start = tab + 1;
        }
        auto it = taxon_of.find(f[1]);
        if (it == taxon_of.end()) {
            order.push_back(f[1]);
            taxon_of.emplace(f[1], f[3]);
        } else {
            it->second = f[3];
        }
    }
    std::unordered_set<std::string_view> seen;
    size_t position = 0, repeats = 0;
    for (const std::string_view &s : order) {
        const std::string_view t = taxon_of[s];
        if (t.empty() || t == "0") continue; /* Perl-false taxon: the hit is not part of the list */
        ++position;
        if (position > 65535) {
            out.err = std::string(path) + ": target list longer than 65535 positions";
            return;
        }
        if (!seen.insert(t).second) {
            /* -dup (ppp.pm:189-200): `$seen{key}` is one counter for the whole list -- the first repeated entry is counted
             * again like a new taxon, from the second on `yes` is frozen and no later step can improve: the list ends */
            if (!dup) continue;
            if (++repeats >= 2) break;
        }
        const auto u = universe.find(t);
        out.idx.push_back(u == universe.end() ? (uint16_t)0xFFFF : u->second);
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
    std::unordered_map<std::string_view, uint16_t> uni;
    uni.reserve(G * 2);
    for (size_t g = 0; g < G; ++g) uni.emplace(std::string_view(universe[g]), (uint16_t)g); /* first occurrence wins */
    std::vector<FileLists> per(n_paths);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < n_paths;) parse_one(paths[i], uni, dup, per[i]);
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
