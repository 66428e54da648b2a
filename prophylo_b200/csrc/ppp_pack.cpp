/*
 * prophylo_b200/csrc/ppp_pack.cpp -- native host-side packers either side of the scan:
 *
 *  ppp_read_profiles      a family of profile files (2 tab columns taxon, 0/1; '#' comments; optional header line -- the
 *      format of /root/reference/lib/PhyloProf/Profile.pm:433-514, written by bin/hmm3profile.pl:175-181, bin/blast2profile.pl)
 *      -> one shared universe (the union of the files' taxa in ascending numeric taxon order, as data/taxdis.txt and
 *      hmm3profile.pl's `sort { $a <=> $b }` list it) and the packed presence / yes bit planes of every profile, uint32 words
 *      in genome order, rows of ppp_words_per_row(G) words -- the layout ppp_set_queries / ppp_set_targets_bits take.
 *      Files are parsed by a pool of host threads; the planes are packed word-wise, not bit by bit through a hash per taxon.
 *
 *  ppp_read_bla_cached    ppp_read_bla_flags with a persistent packed cache: the ranked lists of a genome's `.bla` files are
 *      stored in one binary file keyed by the flags, the universe and every path's size and modification time; a later call
 *      with the same inputs maps the arrays back without opening a single `.bla` (the reference re-parses thousands of
 *      small files on every run, bin/ppp.pl:606-673, 955-1017).
 */
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ppp_b200.h"

namespace {

/* whole file with one open / fstat / read */
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

/* Perl's numeric value of a string: its leading number, 0 when there is none (`$hash{$taxon} != 0`, Profile.pm:463, 477) */
double perl_num(const char *s, size_t len)
{
    char tmp[64];
    const size_t n = std::min(len, sizeof(tmp) - 1);
    memcpy(tmp, s, n);
    tmp[n] = 0;
    char *end = nullptr;
    const double v = strtod(tmp, &end);
    return end == tmp ? 0.0 : v;
}

/* does the whole taxon id read as a number?  (order of the universe: numeric ids ascending, other names after them by name) */
bool whole_number(const std::string &s, double *v)
{
    if (s.empty()) return false;
    char *end = nullptr;
    errno = 0;
    *v = strtod(s.c_str(), &end);
    return end == s.c_str() + s.size() && !(*v != *v);
}

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

/* One parsed file: its bytes stay alive and the items point into them (no string per taxon). */
struct Item {
    uint32_t off, len; /* taxon = buf[off .. off + len) */
    bool yes;
};
struct ParsedProfile {
    std::vector<char> buf;
    std::vector<Item> items; /* (taxon, yes) after the file's own duplicate rule, file order */
    std::string err;
    const char *taxon(size_t k) const { return buf.data() + items[k].off; }
};

void parse_profile(const char *path, bool header, std::vector<uint32_t> &tab, ParsedProfile &out)
{
    if (!slurp(path, out.buf)) {
        out.err = std::string("Can' open ") + path; /* the reference's own spelling (Profile.pm:436) */
        return;
    }
    const char *const base = out.buf.data(), *p = base, *const end = base + out.buf.size();
    size_t lines = 1;
    for (const char *q = p; (q = (const char *)memchr(q, '\n', (size_t)(end - q))) != nullptr; ++q) ++lines;
    tab.assign(pow2_at_least(2 * lines), 0u); /* taxon -> item index + 1 */
    const size_t mask = tab.size() - 1;
    out.items.reserve(lines);
    size_t lineno = 0;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char *line = p;
        p = eol + 1;
        ++lineno;
        if (header && lineno == 1) continue;     /* physical line 1, whatever it holds (Profile.pm:447-449) */
        if (line < eol && line[0] == '#') continue; /* :450 */
        /* split(/\t/): trailing empty fields are dropped; exactly two fields must remain (:452-456), i.e. a tab, a non-empty
         * second field, and nothing but tabs after it */
        const char *t0 = (const char *)memchr(line, '\t', (size_t)(eol - line));
        bool two = false;
        const char *f1 = nullptr, *t1 = nullptr;
        if (t0) {
            f1 = t0 + 1;
            t1 = (const char *)memchr(f1, '\t', (size_t)(eol - f1));
            if (!t1) t1 = eol;
            two = t1 > f1;
            for (const char *c = t1; two && c < eol; ++c) two = (*c == '\t');
        }
        if (!two) {
            out.err = std::string(path) + " must contain only two fields";
            return;
        }
        const size_t l0 = (size_t)(t0 - line), l1 = (size_t)(t1 - f1);
        if (l0 == 0 || (l0 == 1 && line[0] == '0')) continue; /* `next unless $taxon` (:460) */
        const bool yes = (l1 == 1 && f1[0] >= '0' && f1[0] <= '9') ? f1[0] != '0' : perl_num(f1, l1) != 0.0;
        size_t i = hash_bytes(line, l0) & mask;
        while (tab[i] && !same(out.taxon(tab[i] - 1), out.items[tab[i] - 1].len, line, l0)) i = (i + 1) & mask;
        if (!tab[i]) {
            out.items.push_back({(uint32_t)(line - base), (uint32_t)l0, yes});
            tab[i] = (uint32_t)out.items.size();
        } else if (!out.items[tab[i] - 1].yes) {
            out.items[tab[i] - 1].yes = yes; /* a non-zero value is never overwritten, a zero one is (:463-468) */
        }
    }
}

template <typename F>
void parallel_for(size_t n, int nthreads, F fn)
{
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
    };
    size_t nt = nthreads > 0 ? (size_t)nthreads : std::max(1u, std::thread::hardware_concurrency());
    if (nt > n) nt = n ? n : 1;
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (std::thread &t : pool) t.join();
}

int set_err(char *err, size_t errlen, int code, const std::string &msg)
{
    if (err && errlen) snprintf(err, errlen, "%s", msg.c_str());
    return code;
}

} // namespace

extern "C" void ppp_free_profile_set(ppp_profile_set *s)
{
    if (!s) return;
    free(s->universe_buf);
    free(s->universe_off);
    free(s->P);
    free(s->Y);
    free(s->p);
    free(s->yes);
    free(s->total);
    memset(s, 0, sizeof(*s));
}

extern "C" int ppp_read_profiles(const char *const *paths, size_t n_paths, int header, int nthreads, ppp_profile_set *out, char *err,
                                 size_t errlen)
{
    if (err && errlen) err[0] = 0;
    if (!out || (n_paths && !paths)) return set_err(err, errlen, PPP_ERR_INVALID, "ppp_read_profiles: invalid argument");
    memset(out, 0, sizeof(*out));
    std::vector<ParsedProfile> per(n_paths);
    {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            std::vector<uint32_t> tab; /* per-thread scratch of the duplicate rule */
            for (size_t i; (i = next.fetch_add(1)) < n_paths;) parse_profile(paths[i], header != 0, tab, per[i]);
        };
        size_t nt = nthreads > 0 ? (size_t)nthreads : std::max(1u, std::thread::hardware_concurrency());
        if (nt > n_paths) nt = n_paths ? n_paths : 1;
        std::vector<std::thread> pool;
        for (size_t t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
    }
    for (size_t i = 0; i < n_paths; ++i)
        if (!per[i].err.empty()) return set_err(err, errlen, PPP_ERR_INVALID, per[i].err);
    /* The files of a family usually list the SAME taxa in the same order (hmm3profile.pl writes every family over one genome
     * list): such files share file 0's mapping and never touch the union table. */
    std::vector<unsigned char> like0(n_paths, 0);
    parallel_for(n_paths, nthreads, [&](size_t i) {
        const ParsedProfile &a = per[0], &b = per[i];
        if (a.items.size() != b.items.size()) return;
        for (size_t k = 0; k < a.items.size(); ++k)
            if (!same(a.taxon(k), a.items[k].len, b.taxon(k), b.items[k].len)) return;
        like0[i] = 1;
    });
    /* the shared universe: union of the taxa, numeric ids ascending, other names after them by name */
    struct Taxon {
        const char *s;
        uint32_t len, g;
    };
    std::vector<Taxon> taxa;
    std::vector<uint32_t> utab(1024, 0u); /* taxon -> index into taxa + 1 */
    auto intern = [&](const char *t, uint32_t len) -> uint32_t {
        if (2 * (taxa.size() + 1) > utab.size()) { /* grow and re-insert */
            utab.assign(2 * utab.size(), 0u);
            for (size_t j = 0; j < taxa.size(); ++j) {
                size_t i = hash_bytes(taxa[j].s, taxa[j].len) & (utab.size() - 1);
                while (utab[i]) i = (i + 1) & (utab.size() - 1);
                utab[i] = (uint32_t)j + 1;
            }
        }
        size_t i = hash_bytes(t, len) & (utab.size() - 1);
        while (utab[i] && !same(taxa[utab[i] - 1].s, taxa[utab[i] - 1].len, t, len)) i = (i + 1) & (utab.size() - 1);
        if (!utab[i]) {
            taxa.push_back({t, len, 0u});
            utab[i] = (uint32_t)taxa.size();
        }
        return utab[i] - 1;
    };
    for (size_t i = 0; i < n_paths; ++i)
        if (i == 0 || !like0[i])
            for (size_t k = 0; k < per[i].items.size(); ++k) intern(per[i].taxon(k), per[i].items[k].len);
    struct Key {
        bool numeric;
        double v;
        std::string s;
        uint32_t id;
    };
    std::vector<Key> keys(taxa.size());
    for (size_t i = 0; i < taxa.size(); ++i) {
        keys[i].s.assign(taxa[i].s, taxa[i].len);
        keys[i].id = (uint32_t)i;
        double v = 0.0;
        keys[i].numeric = whole_number(keys[i].s, &v);
        keys[i].v = keys[i].numeric ? v : 0.0;
    }
    std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) {
        if (a.numeric != b.numeric) return a.numeric;
        if (a.numeric && a.v != b.v) return a.v < b.v;
        return a.s < b.s;
    });
    const size_t G = keys.size();
    if (G > 8192) return set_err(err, errlen, PPP_ERR_INVALID, "ppp_read_profiles: the profiles span " + std::to_string(G) + " taxa; at most 8192 are supported");
    for (size_t g = 0; g < G; ++g) taxa[keys[g].id].g = (uint32_t)g;
    std::vector<uint32_t> map0; /* genome index of file 0's k-th item */
    if (n_paths) {
        map0.resize(per[0].items.size());
        for (size_t k = 0; k < map0.size(); ++k) map0[k] = taxa[intern(per[0].taxon(k), per[0].items[k].len)].g;
    }
    const size_t wpr = G ? ppp_words_per_row((uint32_t)G) : 4;
    size_t ubytes = 0;
    for (const Key &k : keys) ubytes += k.s.size() + 1;
    out->universe_buf = (char *)malloc(ubytes ? ubytes : 1);
    out->universe_off = (size_t *)malloc((G ? G : 1) * sizeof(size_t));
    out->P = (uint32_t *)calloc(n_paths ? n_paths * wpr : 1, sizeof(uint32_t));
    out->Y = (uint32_t *)calloc(n_paths ? n_paths * wpr : 1, sizeof(uint32_t));
    out->p = (double *)malloc((n_paths ? n_paths : 1) * sizeof(double));
    out->yes = (uint32_t *)malloc((n_paths ? n_paths : 1) * sizeof(uint32_t));
    out->total = (uint32_t *)malloc((n_paths ? n_paths : 1) * sizeof(uint32_t));
    if (!out->universe_buf || !out->universe_off || !out->P || !out->Y || !out->p || !out->yes || !out->total) {
        ppp_free_profile_set(out);
        return set_err(err, errlen, PPP_ERR_NOMEM, "ppp_read_profiles: out of memory");
    }
    size_t o = 0;
    for (size_t g = 0; g < G; ++g) {
        out->universe_off[g] = o;
        memcpy(out->universe_buf + o, keys[g].s.c_str(), keys[g].s.size() + 1);
        o += keys[g].s.size() + 1;
    }
    parallel_for(n_paths, nthreads, [&](size_t i) {
        uint32_t *P = out->P + i * wpr, *Y = out->Y + i * wpr;
        uint32_t ny = 0;
        const ParsedProfile &pp = per[i];
        for (size_t k = 0; k < pp.items.size(); ++k) {
            uint32_t g;
            if (like0[i]) {
                g = map0[k];
            } else { /* read-only probe of the union table */
                size_t j = hash_bytes(pp.taxon(k), pp.items[k].len) & (utab.size() - 1);
                while (!same(taxa[utab[j] - 1].s, taxa[utab[j] - 1].len, pp.taxon(k), pp.items[k].len)) j = (j + 1) & (utab.size() - 1);
                g = taxa[utab[j] - 1].g;
            }
            P[g >> 5] |= 1u << (g & 31);
            if (pp.items[k].yes) {
                Y[g >> 5] |= 1u << (g & 31);
                ++ny;
            }
        }
        out->yes[i] = ny;
        out->total[i] = (uint32_t)per[i].items.size();
        out->p[i] = per[i].items.empty() ? 0.0 / 0.0 : (double)ny / (double)per[i].items.size(); /* ppp.pm:138-142 */
    });
    out->n = n_paths;
    out->G = G;
    out->wpr = wpr;
    return PPP_OK;
}

/* ---------------------------------------------------------------- persistent packed cache of a genome's `.bla` lists */
namespace {

const char CACHE_MAGIC[8] = {'P', 'P', 'P', 'B', 'L', 'A', '0', '2'};

struct CacheHeader {
    char magic[8];
    uint64_t key;
    uint64_t n_targets, n_entries;
};

uint64_t fnv(uint64_t h, const void *data, size_t n)
{
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

/* key of a cache: flags, universe, and every path with its size and modification time; false if a file cannot be stat'ed */
bool cache_key(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, uint32_t flags, uint64_t *key)
{
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, &flags, sizeof(flags));
    h = fnv(h, &G, sizeof(G));
    for (size_t g = 0; g < G; ++g) h = fnv(h, universe[g], strlen(universe[g]) + 1);
    for (size_t i = 0; i < n_paths; ++i) {
        struct stat st;
        if (stat(paths[i], &st) != 0) return false;
        h = fnv(h, paths[i], strlen(paths[i]) + 1);
        const int64_t meta[3] = {(int64_t)st.st_size, (int64_t)st.st_mtim.tv_sec, (int64_t)st.st_mtim.tv_nsec};
        h = fnv(h, meta, sizeof(meta));
    }
    *key = h;
    return true;
}

} // namespace

extern "C" int ppp_read_bla_cached(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, int nthreads,
                                   uint32_t flags, const char *cache_path, ppp_ranked_lists *out, int *from_cache, char *err, size_t errlen)
{
    if (from_cache) *from_cache = 0;
    if (!cache_path || !cache_path[0]) return ppp_read_bla_flags(paths, n_paths, universe, G, nthreads, flags, out, err, errlen);
    if (err && errlen) err[0] = 0;
    if (!out || (n_paths && !paths) || (G && !universe)) return set_err(err, errlen, PPP_ERR_INVALID, "ppp_read_bla_cached: invalid argument");
    uint64_t key = 0;
    const bool have_key = cache_key(paths, n_paths, universe, G, flags, &key);
    if (have_key) {
        FILE *f = fopen(cache_path, "rb");
        if (f) {
            CacheHeader h;
            bool ok = fread(&h, sizeof(h), 1, f) == 1 && !memcmp(h.magic, CACHE_MAGIC, 8) && h.key == key && h.n_targets == n_paths;
            if (ok) {
                memset(out, 0, sizeof(*out));
                out->idx = (uint16_t *)malloc((h.n_entries ? h.n_entries : 1) * 2);
                out->rawdepth = (uint16_t *)malloc((h.n_entries ? h.n_entries : 1) * 2);
                out->row_off = (uint64_t *)malloc((n_paths + 1) * 8);
                ok = out->idx && out->rawdepth && out->row_off && fread(out->row_off, 8, n_paths + 1, f) == n_paths + 1 &&
                     fread(out->idx, 2, h.n_entries, f) == h.n_entries && fread(out->rawdepth, 2, h.n_entries, f) == h.n_entries &&
                     out->row_off[n_paths] == h.n_entries;
                if (ok) {
                    out->n_targets = n_paths;
                    out->n_entries = h.n_entries;
                    fclose(f);
                    if (from_cache) *from_cache = 1;
                    return PPP_OK;
                }
                ppp_free_ranked_lists(out); /* truncated or foreign file: fall through and rebuild it */
            }
            fclose(f);
        }
    }
    const int rc = ppp_read_bla_flags(paths, n_paths, universe, G, nthreads, flags, out, err, errlen);
    if (rc != PPP_OK || !have_key) return rc;
    /* write the cache next to its final name and rename it into place (readers never see a partial file); a cache that
     * cannot be written is not an error */
    const std::string tmp = std::string(cache_path) + ".tmp." + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (f) {
        CacheHeader h;
        memcpy(h.magic, CACHE_MAGIC, 8);
        h.key = key;
        h.n_targets = out->n_targets;
        h.n_entries = out->n_entries;
        const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(out->row_off, 8, n_paths + 1, f) == n_paths + 1 &&
                        fwrite(out->idx, 2, out->n_entries, f) == out->n_entries &&
                        fwrite(out->rawdepth, 2, out->n_entries, f) == out->n_entries;
        const bool closed = fclose(f) == 0;
        if (!(ok && closed && rename(tmp.c_str(), cache_path) == 0)) unlink(tmp.c_str());
    }
    return PPP_OK;
}
