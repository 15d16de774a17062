// Host ingest: k-mer count text files -> packed table (uint64 k-mers + group-planar uint32 counts).
// Replaces the tf.data text path of the reference: dataloader.dataloader (dataloader.py:6-50),
// dataloader.sparse_dataloader (dataloader.py:52-109), core.tf_one_hot's symbol tables
// (core.py:142-153) and the `wc -l` row count (models/train_bear_net.py:54-55).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "bear_b200.h"
#include "bear_host.h"
#include "bear_rank.h"

static thread_local char g_err[512] = "";

void bear_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* bear_last_error(void) { return g_err; }
extern "C" int bear_version(void) { return 1; }
#ifndef BEAR_BUILD_DIGEST
#define BEAR_BUILD_DIGEST "unknown"
#endif
static const char g_build_digest[] = "BEAR_BUILD_DIGEST:" BEAR_BUILD_DIGEST;
extern "C" const char* bear_build_digest(void) { return g_build_digest + 18; }

extern "C" int bear_alphabet_size(int alphabet) {
    if (alphabet == BEAR_ALPHABET_DNA || alphabet == BEAR_ALPHABET_RNA) return 4;
    if (alphabet == BEAR_ALPHABET_PROT) return 20;
    return BEAR_ERR_ARG;
}

extern "C" int bear_max_lag(int alphabet) {
    if (alphabet == BEAR_ALPHABET_DNA || alphabet == BEAR_ALPHABET_RNA) return 29;
    if (alphabet == BEAR_ALPHABET_PROT) return 12;
    return BEAR_ERR_ARG;
}

// core.py:142-153 symbol order; '[' is the start token (last input column).
static const char* kProt = "ARNDCEQGHILKMFPSTWYV";

static int sym_code(int alphabet, char ch) {
    if (alphabet == BEAR_ALPHABET_PROT) {
        if (ch == '[') return 20;
        const char* p = strchr(kProt, ch);
        return (p && ch) ? int(p - kProt) : 31;   // unknown -> all-zero one-hot row
    }
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return alphabet == BEAR_ALPHABET_DNA ? 3 : -1;
        case 'U': return alphabet == BEAR_ALPHABET_RNA ? 3 : -1;
        case '[': return 4;
        default: return -1;
    }
}

// Encode one k-mer of length lag; returns 0 or a negative status.
// What a DNA / RNA k-mer with a symbol outside the alphabet does to a file load: 0 = error (default), 1 = the row is
// marked with BEAR_INVALID_KMER and its counts are zeroed; the caller drops marked rows (dataloader.KmerTable.from_file).
static int g_invalid_policy = 0;
extern "C" int bear_pack_set_invalid_policy(int policy) {
    if (policy != 0 && policy != 1) { bear_set_error("bear_pack_set_invalid_policy: policy must be 0 (error) or 1 (mark)"); return BEAR_ERR_ARG; }
    const int old = g_invalid_policy;
    g_invalid_policy = policy;
    return old;
}

static int encode_one(const char* s, int lag, int alphabet, uint64_t* out) {
    if (alphabet == BEAR_ALPHABET_PROT) {
        uint64_t v = 0;
        for (int j = 0; j < lag; ++j) v = (v << 5) | uint64_t(sym_code(alphabet, s[j]));
        *out = v;
        return 0;
    }
    {   // common case: a run of '[' followed by letters only -- one table look-up per symbol
        static const struct Lut {
            int8_t t[2][256];
            Lut() {
                memset(t, -1, sizeof(t));
                for (int a = 0; a < 2; ++a) {
                    t[a][int('A')] = 0;
                    t[a][int('C')] = 1;
                    t[a][int('G')] = 2;
                    t[a][int(a == 0 ? 'T' : 'U')] = 3;
                }
            }
        } lut;
        const int8_t* t = lut.t[alphabet == BEAR_ALPHABET_DNA ? 0 : 1];
        int j = 0;
        while (j < lag && s[j] == '[') ++j;
        const uint64_t ns = uint64_t(j);
        uint64_t w = 0;
        int bad = 0;
        for (; j < lag; ++j) {
            const int c = t[uint8_t(s[j])];
            bad |= c;
            w = (w << 2) | uint64_t(c & 3);
        }
        if (bad >= 0) {
            *out = w | (ns << 58);
            return 0;
        }
    }   // otherwise: the general loop below finds and reports the offending symbol
    uint64_t v = 0, nstart = 0;
    bool in_prefix = true;
    for (int j = 0; j < lag; ++j) {
        int c = sym_code(alphabet, s[j]);
        if (c < 0) {
            if (g_invalid_policy == 1) {
                *out = BEAR_INVALID_KMER;
                return 0;
            }
            bear_set_error("symbol '%c' outside the alphabet in k-mer '%.*s'", s[j], lag, s);
            return BEAR_ERR_PARSE;
        }
        if (c == 4) {
            if (!in_prefix) {
                bear_set_error("start symbol '[' after a letter in k-mer '%.*s'", lag, s);
                return BEAR_ERR_PARSE;
            }
            ++nstart;
            v <<= 2;
        } else {
            in_prefix = false;
            v = (v << 2) | uint64_t(c);
        }
    }
    *out = v | (nstart << 58);
    return 0;
}

extern "C" int bear_encode_kmers(const char* h_text, int64_t n, int lag, int alphabet, uint64_t* h_kmers) {
    int ml = bear_max_lag(alphabet);
    if (ml < 0 || lag < 1 || !h_text || !h_kmers || n < 0) { bear_set_error("bear_encode_kmers: bad argument"); return BEAR_ERR_ARG; }
    if (lag > ml) { bear_set_error("lag %d exceeds the packed layout's maximum %d", lag, ml); return BEAR_ERR_RANGE; }
    for (int64_t i = 0; i < n; ++i) {
        int rc = encode_one(h_text + i * lag, lag, alphabet, h_kmers + i);
        if (rc) return rc;
    }
    return BEAR_OK;
}

extern "C" int bear_decode_kmers(const uint64_t* h_kmers, int64_t n, int lag, int alphabet, char* h_text) {
    int ml = bear_max_lag(alphabet);
    if (ml < 0 || lag < 1 || lag > ml || !h_text || !h_kmers || n < 0) { bear_set_error("bear_decode_kmers: bad argument"); return BEAR_ERR_ARG; }
    const char* dna = alphabet == BEAR_ALPHABET_RNA ? "ACGU" : "ACGT";
    for (int64_t i = 0; i < n; ++i) {
        uint64_t v = h_kmers[i];
        char* o = h_text + i * lag;
        if (alphabet == BEAR_ALPHABET_PROT) {
            for (int j = lag - 1; j >= 0; --j) {
                int c = int(v & 31); v >>= 5;
                o[j] = c < 20 ? kProt[c] : (c == 20 ? '[' : 'X');
            }
        } else {
            int ns = int(v >> 58);
            for (int j = lag - 1; j >= 0; --j) { o[j] = j < ns ? '[' : dna[v & 3]; v >>= 2; }
        }
    }
    return BEAR_OK;
}

// ---------------------------------------------------------------------------------------------
// file reading: the file is mapped, cut into byte ranges at line boundaries, rows are counted per
// range (pass 1) and parsed into their final slots (pass 2) by a pool of threads.
// ---------------------------------------------------------------------------------------------
namespace {

struct FileMap {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    ~FileMap() {
        if (data && size) munmap(const_cast<char*>(data), size);
        if (fd >= 0) close(fd);
    }
    bool open_ro(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        size = size_t(st.st_size);
        if (size == 0) return true;
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { size = 0; return false; }
        data = static_cast<const char*>(m);
        madvise(m, size, MADV_SEQUENTIAL);
        return true;
    }
};

inline const char* line_end(const char* p, const char* e) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(e - p)));
    return nl ? nl : e;
}

inline bool blank_line(const char* p, const char* e) {
    for (; p < e; ++p)
        if (*p != ' ' && *p != '\t' && *p != '\r') return false;
    return true;
}

// a cursor over one line [p, e)
struct Cur {
    const char* p;
    const char* e;
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
};

int expect(Cur& c, char ch, const char* what) {
    if (c.eat(ch)) return 0;
    bear_set_error("expected '%c' in %s near '%.*s'", ch, what, int(std::min<ptrdiff_t>(20, c.e - c.p)), c.p);
    return BEAR_ERR_PARSE;
}

// non-negative integer-valued JSON number -> uint32.  Digit strings, also with a ".0" / ".000" tail (what
// json.dumps writes for float arrays), take the fast path; anything else (3e0, -2, 2.5) goes through strtod on a
// bounded copy.
int parse_count(Cur& c, uint32_t* out, const char* what) {
    c.ws();
    const char* s = c.p;
    uint64_t v = 0;
    int nd = 0;
    while (s < c.e && *s >= '0' && *s <= '9' && nd < 11) { v = v * 10 + uint64_t(*s - '0'); ++s; ++nd; }
    if (nd > 0 && nd < 11 && s < c.e && *s == '.') {          // "12.0", "12.": skip a tail of zeros
        const char* z = s + 1;
        while (z < c.e && *z == '0') ++z;
        if (z == c.e || !((*z >= '1' && *z <= '9') || *z == 'e' || *z == 'E')) s = z;
    }
    if (nd > 0 && nd < 11 && (s == c.e || (*s != '.' && *s != 'e' && *s != 'E'))) {
        if (v > 4294967295ull) { bear_set_error("count %llu in %s exceeds uint32", (unsigned long long)v, what); return BEAR_ERR_RANGE; }
        *out = uint32_t(v);
        c.p = s;
        return 0;
    }
    char buf[48];
    const size_t len = std::min<size_t>(sizeof(buf) - 1, size_t(c.e - c.p));
    memcpy(buf, c.p, len);
    buf[len] = 0;
    char* end = nullptr;
    const double d = strtod(buf, &end);
    if (end == buf) { bear_set_error("expected a number in %s near '%.20s'", what, buf); return BEAR_ERR_PARSE; }
    if (!(d >= 0) || d != std::floor(d)) { bear_set_error("count %g in %s is negative or not an integer", d, what); return BEAR_ERR_PARSE; }
    if (d > 4294967295.0) { bear_set_error("count %g in %s exceeds uint32", d, what); return BEAR_ERR_RANGE; }
    *out = uint32_t(d);
    c.p += end - buf;
    return 0;
}

struct PackJob {
    int alphabet, num_ds, A1, lag;
    int64_t first_row, max_rows, stride;
    uint64_t* kmers;
    uint32_t* counts;
    bool sparse;
    // rank-sharded ingest (shard_batch > 0): of every global batch of shard_batch consecutive file rows this rank keeps
    // its contiguous slice, ceil(rows of the batch / world) rows from offset rank * that (dataloader.KmerDataset.shard)
    int64_t shard_batch = 0, total = 0, row_offset = 0;     // row_offset: index of the file's first row in the whole dataset
    int world = 1, rank = 0;

    // rows of batch b (of n_b rows) owned by the rank
    int64_t local_rows_of(int64_t n_b) const {
        const int64_t per = (n_b + world - 1) / world;
        const int64_t lo = std::min<int64_t>(int64_t(rank) * per, n_b), hi = std::min<int64_t>(int64_t(rank + 1) * per, n_b);
        return hi - lo;
    }
    // rows of the rank among dataset rows [0, g)
    int64_t rank_rows_before(int64_t g) const {
        const int64_t b = g / shard_batch, i = g - b * shard_batch;
        const int64_t n_b = std::min<int64_t>(shard_batch, total - b * shard_batch);
        int64_t in_b = 0;
        if (n_b > 0) {
            const int64_t per = (n_b + world - 1) / world;
            in_b = std::min<int64_t>(std::max<int64_t>(i - int64_t(rank) * per, 0), local_rows_of(n_b));
        }
        return b * local_rows_of(shard_batch) + in_b;                                       // batches before b are full
    }
    // output row of file row g, or -1 when the row belongs to another rank / lies outside the window
    int64_t out_row_of(int64_t g) const {
        if (shard_batch <= 0) {
            const int64_t o = g - first_row;
            return (o >= 0 && o < max_rows) ? o : -1;
        }
        g += row_offset;                                                                    // index in the whole dataset
        const int64_t b = g / shard_batch, i = g - b * shard_batch;
        const int64_t n_b = std::min<int64_t>(shard_batch, total - b * shard_batch);
        const int64_t per = (n_b + world - 1) / world;
        if (i / per != rank) return -1;
        const int64_t o = rank_rows_before(g) - rank_rows_before(row_offset);
        return o < max_rows ? o : -1;
    }
};

int check_kmer(const PackJob& j, const char* s, int len, int64_t file_row, uint64_t* out) {
    if (len != j.lag) {
        bear_set_error("row %lld: k-mer length %d differs from %d", (long long)(file_row + 1), len, j.lag);
        return BEAR_ERR_PARSE;
    }
    return encode_one(s, len, j.alphabet, out);
}

// The usual shape of a count matrix, "[[1,2,3,4,5],[...]]" with optional blanks and integer-valued decimals
// ("1.0"), parsed without the generic cursor; false = something else was met (nothing is reported: the caller
// re-parses the line with the generic routine, which either succeeds or names the error).
bool fast_tsv_counts(const PackJob& j, const char* p, const char* e, int64_t out_row) {
#define BEAR_SKIP_BLANKS() while (p < e && *p == ' ') ++p
#define BEAR_NEED(ch) do { BEAR_SKIP_BLANKS(); if (p >= e || *p != (ch)) return false; ++p; } while (0)
    BEAR_NEED('[');
    for (int g = 0; g < j.num_ds; ++g) {
        if (g) BEAR_NEED(',');
        BEAR_NEED('[');
        for (int a = 0; a < j.A1; ++a) {
            if (a) BEAR_NEED(',');
            BEAR_SKIP_BLANKS();
            uint32_t v = 0;
            int nd = 0;
            while (p < e && unsigned(*p - '0') < 10u && nd < 10) { v = v * 10u + uint32_t(*p - '0'); ++p; ++nd; }
            if (nd == 0 || nd > 9) return false;                       // up to 9 digits: below 2^32
            if (p < e && *p == '.') {                                    // "12.0", "12."
                const char* z = p + 1;
                while (z < e && *z == '0') ++z;
                if (z < e && ((*z >= '1' && *z <= '9') || *z == 'e' || *z == 'E' || *z == '.')) return false;
                p = z;
            } else if (p < e && (*p == 'e' || *p == 'E')) {
                return false;
            }
            j.counts[(int64_t(g) * j.A1 + a) * j.stride + out_row] = v;
        }
        BEAR_NEED(']');
    }
    BEAR_NEED(']');
#undef BEAR_NEED
#undef BEAR_SKIP_BLANKS
    return true;
}

int parse_tsv_line(const PackJob& j, const char* b, const char* e, int64_t file_row, int64_t out_row) {
    const char* tab = static_cast<const char*>(memchr(b, '\t', size_t(e - b)));
    if (!tab) { bear_set_error("row %lld: no tab separator", (long long)(file_row + 1)); return BEAR_ERR_PARSE; }
    int rc = check_kmer(j, b, int(tab - b), file_row, j.kmers + out_row);
    if (rc) return rc;
    if (fast_tsv_counts(j, tab + 1, e, out_row)) return 0;
    Cur c{tab + 1, e};
    if ((rc = expect(c, '[', "count matrix"))) return rc;
    for (int g = 0; g < j.num_ds; ++g) {
        if (g && (rc = expect(c, ',', "count matrix"))) return rc;
        if ((rc = expect(c, '[', "count matrix"))) return rc;
        for (int a = 0; a < j.A1; ++a) {
            if (a && (rc = expect(c, ',', "count row"))) return rc;
            if ((rc = parse_count(c, &j.counts[(int64_t(g) * j.A1 + a) * j.stride + out_row], "count row"))) return rc;
        }
        if ((rc = expect(c, ']', "count row (wrong alphabet size?)"))) return rc;
    }
    return expect(c, ']', "count matrix (wrong num_ds?)");
}

int parse_sparse_line(const PackJob& j, const char* b, const char* e, int64_t file_row, int64_t out_row) {
    const char* s1 = static_cast<const char*>(memchr(b, ';', size_t(e - b)));
    const char* s2 = s1 ? static_cast<const char*>(memchr(s1 + 1, ';', size_t(e - (s1 + 1)))) : nullptr;
    if (!s1 || !s2) { bear_set_error("row %lld: expected 3 ';'-separated fields", (long long)(file_row + 1)); return BEAR_ERR_PARSE; }
    const char* ks = b;
    while (ks < s1 && *ks == ' ') ++ks;
    const char* ke = s1;
    while (ke > ks && ke[-1] == ' ') --ke;
    int rc = check_kmer(j, ks, int(ke - ks), file_row, j.kmers + out_row);
    if (rc) return rc;
    for (int g = 0; g < j.num_ds; ++g)
        for (int a = 0; a < j.A1; ++a) j.counts[(int64_t(g) * j.A1 + a) * j.stride + out_row] = 0;
    // positions [[g,b],...] and values [v,...] are walked in lockstep
    Cur pc{s1 + 1, s2}, vc{s2 + 1, e};
    if ((rc = expect(pc, '[', "sparse positions")) || (rc = expect(vc, '[', "sparse values"))) return rc;
    bool first = true;
    for (;;) {
        pc.ws();
        if (!(pc.p < pc.e && *pc.p == '[')) break;
        ++pc.p;
        uint32_t g, a, v;
        if ((rc = parse_count(pc, &g, "sparse positions")) || (rc = expect(pc, ',', "sparse positions")) ||
            (rc = parse_count(pc, &a, "sparse positions")) || (rc = expect(pc, ']', "sparse positions"))) return rc;
        if (int(g) >= j.num_ds || int(a) >= j.A1) {
            bear_set_error("row %lld: sparse index [%u,%u] out of range", (long long)(file_row + 1), g, a);
            return BEAR_ERR_PARSE;
        }
        if (!first && (rc = expect(vc, ',', "sparse values (length differs from positions?)"))) return rc;
        if ((rc = parse_count(vc, &v, "sparse values"))) return rc;
        uint32_t& cell = j.counts[(int64_t(g) * j.A1 + a) * j.stride + out_row];
        if (uint64_t(cell) + v > 4294967295ull) { bear_set_error("row %lld: count exceeds uint32", (long long)(file_row + 1)); return BEAR_ERR_RANGE; }
        cell += v;   // duplicate indices add
        first = false;
        pc.eat(',');
    }
    if ((rc = expect(pc, ']', "sparse positions"))) return rc;
    return expect(vc, ']', "sparse values (length differs from positions?)");
}

struct Range {
    const char* b;
    const char* e;
    int64_t rows = 0;        // non-blank lines in the range
    int64_t first = 0;       // file row index of its first line
    int rc = 0;
    int64_t err_row = 0;
    std::string err;
};

int num_threads_for(size_t bytes) {
    if (bytes < (size_t(1) << 20)) return 1;
    int n = int(std::thread::hardware_concurrency());
    if (const char* env = getenv("BEAR_PACK_THREADS")) n = atoi(env);
    n = std::max(1, std::min(n, 64));
    return int(std::min<size_t>(size_t(n), bytes >> 18));
}

// splits [b, e) at line boundaries and counts the data rows of every piece
std::vector<Range> split_and_count(const char* b, const char* e) {
    const int nt = num_threads_for(size_t(e - b));
    std::vector<Range> rs(nt);
    const char* cur = b;
    for (int t = 0; t < nt; ++t) {
        const char* target = t == nt - 1 ? e : b + size_t(e - b) * size_t(t + 1) / size_t(nt);
        const char* stop = target <= cur ? cur : (target >= e ? e : line_end(target, e) + (line_end(target, e) < e ? 1 : 0));
        rs[t].b = cur;
        rs[t].e = stop;
        cur = stop;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&rs, t] {
            int64_t n = 0;
            for (const char* p = rs[t].b; p < rs[t].e;) {
                const char* le = line_end(p, rs[t].e);
                if (!blank_line(p, le)) ++n;
                p = le + 1;
            }
            rs[t].rows = n;
        });
    for (auto& x : th) x.join();
    int64_t acc = 0;
    for (auto& r : rs) { r.first = acc; acc += r.rows; }
    return rs;
}

// start of the data (after the header line, if any)
const char* skip_header(const char* b, const char* e, int header) {
    if (!header) return b;
    for (const char* p = b; p < e;) {
        const char* le = line_end(p, e);
        const bool bl = blank_line(p, le);
        p = le < e ? le + 1 : e;
        if (!bl) return p;
    }
    return e;
}

int pack_file(const char* fn, const char* path, int header, int alphabet, int num_ds, int64_t first_row, int64_t max_rows,
              uint64_t* h_kmers, uint32_t* h_counts, int64_t stride, int64_t* rows_out, int* lag_out, bool sparse,
              int64_t shard_batch = 0, int world = 1, int rank = 0, int64_t row_offset = 0, int64_t dataset_rows = -1) {
    if (!path || bear_alphabet_size(alphabet) < 0 || num_ds < 1 || first_row < 0 || max_rows < 0 || !h_kmers || !h_counts ||
        stride < max_rows || !rows_out || !lag_out || shard_batch < 0 || world < 1 || rank < 0 || rank >= world) {
        bear_set_error("%s: bad argument", fn);
        return BEAR_ERR_ARG;
    }
    FileMap fm;
    if (!fm.open_ro(path)) { bear_set_error("cannot open '%s': %s", path, strerror(errno)); return BEAR_ERR_IO; }
    const char* b = skip_header(fm.data, fm.data + fm.size, header);
    const char* e = fm.data + fm.size;
    *rows_out = 0;
    *lag_out = 0;
    if (b >= e || max_rows == 0) return BEAR_OK;
    std::vector<Range> rs = split_and_count(b, e);
    const int64_t total = rs.back().first + rs.back().rows;
    if (total <= first_row) return BEAR_OK;
    // the lag is the k-mer length of the first data row
    PackJob job{alphabet, num_ds, bear_alphabet_size(alphabet) + 1, 0, first_row, max_rows, stride, h_kmers, h_counts, sparse};
    job.shard_batch = shard_batch;
    job.total = dataset_rows >= 0 ? dataset_rows : total;
    job.row_offset = row_offset;
    job.world = world;
    job.rank = rank;
    if (shard_batch > 0 && (row_offset < 0 || row_offset + total > job.total)) {
        bear_set_error("%s: the file's %lld rows at offset %lld do not fit a dataset of %lld rows", fn, (long long)total,
                       (long long)row_offset, (long long)job.total);
        return BEAR_ERR_ARG;
    }
    for (const char* p = b; p < e;) {
        const char* le = line_end(p, e);
        if (!blank_line(p, le)) {
            const char* sep = static_cast<const char*>(memchr(p, sparse ? ';' : '\t', size_t(le - p)));
            const char* ks = p;
            const char* ke = sep ? sep : le;
            if (sparse) {
                while (ks < ke && *ks == ' ') ++ks;
                while (ke > ks && ke[-1] == ' ') --ke;
            }
            job.lag = int(ke - ks);
            break;
        }
        p = le + 1;
    }
    if (job.lag < 1) { bear_set_error("row 1: empty k-mer"); return BEAR_ERR_PARSE; }
    if (job.lag > bear_max_lag(alphabet)) {
        bear_set_error("lag %d exceeds the packed layout's maximum %d", job.lag, bear_max_lag(alphabet));
        return BEAR_ERR_RANGE;
    }
    std::vector<std::thread> th;
    for (size_t t = 0; t < rs.size(); ++t)
        th.emplace_back([&rs, &job, t] {
            Range& r = rs[t];
            int64_t file_row = r.first;
            for (const char* p = r.b; p < r.e;) {
                const char* le = line_end(p, r.e);
                if (!blank_line(p, le)) {
                    const int64_t out_row = job.out_row_of(file_row);
                    if (job.shard_batch <= 0 && file_row - job.first_row >= job.max_rows) return;
                    if (out_row >= 0) {
                        const char* end = le;
                        while (end > p && end[-1] == '\r') --end;
                        const int rc = job.sparse ? parse_sparse_line(job, p, end, file_row, out_row)
                                                  : parse_tsv_line(job, p, end, file_row, out_row);
                        if (rc) {
                            r.rc = rc;
                            r.err_row = file_row;
                            r.err = bear_last_error();
                            return;
                        }
                    }
                    ++file_row;
                }
                p = le + 1;
            }
        });
    for (auto& x : th) x.join();
    for (const Range& r : rs)      // ranges are in file order: the first failing range holds the earliest error
        if (r.rc) {
            bear_set_error("%s", r.err.c_str());
            return r.rc;
        }
    if (shard_batch > 0) {
        *rows_out = std::min(max_rows, job.rank_rows_before(row_offset + total) - job.rank_rows_before(row_offset));
    } else {
        *rows_out = std::min(max_rows, total - first_row);
    }
    *lag_out = job.lag;
    return BEAR_OK;
}

}  // namespace

extern "C" int64_t bear_count_rows(const char* path, int header) {
    FileMap fm;
    if (!path || !fm.open_ro(path)) { bear_set_error("cannot open '%s': %s", path ? path : "(null)", strerror(errno)); return BEAR_ERR_IO; }
    const char* b = skip_header(fm.data, fm.data + fm.size, header);
    const char* e = fm.data + fm.size;
    if (b >= e) return 0;
    const std::vector<Range> rs = split_and_count(b, e);
    return rs.back().first + rs.back().rows;
}

extern "C" int bear_pack_tsv(const char* path, int header, int alphabet, int num_ds,
                             int64_t first_row, int64_t max_rows,
                             uint64_t* h_kmers, uint32_t* h_counts, int64_t stride,
                             int64_t* rows_out, int* lag_out) {
    return pack_file("bear_pack_tsv", path, header, alphabet, num_ds, first_row, max_rows, h_kmers, h_counts, stride,
                     rows_out, lag_out, false);
}

extern "C" int bear_pack_sparse(const char* path, int header, int alphabet, int num_ds,
                                int64_t first_row, int64_t max_rows,
                                uint64_t* h_kmers, uint32_t* h_counts, int64_t stride,
                                int64_t* rows_out, int* lag_out) {
    return pack_file("bear_pack_sparse", path, header, alphabet, num_ds, first_row, max_rows, h_kmers, h_counts, stride,
                     rows_out, lag_out, true);
}

extern "C" int bear_pack_shard(const char* path, int sparse, int header, int alphabet, int num_ds, int64_t batch_rows,
                               int world, int rank, int64_t row_offset, int64_t dataset_rows, int64_t max_rows,
                               uint64_t* h_kmers, uint32_t* h_counts, int64_t stride, int64_t* rows_out, int* lag_out) {
    if (batch_rows < 1 || row_offset < 0) {
        bear_set_error("bear_pack_shard: bad argument");
        return BEAR_ERR_ARG;
    }
    return pack_file("bear_pack_shard", path, header, alphabet, num_ds, 0, max_rows, h_kmers, h_counts, stride, rows_out, lag_out,
                     sparse != 0, batch_rows, world, rank, row_offset, dataset_rows);
}

// ------------------------------------------------------------------------------------------------
// compact transfer format (see include/bear_b200.h)
// ------------------------------------------------------------------------------------------------
static inline int compact_kbits(int lag, int alphabet, int wire = 0) {
    if (alphabet == BEAR_ALPHABET_PROT) return 5 * lag;
    return (wire & BEAR_WIRE_START_ESC) ? 2 * lag : 2 * lag + 6;
}
static inline bool wire_ok(int wire, int alphabet) {
    const int bits = wire & 15;
    if (bits != 4 && bits != 8 && bits != 12) return false;
    if (wire & ~(15 | BEAR_WIRE_START_ESC)) return false;
    return !(wire & BEAR_WIRE_START_ESC) || alphabet != BEAR_ALPHABET_PROT;
}
static inline int64_t compact_pitch(int64_t n) { return (n + 15) / 16 * 16; }

static int pack_threads(int64_t n) {
    int nthreads = int(std::thread::hardware_concurrency());
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 16) nthreads = 16;
    if (const char* e = getenv("BEAR_PACK_THREADS")) nthreads = atoi(e) > 0 ? atoi(e) : nthreads;
    if (n < 65536) nthreads = 1;
    return nthreads;
}

extern "C" int64_t bear_compact_bytes(int64_t n, int lag, int alphabet, int G, int wire) {
    const int a = bear_alphabet_size(alphabet);
    if (a <= 0 || n < 0 || lag < 1 || lag > bear_max_lag(alphabet) || G < 1) return -1;
    if (!wire_ok(wire, alphabet)) return -1;
    const int kb = (compact_kbits(lag, alphabet, wire) + 7) / 8;
    const int64_t pitch = compact_pitch(n);
    if ((wire & 15) == 12) return pitch * kb + (pitch + pitch / 2) * int64_t(G);      // a byte + a nibble plane per group
    return pitch * kb + (pitch * (wire & 15) / 8) * int64_t(G) * (a + 1);
}

extern "C" int bear_compact_choose_wire(const uint64_t* h_kmers, const uint32_t* h_counts, int64_t stride, int64_t row0,
                                        int64_t n, int lag, int alphabet, int G) {
    const char* fn = "bear_compact_choose_wire";
    const int a = bear_alphabet_size(alphabet);
    BEAR_REQUIRE(a > 0 && G >= 1 && n >= 0 && row0 >= 0 && stride >= row0 + n, fn);
    BEAR_REQUIRE(lag >= 1 && lag <= bear_max_lag(alphabet), fn);
    if (n == 0) return 8;
    BEAR_REQUIRE(h_counts != nullptr && h_kmers != nullptr, fn);
    const int nplanes = G * (a + 1);
    const int nthreads = pack_threads(n);
    std::vector<int64_t> big15(nthreads, 0), big255(nthreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t]() {
            int64_t e15 = 0, e255 = 0;
            for (int pl = t; pl < nplanes; pl += nthreads) {
                const uint32_t* src = h_counts + int64_t(pl) * stride + row0;
                for (int64_t i = 0; i < n; ++i) {
                    e15 += src[i] >= 15u;
                    e255 += src[i] >= 255u;
                }
            }
            big15[t] = e15;
            big255[t] = e255;
        });
    }
    for (auto& th : pool) th.join();
    int64_t e15 = 0, e255 = 0;
    for (int t = 0; t < nthreads; ++t) {
        e15 += big15[t];
        e255 += big255[t];
    }
    const int64_t bytes8 = n * nplanes + 12 * e255, bytes4 = n * nplanes / 2 + 12 * e15;
    int wire = bytes4 < bytes8 ? 4 : 8;
    // 12-bit rank of the whole count vector of a (row, group): rows whose counts sum to more than nmax escape, one entry
    // per non-zero count
    {
        const int A1 = a + 1, nmx = bear_rank::nmax(A1);
        std::vector<int64_t> esc12(nthreads, 0);
        std::vector<std::thread> pool2;
        for (int t = 0; t < nthreads; ++t) {
            pool2.emplace_back([&, t]() {
                const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
                int64_t e = 0;
                for (int g = 0; g < G; ++g) {
                    const uint32_t* src = h_counts + int64_t(g) * A1 * stride + row0;
                    for (int64_t i = lo; i < hi; ++i) {
                        uint64_t sum = 0;
                        int nz = 0;
                        for (int b = 0; b < A1; ++b) {
                            const uint32_t c = src[int64_t(b) * stride + i];
                            sum += c;
                            nz += c != 0u;
                        }
                        if (sum > uint64_t(nmx)) e += nz;
                    }
                }
                esc12[t] = e;
            });
        }
        for (auto& th : pool2) th.join();
        int64_t e12 = 0;
        for (int t = 0; t < nthreads; ++t) e12 += esc12[t];
        const int64_t bytes12 = n * G * 3 / 2 + 12 * e12;
        if (bytes12 < (wire == 4 ? bytes4 : bytes8)) wire = 12;
    }
    // start-run lengths as escapes: pays when it saves a k-mer plane and few rows are start-padded
    if (alphabet != BEAR_ALPHABET_PROT) {
        const int planes_saved = (compact_kbits(lag, alphabet, 0) + 7) / 8 - (compact_kbits(lag, alphabet, BEAR_WIRE_START_ESC) + 7) / 8;
        if (planes_saved > 0) {
            int64_t padded = 0;
            for (int64_t i = 0; i < n; ++i) padded += (h_kmers[row0 + i] >> 58) != 0;
            if (12 * padded < n * planes_saved) wire |= BEAR_WIRE_START_ESC;
        }
    }
    return wire;
}

extern "C" int bear_compact_table(const uint64_t* h_kmers, const uint32_t* h_counts, int64_t stride, int64_t row0,
                                  int64_t n, int lag, int alphabet, int G, int wire, uint8_t* h_out,
                                  uint32_t* h_esc, int64_t esc_cap, int64_t* n_esc_out) {
    const char* fn = "bear_compact_table";
    const int a = bear_alphabet_size(alphabet);
    BEAR_REQUIRE(a > 0 && lag >= 1 && lag <= bear_max_lag(alphabet) && G >= 1, fn);
    BEAR_REQUIRE(n >= 0 && row0 >= 0 && stride >= row0 + n && n < (int64_t(1) << 32) && esc_cap >= 0, fn);
    BEAR_REQUIRE(n_esc_out != nullptr && wire_ok(wire, alphabet), fn);
    const int count_bits = wire & 15;
    const bool start_esc = (wire & BEAR_WIRE_START_ESC) != 0;
    *n_esc_out = 0;
    if (n == 0) return BEAR_OK;
    BEAR_REQUIRE(h_kmers && h_counts && h_out && (h_esc || esc_cap == 0), fn);
    const int A1 = a + 1, kb = (compact_kbits(lag, alphabet, wire) + 7) / 8;
    const int64_t pitch = compact_pitch(n);
    const bool dna = alphabet != BEAR_ALPHABET_PROT;
    const int nplanes = G * A1;
    const int64_t cpitch = pitch * count_bits / 8;      // bytes per count plane
    const uint32_t marker = count_bits == 4 ? 15u : 255u;
    const int nthreads = pack_threads(n);
    const int nmx = bear_rank::nmax(A1);
    // tables for the rank coding: comp[p][m] = vectors of p entries with sum m; off[N] = vectors with sum below N
    uint32_t comp[22][12] = {}, off[13] = {};
    for (int p2 = 1; p2 <= A1 && p2 < 22; ++p2)
        for (int m = 0; m <= nmx && m < 12; ++m) comp[p2][m] = bear_rank::compositions(m, p2);
    for (int N = 1; N <= nmx + 1 && N < 13; ++N) off[N] = off[N - 1] + comp[A1][N - 1];
    // k-mer planes: rows split over threads (start-run escapes per thread, i.e. ordered by row)
    std::vector<std::vector<uint32_t>> kesc(nthreads);
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            pool.emplace_back([&, t]() {
                const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
                for (int64_t i = lo; i < hi; ++i) {
                    const uint64_t code = h_kmers[row0 + i];
                    uint64_t v = code;
                    if (dna) {
                        const uint64_t pay = code & ((uint64_t(1) << 58) - 1), ns = code >> 58;
                        v = start_esc ? pay : (pay | (ns << (2 * lag)));
                        if (start_esc && ns != 0) {
                            kesc[t].push_back(0xffffffffu);
                            kesc[t].push_back(uint32_t(i));
                            kesc[t].push_back(uint32_t(ns));
                        }
                    }
                    for (int b = 0; b < kb; ++b) h_out[int64_t(b) * pitch + i] = uint8_t(v >> (8 * b));
                }
                for (int b = 0; b < kb; ++b)
                    if (t == nthreads - 1)
                        for (int64_t i = n; i < pitch; ++i) h_out[int64_t(b) * pitch + i] = 0;
            });
        }
        for (auto& th : pool) th.join();
    }
    // count planes: one plane at a time per thread (keeps the escapes ordered by plane, row)
    std::vector<std::vector<uint32_t>> esc(count_bits == 12 ? nthreads : nplanes);
    if (count_bits == 12) {
        // rank coding: per group a byte plane (low 8 bits of the rank) and a nibble plane (high 4 bits); rows are split
        // over the threads at even row indices (two rows share a nibble-plane byte)
        uint8_t* base = h_out + int64_t(kb) * pitch;
        memset(base, 0, size_t((pitch + pitch / 2) * int64_t(G)));
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            pool.emplace_back([&, t]() {
                const int64_t lo = (n * t / nthreads) & ~int64_t(1), hi = t == nthreads - 1 ? n : ((n * (t + 1) / nthreads) & ~int64_t(1));
                for (int g = 0; g < G; ++g) {
                    const uint32_t* src = h_counts + int64_t(g) * A1 * stride + row0;
                    uint8_t* lo8 = base + int64_t(g) * (pitch + pitch / 2);
                    uint8_t* hi4 = lo8 + pitch;
                    for (int64_t i = lo; i < hi; ++i) {
                        uint32_t r = bear_rank::ESCAPE, N = 0;
                        {
                            bool ok = true;
                            for (int b = 0; b < A1; ++b) {
                                const uint32_t c = src[int64_t(b) * stride + i];
                                ok = ok && c <= uint32_t(nmx);
                                N += ok ? c : 0u;
                            }
                            if (ok && N <= uint32_t(nmx)) {
                                r = off[N];
                                int rem = int(N);
                                for (int b = 0; b + 1 < A1 && rem > 0; ++b) {
                                    const int cb = int(src[int64_t(b) * stride + i]);
                                    for (int v = 0; v < cb; ++v) r += comp[A1 - 1 - b][rem - v];
                                    rem -= cb;
                                }
                            }
                        }
                        if (r == bear_rank::ESCAPE) {
                            for (int b = 0; b < A1; ++b) {
                                const uint32_t c = src[int64_t(b) * stride + i];
                                if (c != 0u) {
                                    esc[t].push_back(uint32_t(g * A1 + b));
                                    esc[t].push_back(uint32_t(i));
                                    esc[t].push_back(c);
                                }
                            }
                        }
                        lo8[i] = uint8_t(r);
                        hi4[i >> 1] |= uint8_t((r >> 8) << (4 * (i & 1)));
                    }
                }
            });
        }
        for (auto& th : pool) th.join();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            pool.emplace_back([&, t]() {
                for (int pl = t; pl < nplanes; pl += nthreads) {
                    const uint32_t* src = h_counts + int64_t(pl) * stride + row0;
                    uint8_t* dst = h_out + int64_t(kb) * pitch + int64_t(pl) * cpitch;
                    if (count_bits == 4) memset(dst, 0, size_t(cpitch));
                    for (int64_t i = 0; i < n; ++i) {
                        uint32_t c = src[i];
                        if (c >= marker) {
                            esc[pl].push_back(uint32_t(pl));
                            esc[pl].push_back(uint32_t(i));
                            esc[pl].push_back(c);
                            c = marker;
                        }
                        if (count_bits == 4) dst[i >> 1] |= uint8_t(c << (4 * (i & 1)));
                        else dst[i] = uint8_t(c);
                    }
                    if (count_bits == 8)
                        for (int64_t i = n; i < pitch; ++i) dst[i] = 0;
                }
            });
        }
        for (auto& th : pool) th.join();
    }
    int64_t total = 0;
    for (const auto& e : esc) total += int64_t(e.size() / 3);
    for (int t = 0; t < nthreads; ++t) total += int64_t(kesc[t].size() / 3);
    *n_esc_out = total;
    if (total > esc_cap) return BEAR_OK;              // caller re-calls with room for *n_esc_out entries
    int64_t o = 0;
    for (const auto& e : esc) {
        if (!e.empty()) memcpy(h_esc + o * 3, e.data(), e.size() * sizeof(uint32_t));
        o += int64_t(e.size() / 3);
    }
    for (int t = 0; t < nthreads; ++t) {                 // start-run lengths after the count escapes
        if (!kesc[t].empty()) memcpy(h_esc + o * 3, kesc[t].data(), kesc[t].size() * sizeof(uint32_t));
        o += int64_t(kesc[t].size() / 3);
    }
    return BEAR_OK;
}
