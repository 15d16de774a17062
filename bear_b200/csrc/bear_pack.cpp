// Host ingest: k-mer count text files -> packed table (uint64 k-mers + group-planar uint32 counts).
// Replaces the tf.data text path of the reference: dataloader.dataloader (dataloader.py:6-50),
// dataloader.sparse_dataloader (dataloader.py:52-109), core.tf_one_hot's symbol tables
// (core.py:142-153) and the `wc -l` row count (models/train_bear_net.py:54-55).
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bear_b200.h"
#include "bear_host.h"

static thread_local char g_err[512] = "";

void bear_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* bear_last_error(void) { return g_err; }
extern "C" int bear_version(void) { return 1; }

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
static int encode_one(const char* s, int lag, int alphabet, uint64_t* out) {
    if (alphabet == BEAR_ALPHABET_PROT) {
        uint64_t v = 0;
        for (int j = 0; j < lag; ++j) v = (v << 5) | uint64_t(sym_code(alphabet, s[j]));
        *out = v;
        return 0;
    }
    uint64_t v = 0, nstart = 0;
    bool in_prefix = true;
    for (int j = 0; j < lag; ++j) {
        int c = sym_code(alphabet, s[j]);
        if (c < 0) {
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
// file reading
// ---------------------------------------------------------------------------------------------
struct LineReader {
    FILE* fp = nullptr;
    char* buf = nullptr;
    size_t cap = 0;
    ~LineReader() { if (fp) fclose(fp); free(buf); }
    bool open(const char* path) { fp = fopen(path, "rb"); return fp != nullptr; }
    // returns length without trailing newline / CR, or -1 at EOF
    ssize_t next() {
        ssize_t n = getline(&buf, &cap, fp);
        if (n < 0) return -1;
        while (n > 0 && (buf[n - 1] == '\n' || buf[n - 1] == '\r')) buf[--n] = 0;
        return n;
    }
};

static bool blank(const char* s, ssize_t n) {
    for (ssize_t i = 0; i < n; ++i) if (s[i] != ' ' && s[i] != '\t') return false;
    return true;
}

extern "C" int64_t bear_count_rows(const char* path, int header) {
    LineReader r;
    if (!path || !r.open(path)) { bear_set_error("cannot open '%s': %s", path ? path : "(null)", strerror(errno)); return BEAR_ERR_IO; }
    int64_t rows = 0;
    bool skip = header != 0;
    ssize_t n;
    while ((n = r.next()) >= 0) {
        if (blank(r.buf, n)) continue;
        if (skip) { skip = false; continue; }
        ++rows;
    }
    return rows;
}

// parse a non-negative integer-valued number at *p (JSON number); advances *p.
static int parse_count(const char** p, uint32_t* out, const char* what) {
    char* end = nullptr;
    errno = 0;
    double v = strtod(*p, &end);
    if (end == *p) { bear_set_error("expected a number in %s near '%.20s'", what, *p); return BEAR_ERR_PARSE; }
    if (!(v >= 0) || v != std::floor(v)) { bear_set_error("count %g in %s is negative or not an integer", v, what); return BEAR_ERR_PARSE; }
    if (v > 4294967295.0) { bear_set_error("count %g in %s exceeds uint32", v, what); return BEAR_ERR_RANGE; }
    *out = uint32_t(v);
    *p = end;
    return 0;
}

static void skip_ws(const char** p) { while (**p == ' ' || **p == '\t') ++*p; }

static int expect(const char** p, char c, const char* what) {
    skip_ws(p);
    if (**p != c) { bear_set_error("expected '%c' in %s near '%.20s'", c, what, *p); return BEAR_ERR_PARSE; }
    ++*p;
    return 0;
}

struct PackArgs {
    int alphabet, num_ds, A1;
    int64_t first_row, max_rows, stride;
    uint64_t* kmers;
    uint32_t* counts;
    int64_t rows = 0;
    int lag = 0;
};

static int check_args(const char* fn, const char* path, int alphabet, int num_ds, int64_t first_row,
                      int64_t max_rows, const uint64_t* k, const uint32_t* c, int64_t stride,
                      const int64_t* rows_out, const int* lag_out) {
    if (!path || bear_alphabet_size(alphabet) < 0 || num_ds < 1 || first_row < 0 || max_rows < 0 ||
        !k || !c || stride < max_rows || !rows_out || !lag_out) {
        bear_set_error("%s: bad argument", fn);
        return BEAR_ERR_ARG;
    }
    return 0;
}

static int store_kmer(PackArgs& a, const char* s, int len, int64_t file_row) {
    if (a.lag == 0) {
        if (len < 1) { bear_set_error("row %lld: empty k-mer", (long long)file_row); return BEAR_ERR_PARSE; }
        if (len > bear_max_lag(a.alphabet)) { bear_set_error("lag %d exceeds the packed layout's maximum %d", len, bear_max_lag(a.alphabet)); return BEAR_ERR_RANGE; }
        a.lag = len;
    } else if (len != a.lag) {
        bear_set_error("row %lld: k-mer length %d differs from %d", (long long)file_row, len, a.lag);
        return BEAR_ERR_PARSE;
    }
    return encode_one(s, len, a.alphabet, a.kmers + a.rows);
}

extern "C" int bear_pack_tsv(const char* path, int header, int alphabet, int num_ds,
                             int64_t first_row, int64_t max_rows,
                             uint64_t* h_kmers, uint32_t* h_counts, int64_t stride,
                             int64_t* rows_out, int* lag_out) {
    int rc = check_args("bear_pack_tsv", path, alphabet, num_ds, first_row, max_rows, h_kmers, h_counts, stride, rows_out, lag_out);
    if (rc) return rc;
    LineReader r;
    if (!r.open(path)) { bear_set_error("cannot open '%s': %s", path, strerror(errno)); return BEAR_ERR_IO; }
    PackArgs a{alphabet, num_ds, bear_alphabet_size(alphabet) + 1, first_row, max_rows, stride, h_kmers, h_counts};
    bool skip = header != 0;
    int64_t file_row = 0;
    ssize_t n;
    while (a.rows < max_rows && (n = r.next()) >= 0) {
        if (blank(r.buf, n)) continue;
        if (skip) { skip = false; continue; }
        if (file_row++ < first_row) continue;
        const char* tab = (const char*)memchr(r.buf, '\t', size_t(n));
        if (!tab) { bear_set_error("row %lld: no tab separator", (long long)file_row); return BEAR_ERR_PARSE; }
        if ((rc = store_kmer(a, r.buf, int(tab - r.buf), file_row))) return rc;
        const char* p = tab + 1;
        if ((rc = expect(&p, '[', "count matrix"))) return rc;
        for (int g = 0; g < num_ds; ++g) {
            if (g && (rc = expect(&p, ',', "count matrix"))) return rc;
            if ((rc = expect(&p, '[', "count matrix"))) return rc;
            for (int b = 0; b < a.A1; ++b) {
                if (b && (rc = expect(&p, ',', "count row"))) return rc;
                skip_ws(&p);
                if ((rc = parse_count(&p, &h_counts[(int64_t(g) * a.A1 + b) * stride + a.rows], "count row"))) return rc;
            }
            if ((rc = expect(&p, ']', "count row (wrong alphabet size?)"))) return rc;
        }
        if ((rc = expect(&p, ']', "count matrix (wrong num_ds?)"))) return rc;
        ++a.rows;
    }
    *rows_out = a.rows;
    *lag_out = a.lag;
    return BEAR_OK;
}

extern "C" int bear_pack_sparse(const char* path, int header, int alphabet, int num_ds,
                                int64_t first_row, int64_t max_rows,
                                uint64_t* h_kmers, uint32_t* h_counts, int64_t stride,
                                int64_t* rows_out, int* lag_out) {
    int rc = check_args("bear_pack_sparse", path, alphabet, num_ds, first_row, max_rows, h_kmers, h_counts, stride, rows_out, lag_out);
    if (rc) return rc;
    LineReader r;
    if (!r.open(path)) { bear_set_error("cannot open '%s': %s", path, strerror(errno)); return BEAR_ERR_IO; }
    PackArgs a{alphabet, num_ds, bear_alphabet_size(alphabet) + 1, first_row, max_rows, stride, h_kmers, h_counts};
    bool skip = header != 0;
    int64_t file_row = 0;
    ssize_t n;
    std::vector<std::pair<int, int>> pos;
    while (a.rows < max_rows && (n = r.next()) >= 0) {
        if (blank(r.buf, n)) continue;
        if (skip) { skip = false; continue; }
        if (file_row++ < first_row) continue;
        const char* s1 = (const char*)memchr(r.buf, ';', size_t(n));
        const char* s2 = s1 ? (const char*)memchr(s1 + 1, ';', size_t(n - (s1 + 1 - r.buf))) : nullptr;
        if (!s1 || !s2) { bear_set_error("row %lld: expected 3 ';'-separated fields", (long long)file_row); return BEAR_ERR_PARSE; }
        const char* ks = r.buf;
        while (*ks == ' ') ++ks;
        const char* ke = s1;
        while (ke > ks && ke[-1] == ' ') --ke;
        if ((rc = store_kmer(a, ks, int(ke - ks), file_row))) return rc;
        for (int g = 0; g < num_ds; ++g)
            for (int b = 0; b < a.A1; ++b) h_counts[(int64_t(g) * a.A1 + b) * stride + a.rows] = 0;
        // positions [[g,b],...]
        pos.clear();
        const char* p = s1 + 1;
        if ((rc = expect(&p, '[', "sparse positions"))) return rc;
        skip_ws(&p);
        while (*p == '[') {
            ++p;
            uint32_t g, b;
            skip_ws(&p);
            if ((rc = parse_count(&p, &g, "sparse positions"))) return rc;
            if ((rc = expect(&p, ',', "sparse positions"))) return rc;
            skip_ws(&p);
            if ((rc = parse_count(&p, &b, "sparse positions"))) return rc;
            if ((rc = expect(&p, ']', "sparse positions"))) return rc;
            if (int(g) >= num_ds || int(b) >= a.A1) { bear_set_error("row %lld: sparse index [%u,%u] out of range", (long long)file_row, g, b); return BEAR_ERR_PARSE; }
            pos.emplace_back(int(g), int(b));
            skip_ws(&p);
            if (*p == ',') { ++p; skip_ws(&p); }
        }
        if ((rc = expect(&p, ']', "sparse positions"))) return rc;
        // values [v,...]
        p = s2 + 1;
        if ((rc = expect(&p, '[', "sparse values"))) return rc;
        for (size_t i = 0; i < pos.size(); ++i) {
            if (i && (rc = expect(&p, ',', "sparse values"))) return rc;
            skip_ws(&p);
            uint32_t v;
            if ((rc = parse_count(&p, &v, "sparse values"))) return rc;
            uint32_t& cell = h_counts[(int64_t(pos[i].first) * a.A1 + pos[i].second) * stride + a.rows];
            if (uint64_t(cell) + v > 4294967295ull) { bear_set_error("row %lld: count exceeds uint32", (long long)file_row); return BEAR_ERR_RANGE; }
            cell += v;   // duplicate indices add, as in tf.sparse.to_dense after reorder
        }
        if ((rc = expect(&p, ']', "sparse values (length differs from positions?)"))) return rc;
        ++a.rows;
    }
    *rows_out = a.rows;
    *lag_out = a.lag;
    return BEAR_OK;
}
