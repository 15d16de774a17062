// Linear AR head as a gather from per-chunk tables of exp-ratios (ar_funcs.make_ar_func_linear, ar_funcs.py:23-46),
// shared by the fused train and evaluation kernels.
#pragma once
#include <math.h>

#include "bear_common.cuh"
#include "bear_dm_row.cuh"

namespace {

using namespace bear;

constexpr int CHUNK = 4;         // at most 4 positions per chunk table (forward ratios R, gradient G)
constexpr int THREADS = 256;
constexpr int MAX_GRID = 148 * 4;
constexpr int TABN = 64;         // counts below TABN index the per-CTA tables of row-independent terms
constexpr uint64_t PAYLOAD_MASK = (1ull << 58) - 1;

__host__ __device__ inline int num_chunks(int lag) { return (lag + CHUNK - 1) / CHUNK; }

// The lag positions are spread as evenly as possible over the chunks (13 = 4+3+3+3, not 4+4+4+1):
// a chunk with very few keys would make every row of a tile collide in the gradient scatter.
struct ChunkGeom {
    int start, size;
};
struct ChunkKeys {               // chunk sizes: base + 1 for the first `extra` chunks, base for the rest
    int base, extra;
};
__host__ __device__ inline ChunkGeom chunk_geom(int lag, int nch, int ch) {
    const int base = lag / nch, extra = lag % nch;
    ChunkGeom g;
    g.size = base + (ch < extra ? 1 : 0);
    g.start = ch * base + (ch < extra ? ch : extra);
    return g;
}

__device__ __forceinline__ int symbol_at(uint64_t v, int j, int lag, int nstart) {
    return j < nstart ? 4 : int((v >> (2 * (lag - 1 - j))) & 3u);
}

// Fixed-order second stage: out[p] += mult * sum_blk partials[blk, p]
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblk, int P, double mult,
                                       double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partials[int64_t(b) * P + p];
    out[p] += mult * s;
}

constexpr int ENT = 344;            // entries per extended chunk table: 256 + 85, padded to a multiple of 8

__host__ __device__ inline int ext_entries(int r) { return (1 << (2 * r)) + ((1 << (2 * r)) - 1) / 3; }

// key of a chunk of r positions whose first s positions are start symbols; `plain` = the chunk's payload bits
// (zero under the start run)
__device__ __forceinline__ int ext_key(uint32_t plain, int r, int s) {
    if (s <= 0) return int(plain);
    if (s > r) s = r;
    const int full = 1 << (2 * r);
    return full + (full - (1 << (2 * (r - s + 1)))) / 3 + int(plain & ((1u << (2 * (r - s))) - 1u));
}

// symbols (0..3 letters, 4 start) of entry idx of a chunk of r positions, 3 bits each, position 0 in the low
// bits; 0xffff for an index past the last entry
__device__ inline uint16_t ext_symbols(int idx, int r) {
    const int full = 1 << (2 * r);
    int s = 0, v = idx;
    if (idx >= full) {
        int e = idx - full;
        s = 1;
        while (s <= r && e >= (1 << (2 * (r - s)))) {
            e -= 1 << (2 * (r - s));
            ++s;
        }
        if (s > r) return 0xffff;
        v = e;
    }
    uint32_t out = 0;
    for (int p = 0; p < r; ++p) {
        const uint32_t sym = p < s ? 4u : uint32_t((v >> (2 * (r - 1 - p))) & 3);
        out |= sym << (3 * p);
    }
    return uint16_t(out);
}

// Rows of the chunk tables are 32 bytes (4 doubles) and are read as two 128-bit halves.  A warp instruction touches
// the same half of 32 random rows, which would use only every other 16-byte bank group; swapping the halves of
// rows with bit 2 of the key set spreads a half over all eight groups (about a third fewer conflict wavefronts).
__device__ __forceinline__ int half_swizzle(int q) { return (q >> 1) & 2; }      // offset (doubles) of logical half 0

// exact softmax(sum_j mat[j, s_j, :]) of one k-mer from the weight table in global memory: only taken when
// the ratio product of the chunk tables left the double range (logit spreads of several hundred)
__device__ __noinline__ void linear_head_exact(const double* __restrict__ mat, uint64_t code, int lag, double (&f)[A1]) {
    const int ns = int(code >> 58);
    const uint64_t v = code & PAYLOAD_MASK;
    double l[A1] = {0, 0, 0, 0, 0};
    for (int j = 0; j < lag; ++j) {
        const double* row = mat + (j * A1 + symbol_at(v, j, lag, ns)) * A1;
        for (int b = 0; b < A1; ++b) l[b] += __ldg(row + b);
    }
    double m = l[0], z = 0.0;
    for (int b = 1; b < A1; ++b) m = fmax(m, l[b]);
    for (int b = 0; b < A1; ++b) {
        f[b] = exp(l[b] - m);
        z += f[b];
    }
    for (int b = 0; b < A1; ++b) f[b] /= z;
}

// Extended, swizzled ratio tables R[nch][ENT][4] of the linear head and the symbol table they are built from.
// Needs a __syncthreads() by the caller afterwards.
__device__ void build_ext_tables(const double* __restrict__ mat, double* R, uint16_t* symtab, int lag, const ChunkKeys& ck) {
    const int nch = num_chunks(lag);
    for (int i = threadIdx.x; i < 2 * ENT; i += blockDim.x) {
        const int r = ck.base + (i < ENT ? 1 : 0);           // size class 0: base + 1 positions, class 1: base
        symtab[i] = (r >= 1 && r <= CHUNK) ? ext_symbols(i % ENT, r) : uint16_t(0xffff);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nch * ENT; idx += blockDim.x) {
        const int ch = idx / ENT, e = idx - ch * ENT;
        const ChunkGeom cg = chunk_geom(lag, nch, ch);
        const uint32_t syms = symtab[(ch < ck.extra ? 0 : ENT) + e];
        double l[A1] = {0, 0, 0, 0, 0};
        if (syms != 0xffffu) {
            for (int p = 0; p < cg.size; ++p) {
                const double* row = mat + ((cg.start + p) * A1 + int((syms >> (3 * p)) & 7u)) * A1;
#pragma unroll
                for (int b = 0; b < A1; ++b) l[b] += __ldg(row + b);
            }
        }
        const int sw = half_swizzle(e);
#pragma unroll
        for (int b = 0; b < 4; ++b) R[idx * 4 + ((b & 2) ^ sw) + (b & 1)] = exp(l[b] - l[4]);
    }
}

// softmax(sum_j mat[j, s_j, :]) of one k-mer as the normalised product of its chunk-table rows
__device__ __forceinline__ void linear_head_ext(const double* R, const double* __restrict__ mat, uint64_t code, int lag,
                                                const ChunkKeys& ck, int nch, double (&f)[A1]) {
    const int ns = int(code >> 58);
    const uint64_t v = code & PAYLOAD_MASK;
    double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
    int sh = 2 * lag, c0 = 0;
    for (int ch = 0; ch < nch; ++ch) {
        const int rr = ck.base + (ch < ck.extra ? 1 : 0);
        sh -= 2 * rr;
        int q = int(uint32_t(v >> sh) & ((1u << (2 * rr)) - 1u));
        if (ns > c0) q = ext_key(uint32_t(q), rr, ns - c0);
        c0 += rr;
        const int sw = half_swizzle(q);
        const double2 a = *reinterpret_cast<const double2*>(R + (ch * ENT + q) * 4 + sw);
        const double2 b = *reinterpret_cast<const double2*>(R + (ch * ENT + q) * 4 + (sw ^ 2));
        p0 *= a.x;
        p1 *= a.y;
        p2 *= b.x;
        p3 *= b.y;
    }
    const double z = 1.0 + ((p0 + p1) + (p2 + p3));
    if (z < 1e300 && z > 1e-300) {
        const double zi = 1.0 / z;
        f[0] = p0 * zi;
        f[1] = p1 * zi;
        f[2] = p2 * zi;
        f[3] = p3 * zi;
        f[4] = zi;
    } else {
        linear_head_exact(mat, code, lag, f);
    }
}


// Chunk geometry of the head tables, precomputed on the host (kernel parameter: its fields are constant-bank operands):
// chunk ch covers positions [c0, c0 + rr) and its key sits `sh` bits above the low end of the packed k-mer.
struct HeadGeom {
    uint8_t sh[8], rr[8], c0[8];
};

inline HeadGeom make_head_geom(int lag) {
    HeadGeom hg;
    const int nch = num_chunks(lag);
    for (int ch = 0; ch < 8; ++ch) {
        const ChunkGeom cg = chunk_geom(lag, nch, ch < nch ? ch : nch - 1);
        hg.rr[ch] = uint8_t(cg.size);
        hg.c0[ch] = uint8_t(cg.start);
        hg.sh[ch] = uint8_t(2 * (lag - cg.start - cg.size));
    }
    return hg;
}

// softmax(sum_j mat[j, s_j, :]) of one k-mer as the normalised product of its chunk-table rows, with the chunk loop fully
// unrolled over the precomputed geometry (NCH > 0: compile-time chunk count; NCH = 0: up to 8 chunks, `nch` at run time).
// Rows are 32 bytes; the two 16-byte halves are swapped when bit 2 of the key is set (half_swizzle).
// `any_start` (warp-uniform): some k-mer of the warp's tile is start-padded.  Start-padded rows are rare (only the first
// lag positions of a sequence), so the usual tile takes the loop without the extended-key arithmetic -- as predicated
// code it would issue for every row and chunk.
template <int NCH>
__device__ __forceinline__ void linear_head_geom(const double* R, const double* __restrict__ mat, uint64_t code, int lag,
                                                 const HeadGeom& hg, int nch, double (&f)[A1], bool any_start = true) {
    const int ns = int(code >> 58);
    const uint64_t v5 = (code & PAYLOAD_MASK) << 5;
    double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
    auto chunk = [&](int ch, uint32_t q32) {
        const uint32_t o = q32 | ((q32 >> 3) & 16u);
        const unsigned char* row = reinterpret_cast<const unsigned char*>(R) + ch * (ENT * 32);
        const double2 a = *reinterpret_cast<const double2*>(row + o);
        const double2 b = *reinterpret_cast<const double2*>(row + (o ^ 16u));
        p0 *= a.x;
        p1 *= a.y;
        p2 *= b.x;
        p3 *= b.y;
    };
    if (any_start) {
#pragma unroll
        for (int ch = 0; ch < (NCH ? NCH : 8); ++ch) {
            if (!NCH && ch >= nch) break;
            const int rr = hg.rr[ch];
            uint32_t q32 = uint32_t(v5 >> hg.sh[ch]) & (((1u << (2 * rr)) - 1u) << 5);     // 32 * key
            if (ns > hg.c0[ch]) q32 = uint32_t(ext_key(q32 >> 5, rr, ns - hg.c0[ch])) << 5;
            chunk(ch, q32);
        }
    } else {
#pragma unroll
        for (int ch = 0; ch < (NCH ? NCH : 8); ++ch) {
            if (!NCH && ch >= nch) break;
            chunk(ch, uint32_t(v5 >> hg.sh[ch]) & (((1u << (2 * hg.rr[ch])) - 1u) << 5));
        }
    }
    const double z = 1.0 + ((p0 + p1) + (p2 + p3));
    if (z < 1e300 && z > 1e-300) {
        const double zi = 1.0 / z;
        f[0] = p0 * zi;
        f[1] = p1 * zi;
        f[2] = p2 * zi;
        f[3] = p3 * zi;
        f[4] = zi;
    } else {
        linear_head_exact(mat, code, lag, f);
    }
}

struct RowIn {
    uint64_t code;
    uint32_t c[A1];
};

__device__ __forceinline__ RowIn load_row(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ col,
                                          int64_t stride, int64_t i, int64_t n) {
    RowIn r;
    const bool ok = i < n;
    r.code = (ok && kmers) ? __ldg(kmers + i) : 0ull;
#pragma unroll
    for (int b = 0; b < A1; ++b) r.c[b] = ok ? __ldg(col + b * stride + i) : 0u;
    return r;
}

inline ChunkKeys make_chunk_keys(int lag) {
    ChunkKeys ck;
    const int nch = num_chunks(lag);
    ck.base = lag / nch;
    ck.extra = lag % nch;
    return ck;
}

}  // namespace
