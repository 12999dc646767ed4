// latok_tokbytes.cu -- token spans as BYTE ranges of the packed UTF-8 buffer.
//
// The reference materialises every token with `text[s:e].strip()` (default_tokenizer.py:151-158, 47-59 % of its
// CPU time).  Here the (start, end) code-point pairs written by the tokenize kernel are turned into byte offsets into
// the flat buffer, already trimmed the way strip() trims them (a span holds whitespace only at its first character,
// SURVEY §8 A5), so that utf8[b:e] IS the token text and an Arrow-style (data, begin, end) view needs no Python loop.
//
//   lead_count_kernel   characters (non-continuation bytes) per 512-byte block, as an exclusive prefix inside each
//                       group of 256 blocks + one total per group
//   group_scan_kernel   exclusive prefix of the group totals (one CTA)
//   token_bytes_kernel  one warp per string, one lane per token: ASCII-only strings map index -> byte directly; other
//                       strings find the block by bisection on the prefix and the byte by population counts
// All three are HBM/L2-bound integer kernels; no tensor cores.
#include "latok_device.cuh"

namespace latok {

constexpr int TB_BLOCK = 512;            // bytes per counted block (one coalesced 16-byte load per lane)
constexpr int TB_GROUP = 256;            // blocks per CTA of lead_count_kernel (8 warps x 32 blocks)

__device__ __forceinline__ unsigned leads16(const uint4 v)
{
    // bytes that are not UTF-8 continuation bytes (10xxxxxx)
    auto cnt = [](uint32_t w) { return 4u - (unsigned)__popc((w >> 7) & ~(w >> 6) & 0x01010101u); };
    return cnt(v.x) + cnt(v.y) + cnt(v.z) + cnt(v.w);
}

__device__ __forceinline__ uint4 load16_clamped(const uint8_t *in, long long pos, long long n_bytes)
{
    // bytes at and beyond n_bytes read as continuation bytes (0x80): they count as no character
    if (pos + 16 <= n_bytes) return *reinterpret_cast<const uint4 *>(in + pos);
    uint32_t w[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};
    for (int q = 0; q < 16; ++q)
        if (pos + q < n_bytes) w[q >> 2] = (w[q >> 2] & ~(0xFFu << (8 * (q & 3)))) | ((uint32_t)in[pos + q] << (8 * (q & 3)));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) lead_count_kernel(const uint8_t *in, long long n_bytes, long long nblocks,
                                                         unsigned *blk_local, unsigned *group_tot)
{
    __shared__ unsigned wtot[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b0 = ((long long)blockIdx.x * 8 + warp) * 32;       // this warp's 32 blocks
    unsigned mine = 0;                                                   // lane i keeps the count of block b0 + i
    for (int i = 0; i < 32; ++i) {
        const long long b = b0 + i;
        if (b >= nblocks) break;
        const unsigned c = __reduce_add_sync(0xFFFFFFFFu, leads16(load16_clamped(in, b * TB_BLOCK + lane * 16, n_bytes)));
        if (lane == i) mine = c;
    }
    unsigned inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) before += wtot[w]; total += wtot[w]; }
    if (b0 + lane < nblocks) blk_local[b0 + lane] = before + inc - mine;
    if (threadIdx.x == 0) group_tot[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) group_scan_kernel(const unsigned *group_tot, long long ngroups, unsigned long long *group_pref)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < ngroups; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < ngroups ? group_tot[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        unsigned long long before = carry_s;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (i < ngroups) group_pref[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + inc;
        __syncthreads();
    }
}

struct TokBytesParams {
    const uint8_t *in;
    long long n_bytes;
    const long long *offsets, *char_off, *tok_off;
    long long n_strings;
    const int32_t *spans;
    long long *out;                 // [T,2]
    const unsigned *blk_local;
    const unsigned long long *group_pref;
    long long nblocks;
    const uint8_t *table_blob;
    TableLayout tl;
};

__global__ void __launch_bounds__(256) token_bytes_kernel(const TokBytesParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    Tables tb;
    tb.ascii_feat = reinterpret_cast<const uint16_t *>(p.table_blob + p.tl.ascii_feat);
    tb.class_feat = reinterpret_cast<const uint16_t *>(p.table_blob + p.tl.class_feat);
    tb.stage1 = p.table_blob + p.tl.stage1; tb.stage2 = p.table_blob + p.tl.stage2;
    tb.low_limit = p.tl.low_limit; tb.high_first = p.tl.high_first; tb.high_last = p.tl.high_last; tb.high_feat = p.tl.high_feat;
    auto pref = [&](long long b) -> unsigned long long { return p.group_pref[b / TB_GROUP] + p.blk_local[b]; };
    for (long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < p.n_strings; s += warps) {
        const long long o0 = p.offsets[s], o1 = p.offsets[s + 1], c0 = p.char_off[s], c1 = p.char_off[s + 1];
        const long long k0 = p.tok_off[s], k1 = p.tok_off[s + 1];
        const bool ascii = (o1 - o0) == (c1 - c0);
        // byte position of the string's character number idx (0 <= idx <= length; length -> end of the string)
        auto byte_of = [&](long long idx) -> long long {
            if (ascii) return o0 + idx;
            if (idx >= c1 - c0) return o1;
            const unsigned long long g = (unsigned long long)(c0 + idx);        // global character number
            long long lo = o0 / TB_BLOCK, hi = (o1 - 1) / TB_BLOCK;              // last block whose prefix is <= g
            while (lo < hi) {
                const long long mid = (lo + hi + 1) >> 1;
                if (pref(mid) <= g) lo = mid; else hi = mid - 1;
            }
            unsigned r = (unsigned)(g - pref(lo));                               // characters of the block to skip
            long long pos = lo * TB_BLOCK;
            for (int i = 0; i < TB_BLOCK / 16; ++i, pos += 16) {
                const uint4 v = load16_clamped(p.in, pos, p.n_bytes);
                const unsigned c = leads16(v);
                if (r < c) {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    for (int q = 0; q < 16; ++q) {
                        const uint32_t b = (w[q >> 2] >> (8 * (q & 3))) & 0xFFu;
                        if ((b & 0xC0u) != 0x80u) { if (r == 0) return pos + q; --r; }
                    }
                }
                r -= c;
            }
            return o1;     // (not reached on well-formed input)
        };
        for (long long k = k0 + lane; k < k1; k += 32) {
            const int2 sp = reinterpret_cast<const int2 *>(p.spans)[k];
            long long bs = byte_of(sp.x);
            const long long be = byte_of(sp.y);
            // strip(): the span's first character is the only one that can be whitespace
            if (bs < be) {
                const uint8_t *q = p.in + bs;
                const uint32_t b0 = q[0];
                uint32_t w;
                int len = 1;
                if (b0 < 0x80u) w = tb.ascii_feat[b0];
                else {
                    len = b0 >= 0xF0u ? 4 : (b0 >= 0xE0u ? 3 : 2);
                    uint32_t v = b0;
                    for (int j = 1; j < len && bs + j < p.n_bytes; ++j) v |= (uint32_t)q[j] << (8 * j);
                    w = mb_features(v, tb);
                }
                if ((w >> PL_SP) & 1u) bs = min(bs + len, be);
            }
            p.out[2 * k] = bs; p.out[2 * k + 1] = be;
        }
    }
}

cudaError_t launch_token_bytes(const uint8_t *in, long long n_bytes, const long long *offsets, const long long *char_off,
                               const long long *tok_off, long long n_strings, const int32_t *spans, long long *out,
                               unsigned *blk_local, unsigned *group_tot, unsigned long long *group_pref,
                               const uint8_t *table_blob, const TableLayout &tl, int n_sm, cudaStream_t s)
{
    const long long nblocks = token_bytes_blocks(n_bytes), ngroups = token_bytes_groups(n_bytes);
    if (n_strings == 0) return cudaSuccess;
    lead_count_kernel<<<(unsigned)ngroups, 256, 0, s>>>(in, n_bytes, nblocks, blk_local, group_tot);
    group_scan_kernel<<<1, 1024, 0, s>>>(group_tot, ngroups, group_pref);
    TokBytesParams p;
    p.in = in; p.n_bytes = n_bytes; p.offsets = offsets; p.char_off = char_off; p.tok_off = tok_off; p.n_strings = n_strings;
    p.spans = spans; p.out = out; p.blk_local = blk_local; p.group_pref = group_pref; p.nblocks = nblocks;
    p.table_blob = table_blob; p.tl = tl;
    const long long want = (n_strings + 7) / 8;
    const long long cap = (long long)n_sm * 8 * 4;
    token_bytes_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(p);
    return cudaGetLastError();
}

long long token_bytes_blocks(long long n_bytes) { return n_bytes / TB_BLOCK + 1; }
long long token_bytes_groups(long long n_bytes) { return (token_bytes_blocks(n_bytes) + TB_GROUP - 1) / TB_GROUP; }

}  // namespace latok
