// latok_kernels.cu -- sm_100a kernels for LaTok's tokenization hot path.
//
// One persistent kernel (tokenize_kernel) makes a single pass over the flat UTF-8 buffer (bit-plane form,
// 32 characters per register):
//
//   TMA bulk copy of a 16 KB window (owned tile + halos) into shared memory
//   phase 1  byte space : lead-byte detection, UTF-8 decode, class-table lookup (shared memory)
//                         -> 12 base feature bits per character           (latok.c:15-29, 77-98)
//   phase 2a char space : prev / next / after-next context features        (latok.c:68-73, 99-134)
//                         + rule evaluation split_cnt / mark / sym         (latok.c:318-341 with
//                           default_tokenizer.py:49-55, 80-91, 100-102)
//   chain 1  decoupled look-back: character count, current string start, block-mask backlog
//   phase 2b block mask in scan form (latok.c:218-244), closing the chunk that is open at the
//            tile end from the right halo (or a look-ahead walk for very long chunks)
//   phase 3  split values (default_tokenizer.py:121-132), token start/end flags
//            (default_tokenizer.py:148-158)
//   chain 2  decoupled look-back: token ordinal (+ feature sums of the token open at the boundary)
//   phase 4  emission: int8 split mask, int32 spans, CSR offsets, int8 token feature sums
//            (latok.c:342-354), int8 feature matrix
//
// No tensor cores: nothing on this path is a dense contraction; the bound is HBM bandwidth.
#include "latok_internal.h"

namespace latok {

// ---- word layout: bits 0..24 = the 25 feature columns (offsets.py:24-48), then flags ------------
constexpr uint32_t FIRSTBIT = 1u << 25;   // character starts a string
constexpr uint32_t LASTBIT = 1u << 26;    // character ends a string
constexpr uint32_t FEATMASK = (1u << NFEAT) - 1;
constexpr int NEG = -(1 << 28);           // "-infinity" of the (max,+) backlog functions
constexpr unsigned SPIN_LIMIT = 1u << 27; // watchdog for look-back spins

// ---- small helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mask_lt(int k) { return k >= 32 ? 0xFFFFFFFFu : (k <= 0 ? 0u : ((1u << k) - 1u)); }
// bits of a thread's 32-character word that fall inside the character range [lo, hi)
__device__ __forceinline__ uint32_t range_mask(int base, int lo, int hi)
{
    return mask_lt(hi - base) & ~mask_lt(lo - base);
}
__device__ __forceinline__ int widx(int c) { return 1 + c + ((c + 32) >> 5); }  // bank-skewed slot of character c >= -1

struct Fn { int u, v; };  // x -> max(x + u, v)
__device__ __forceinline__ Fn fn_id() { return Fn{0, NEG}; }
__device__ __forceinline__ Fn fn_compose(Fn f, Fn g)  // g after f
{
    return Fn{max(f.u + g.u, NEG), max(max(f.v + g.u, g.v), NEG)};
}
__device__ __forceinline__ int fn_apply(Fn f, int x) { return max(x + f.u, f.v); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- Unicode class lookup (replaces gettyperecord, latok.c:15-29, + the tests of latok.c:87-98) ----
struct Tables {
    const uint16_t *ascii_feat;
    const uint16_t *class_feat;
    const uint8_t *stage1;
    const uint8_t *stage2;
    uint32_t low_limit, high_first, high_last, high_feat;
};

__device__ __forceinline__ uint32_t class_of_cp(uint32_t cp, const Tables &t)
{
    if (cp < 0x80u) return t.ascii_feat[cp];
    if (cp < t.low_limit) {
        uint32_t blk = t.stage1[cp >> 7];
        uint32_t b = t.stage2[blk * 64u + ((cp & 127u) >> 1)];
        return t.class_feat[(cp & 1u) ? (b >> 4) : (b & 15u)];
    }
    return (cp >= t.high_first && cp <= t.high_last) ? t.high_feat : 0u;
}

// p points at a character's first byte; up to 3 following bytes are read unconditionally
template <class BytePtr>
__device__ __forceinline__ uint32_t classify_at(BytePtr p, const Tables &t)
{
    uint32_t b0 = p[0];
    if (b0 < 0x80u) return t.ascii_feat[b0];
    if (b0 < 0xC0u || b0 >= 0xF8u) return 0u;  // stray continuation / invalid lead: no features
    uint32_t cp;
    if (b0 < 0xE0u) cp = ((b0 & 0x1Fu) << 6) | (p[1] & 0x3Fu);
    else if (b0 < 0xF0u) cp = ((b0 & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu);
    else cp = ((b0 & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu);
    return class_of_cp(cp, t);
}

// ---- context features (latok.c:68-73, 99-134) ------------------------------------------------------
// pw / nw / aw: words of the previous / next / after-next character (only base bits are used);
// F / L: this character starts / ends its string; L2: the next character ends the string.
__device__ __forceinline__ uint32_t make_word(uint32_t pw, uint32_t w, uint32_t nw, uint32_t aw, bool F, bool L, bool L2)
{
    pw = F ? 0x20u : pw;                 // start of string behaves as a space (latok.c:69-73,114-117)
    nw = L ? 0x20u : nw;                 // end of string behaves as a space   (latok.c:122-130)
    aw = (L || L2) ? 0u : aw;            // latok.c:131-134
    uint32_t x = w & 0xFFFu;
    x |= ((pw >> 0) & 1u) << 12;         // PREV_ALPHA
    x |= ((nw >> 0) & 1u) << 13;         // NEXT_ALPHA
    x |= ((pw >> 1) & 1u) << 14;         // PREV_ALPHA_NUM
    x |= ((nw >> 1) & 1u) << 15;         // NEXT_ALPHA_NUM
    x |= ((pw >> 3) & 1u) << 16;         // PREV_LOWER
    x |= ((nw >> 3) & 1u) << 17;         // NEXT_LOWER
    x |= ((pw >> 5) & 1u) << 18;         // PREV_SPACE
    x |= ((nw >> 5) & 1u) << 19;         // NEXT_SPACE
    x |= ((pw >> 6) & 1u) << 20;         // PREV_SYMBOL
    x |= ((nw >> 8) & 1u) << 21;         // NEXT_AT
    x |= ((nw >> 10) & 1u) << 22;        // NEXT_SLASH
    x |= ((aw >> 0) & 1u) << 23;         // AFTER_NEXT_ALPHA
    x |= ((aw >> 10) & 1u) << 24;        // AFTER_NEXT_SLASH
    return x;
}

// ---- rule evaluation (combine_matrix_rows 2-D, latok.c:318-341) -------------------------------------
__device__ __forceinline__ void eval_rules(const RuleSet &c_rules, uint32_t w, uint32_t &cnt, bool &mark, uint32_t &sym)
{
    if (c_rules.is_default) {
        // C_SPLIT: SPACE + SYMBOL + PREV_SYMBOL + UPPER*NEXT_LOWER + UPPER*PREV_LOWER (default_tokenizer.py:49-55)
        uint32_t upper = (w >> 4) & 1u;
        cnt = ((w >> 5) & 1u) + ((w >> 6) & 1u) + ((w >> 20) & 1u) + (upper & (w >> 17)) + (upper & (w >> 16));
        // C_MASK (default_tokenizer.py:80-91)
        const uint32_t M0 = (1u << 7) | (1u << 18) | (1u << 13);
        const uint32_t M1 = (1u << 11) | (1u << 18) | (1u << 21) | (1u << 23);
        const uint32_t M2 = (1u << 8) | (1u << 14) | (1u << 15);
        const uint32_t M3 = (1u << 9) | (1u << 22) | (1u << 24) | (1u << 12);
        mark = ((w & M0) == M0) | ((w & M1) == M1) | ((w & M2) == M2) | ((w & M3) == M3);
        // C_SYM: SYMBOL*NEXT_SPACE (default_tokenizer.py:100-102)
        sym = ((w >> 6) & (w >> 19)) & 1u;
    } else {
        cnt = 0; sym = 0; uint32_t mk = 0;
        for (int i = 0; i < c_rules.n_split; ++i) cnt += ((w & c_rules.split[i]) == c_rules.split[i]);
        for (int i = 0; i < c_rules.n_mask; ++i) mk += ((w & c_rules.mask[i]) == c_rules.mask[i]);
        for (int i = 0; i < c_rules.n_sym; ++i) sym += ((w & c_rules.sym[i]) == c_rules.sym[i]);
        mark = mk != 0;
    }
}

// 4 feature bits -> 4 byte counters (bit k -> byte k)
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// ---- chain combine operators (older `a`, newer `b`; a reset element discards everything older) ----
__device__ __forceinline__ Chain1 combine1(const Chain1 &a, const Chain1 &b)
{
    if (b.reset) return b;
    Chain1 r;
    r.reset = a.reset;
    r.n = a.n + b.n;
    r.has = a.has | b.has;
    r.lf = b.has ? a.n + b.lf : a.lf;
    Fn f = fn_compose(Fn{a.u, a.v}, Fn{b.u, b.v});
    r.u = f.u; r.v = f.v;
    return r;
}
__device__ __forceinline__ Chain2 combine2(const Chain2 &a, const Chain2 &b)
{
    if (b.reset) return b;
    Chain2 r;
    r.reset = a.reset;
    r.k = a.k + b.k;
    r.has_split = a.has_split | b.has_split;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.sums[i] = b.has_split ? b.sums[i] : __vadd4(a.sums[i], b.sums[i]);
    return r;
}
__device__ __forceinline__ Chain1 shfl_up_c(const Chain1 &e, int d)
{
    Chain1 o;
    o.n = __shfl_up_sync(0xFFFFFFFFu, e.n, d); o.lf = __shfl_up_sync(0xFFFFFFFFu, e.lf, d);
    o.u = __shfl_up_sync(0xFFFFFFFFu, e.u, d); o.v = __shfl_up_sync(0xFFFFFFFFu, e.v, d);
    o.has = __shfl_up_sync(0xFFFFFFFFu, e.has, d); o.reset = __shfl_up_sync(0xFFFFFFFFu, e.reset, d);
    return o;
}
__device__ __forceinline__ Chain2 shfl_up_c(const Chain2 &e, int d)
{
    Chain2 o;
    o.k = __shfl_up_sync(0xFFFFFFFFu, e.k, d);
    o.has_split = __shfl_up_sync(0xFFFFFFFFu, e.has_split, d); o.reset = __shfl_up_sync(0xFFFFFFFFu, e.reset, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) o.sums[i] = __shfl_up_sync(0xFFFFFFFFu, e.sums[i], d);
    return o;
}
__device__ __forceinline__ Chain1 shfl_c(const Chain1 &e, int src)
{
    Chain1 o;
    o.n = __shfl_sync(0xFFFFFFFFu, e.n, src); o.lf = __shfl_sync(0xFFFFFFFFu, e.lf, src);
    o.u = __shfl_sync(0xFFFFFFFFu, e.u, src); o.v = __shfl_sync(0xFFFFFFFFu, e.v, src);
    o.has = __shfl_sync(0xFFFFFFFFu, e.has, src); o.reset = __shfl_sync(0xFFFFFFFFu, e.reset, src);
    return o;
}
__device__ __forceinline__ Chain2 shfl_c(const Chain2 &e, int src)
{
    Chain2 o;
    o.k = __shfl_sync(0xFFFFFFFFu, e.k, src);
    o.has_split = __shfl_sync(0xFFFFFFFFu, e.has_split, src); o.reset = __shfl_sync(0xFFFFFFFFu, e.reset, src);
#pragma unroll
    for (int i = 0; i < 8; ++i) o.sums[i] = __shfl_sync(0xFFFFFFFFu, e.sums[i], src);
    return o;
}
__device__ __forceinline__ Chain1 origin1() { Chain1 o; o.n = 0; o.lf = 0; o.u = NEG; o.v = 0; o.has = 1; o.reset = 1; return o; }
__device__ __forceinline__ Chain2 origin2() { Chain2 o; o.k = 0; o.has_split = 1; o.reset = 1; for (int i = 0; i < 8; ++i) o.sums[i] = 0; return o; }
__device__ __forceinline__ Chain1 combine_c(const Chain1 &a, const Chain1 &b) { return combine1(a, b); }
__device__ __forceinline__ Chain2 combine_c(const Chain2 &a, const Chain2 &b) { return combine2(a, b); }
__device__ __forceinline__ void origin_c(Chain1 &o) { o = origin1(); }
__device__ __forceinline__ void origin_c(Chain2 &o) { o = origin2(); }

template <class T>
__device__ __forceinline__ T load_state(const T *p)
{
    // 16-byte L2 loads (never the non-coherent L1 path)
    T r;
    const int4 *s = reinterpret_cast<const int4 *>(p);
    int4 *d = reinterpret_cast<int4 *>(&r);
#pragma unroll
    for (int i = 0; i < int(sizeof(T) / 16); ++i) d[i] = __ldcg(s + i);
    return r;
}
template <class T>
__device__ __forceinline__ void store_state(T *p, const T &v)
{
    const int4 *s = reinterpret_cast<const int4 *>(&v);
    int4 *d = reinterpret_cast<int4 *>(p);
#pragma unroll
    for (int i = 0; i < int(sizeof(T) / 16); ++i) __stcg(d + i, s[i]);
}

// Publish this tile's aggregate (state 1) or inclusive prefix (state 2).  Called by one lane.
template <class T>
__device__ __forceinline__ void publish(T *slot, unsigned *status, const T &v, unsigned epoch, unsigned state)
{
    store_state(slot, v);
    __threadfence();
    st_volatile_u32(status, (epoch << 2) | state);
}

// Decoupled look-back over the preceding tiles, 32 at a time (warp 0 only).  Returns the
// exclusive prefix of `tile` as a reset element in every lane.
template <class T>
__device__ T lookback(long long tile, const unsigned *status, const T *agg, const T *inc, unsigned epoch,
                      Result *result, int lane)
{
    T acc;
    bool have = false;
    for (long long base = tile - 1;; base -= 32) {
        long long idx = base - (31 - lane);  // lane 31 looks at the nearest predecessor
        T e;
        if (idx < 0) {
            origin_c(e);
        } else {
            unsigned s, spins = 0;
            for (;;) {
                s = ld_volatile_u32(status + idx);
                if ((s >> 2) == epoch && (s & 3u) != 0u) break;
                if (++spins > SPIN_LIMIT || (((spins & 1023u) == 0u) && ld_volatile_u32(&result->abort_flag))) {
                    atomicOr(&result->error, 1u);
                    st_volatile_u32(&result->abort_flag, 1u);
                    s = 2u;  // give up: behave as if an (arbitrary) prefix was found so the kernel terminates
                    break;
                }
                __nanosleep(20);
            }
            __threadfence();
            const bool is_inc = (s & 3u) == 2u;
            e = load_state(is_inc ? inc + idx : agg + idx);
            e.reset = is_inc ? 1u : 0u;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = shfl_up_c(e, d);
            if (lane >= d) e = combine_c(o, e);
        }
        T win = shfl_c(e, 31);
        acc = have ? combine_c(win, acc) : win;
        have = true;
        if (acc.reset) break;
    }
    return acc;
}

// ---- look-ahead walk (rare): the whitespace chunk open at the end of a tile did not close inside
// the right halo.  Warp 0 scans forward from global byte `pos0` until the chunk closes (a SPACE
// character or the end of the string) and reports whether a mark occurs up to and including the
// closing character.  Exact but slow; it exists so that arbitrarily long space-free runs stay
// bit-exact (latok.c:218-244 has unbounded reach).
__device__ bool walk_ahead(const Params &p, const Tables &t, long long pos0, int lane)
{
    // end of the string that contains pos0: first offset > pos0
    long long e = p.n_bytes;
    if (lane == 0) {
        long long lo = 0, hi = p.n_strings;  // offsets[lo] <= pos0 < offsets[hi] invariant target
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (p.offsets[mid] > pos0) hi = mid; else lo = mid + 1;
        }
        e = p.offsets[lo <= p.n_strings ? lo : p.n_strings];
    }
    e = __shfl_sync(0xFFFFFFFFu, e, 0);
    const uint8_t *in = p.in;
    auto byte_at = [&](long long q) -> uint32_t { return (q >= 0 && q < p.n_bytes) ? (uint32_t)in[q] : 0u; };
    auto is_lead = [&](long long q) -> bool { return (byte_at(q) & 0xC0u) != 0x80u; };
    struct G { const uint8_t *in; long long q, n; __device__ uint32_t operator[](int k) const { long long a = q + k; return a < n ? (uint32_t)in[a] : 0u; } };
    bool any = false;
    for (long long q = pos0; q < e; q += 32) {
        long long pos = q + lane;
        bool lead = pos < e && is_lead(pos);
        bool closer = false, mk = false;
        if (lead) {
            uint32_t w = classify_at(G{in, pos, p.n_bytes}, t);
            // previous character (we are strictly inside the string, so it exists)
            long long pp = pos - 1;
            for (int k = 0; k < 8 && pp > 0 && !is_lead(pp); ++k) --pp;
            uint32_t pw = classify_at(G{in, pp, p.n_bytes}, t);
            long long n1 = pos + 1;
            for (int k = 0; k < 8 && n1 < e && !is_lead(n1); ++k) ++n1;
            bool has_next = n1 < e;
            uint32_t nw = has_next ? classify_at(G{in, n1, p.n_bytes}, t) : 0u;
            long long n2 = n1 + 1;
            for (int k = 0; k < 8 && n2 < e && !is_lead(n2); ++k) ++n2;
            bool has_an = has_next && n2 < e;
            uint32_t aw = has_an ? classify_at(G{in, n2, p.n_bytes}, t) : 0u;
            uint32_t full = make_word(pw, w, nw, aw, false, !has_next, !has_an);
            uint32_t cnt, sy;
            eval_rules(p.rules, full, cnt, mk, sy);
            closer = ((full >> 5) & 1u) || !has_next;
        }
        unsigned bc = __ballot_sync(0xFFFFFFFFu, closer), bm = __ballot_sync(0xFFFFFFFFu, lead && mk);
        if (bc) {
            int first = __ffs(bc) - 1;
            return any || (bm & (first == 31 ? 0xFFFFFFFFu : ((2u << first) - 1u))) != 0u;
        }
        any = any || bm != 0u;
        if (any) return true;
    }
    return any;
}

// =====================================================================================================
// tokenize_kernel (v2, bit-plane formulation)
//
// Every thread owns 32 window bytes and keeps its characters as 32-bit BIT-PLANES in registers: bit j of
// plane f = feature f of the thread's j-th character.  All per-character logic of the reference
// (context features, the three combo-matrix rules, the block mask, split values, token flags) then runs
// 32 characters per instruction.
// =====================================================================================================
enum { PL_A = 0, PL_N = 1, PL_NUM = 2, PL_LO = 3, PL_UP = 4, PL_SP = 5, PL_SY = 6, PL_TW = 7, PL_AT = 8, PL_CO = 9,
       PL_SL = 10, PL_PE = 11 };

struct SmemPlan {
    int mbar, scal, tile, table, spans, startbits, leadmask, cpref, emit, tokpref, split, edge, scratch, words, total;
};
__host__ __device__ inline SmemPlan smem_plan(int table_bytes, bool want_words)
{
    SmemPlan s; int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    s.mbar = take(16);
    s.scal = take(256);
    s.tile = take(WINB + 64);          // window bytes; re-used as the split-value staging buffer
    s.table = take(table_bytes);
    s.spans = take(SPAN_STAGE * 8);
    s.startbits = take(NT * 4);
    s.leadmask = take(NT * 4);
    s.cpref = take((NT + 1) * 4);
    s.emit = take(NT * 4);
    s.tokpref = take((NT + 1) * 4);
    s.split = take(NT * 4);
    s.edge = take(NWARP * 16 * 4);
    s.scratch = take(1024);
    s.words = take(want_words ? (WINB + WINB / 32 + 64) * 4 : 0);
    s.total = o;
    return s;
}
size_t tokenize_smem_bytes(const TableLayout &tl, bool want_words) { return (size_t)smem_plan(tl.total, want_words).total; }

struct Scalars {          // block-shared scalars
    long long tile;
    unsigned long long G_in, base_in, K_in;
    int x_in, x_end;
    int agg_u, agg_v;
    int need_walk, far, slow_vals;
    int lf_tile;
    unsigned open_has;
    unsigned open_sums[8];
    unsigned carry_sums[8];
};

__device__ __forceinline__ int block_excl_sum(int v, int *scratch, int &total, int lane, int warp)
{
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) { int t = scratch[w]; if (w < warp) base += t; tot += t; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

// exclusive scan of backlog functions (composition)
__device__ __forceinline__ void block_excl_fn(Fn f, int *scratch, Fn &excl, Fn &total, int lane, int warp)
{
    Fn inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int ou = __shfl_up_sync(0xFFFFFFFFu, inc.u, d), ov = __shfl_up_sync(0xFFFFFFFFu, inc.v, d);
        if (lane >= d) inc = fn_compose(Fn{ou, ov}, inc);
    }
    int eu = __shfl_up_sync(0xFFFFFFFFu, inc.u, 1), ev = __shfl_up_sync(0xFFFFFFFFu, inc.v, 1);
    Fn wex = lane ? Fn{eu, ev} : fn_id();
    if (lane == 31) { scratch[2 * warp] = inc.u; scratch[2 * warp + 1] = inc.v; }
    __syncthreads();
    Fn base = fn_id(), tot = fn_id();
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
        Fn t = Fn{scratch[2 * w], scratch[2 * w + 1]};
        if (w < warp) base = fn_compose(base, t);
        tot = fn_compose(tot, t);
    }
    __syncthreads();
    excl = fn_compose(base, wex);
    total = tot;
}

// LUT entry (256 + class) of the multi-byte character whose lead byte is p[0] >= 0xC0
__device__ __forceinline__ uint32_t mb_entry(const uint8_t *p, const Tables &t, uint32_t high_class)
{
    const uint32_t b0 = p[0];
    if (b0 >= 0xF8u) return 256u;  // invalid lead: class 0 (no features)
    uint32_t cp;
    if (b0 < 0xE0u) cp = ((b0 & 0x1Fu) << 6) | (p[1] & 0x3Fu);
    else if (b0 < 0xF0u) cp = ((b0 & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu);
    else cp = ((b0 & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu);
    if (cp < 0x80u) return cp;     // over-long form of an ASCII character
    if (cp < t.low_limit) {
        const uint32_t blk = t.stage1[cp >> 7];
        const uint32_t b = t.stage2[blk * 64u + ((cp & 127u) >> 1)];
        return 256u + ((cp & 1u) ? (b >> 4) : (b & 15u));
    }
    return (cp >= t.high_first && cp <= t.high_last) ? 256u + high_class : 256u;
}

// 4x4 byte transpose: four accumulators (8 characters each, one byte per feature) -> four 32-character planes
__device__ __forceinline__ void planes4(const uint32_t a[4], uint32_t &p0, uint32_t &p1, uint32_t &p2, uint32_t &p3)
{
    const uint32_t t0 = __byte_perm(a[0], a[1], 0x5140), t1 = __byte_perm(a[0], a[1], 0x7362);
    const uint32_t t2 = __byte_perm(a[2], a[3], 0x5140), t3 = __byte_perm(a[2], a[3], 0x7362);
    p0 = __byte_perm(t0, t2, 0x5410); p1 = __byte_perm(t0, t2, 0x7632);
    p2 = __byte_perm(t1, t3, 0x5410); p3 = __byte_perm(t1, t3, 0x7632);
}

// In-register 32x32 bit-matrix transpose: planes (bit j of a[f]) -> per-character words (bit f of a[j])
__device__ __forceinline__ void transpose32(uint32_t a[32])
{
#pragma unroll
    for (int j = 16, sh = 0; j != 0; j >>= 1, ++sh) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
                a[k] ^= t << j;
                a[k + j] ^= t;
            }
        }
    }
}

template <bool kDefault, bool kWords>
__global__ void __launch_bounds__(NT) tokenize_kernel(const Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemPlan sp = smem_plan(p.tl.total, kWords);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + sp.mbar);
    Scalars &sc = *reinterpret_cast<Scalars *>(smem + sp.scal);
    uint8_t *tileS = smem + sp.tile;
    uint8_t *valS = smem + sp.tile;      // aliases the window bytes (dead after phase 1)
    uint8_t *tableS = smem + sp.table;
    int32_t *spanS = reinterpret_cast<int32_t *>(smem + sp.spans);
    uint32_t *startbits = reinterpret_cast<uint32_t *>(smem + sp.startbits);
    uint32_t *leadmaskS = reinterpret_cast<uint32_t *>(smem + sp.leadmask);
    int *cprefS = reinterpret_cast<int *>(smem + sp.cpref);
    uint32_t *emitS = reinterpret_cast<uint32_t *>(smem + sp.emit);
    int *tokprefS = reinterpret_cast<int *>(smem + sp.tokpref);
    uint32_t *splitS = reinterpret_cast<uint32_t *>(smem + sp.split);
    uint32_t *edgeS = reinterpret_cast<uint32_t *>(smem + sp.edge);
    int *scratch = reinterpret_cast<int *>(smem + sp.scratch);
    uint32_t *wordS = reinterpret_cast<uint32_t *>(smem + sp.words);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xFFFFFFFFu;

    if (ld_volatile_u32(&p.result->error) & 2u) return;  // offsets failed validation in tile_index_kernel

    // one-time per CTA: tables into shared memory, mbarrier init
    for (int i = tid; i < p.tl.total / 16; i += NT)
        reinterpret_cast<uint4 *>(tableS)[i] = __ldg(reinterpret_cast<const uint4 *>(p.table_blob) + i);
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    Tables tb;
    tb.ascii_feat = reinterpret_cast<const uint16_t *>(tableS + p.tl.ascii_feat);
    tb.class_feat = reinterpret_cast<const uint16_t *>(tableS + p.tl.class_feat);
    tb.stage1 = tableS + p.tl.stage1;
    tb.stage2 = tableS + p.tl.stage2;
    tb.low_limit = p.tl.low_limit; tb.high_first = p.tl.high_first; tb.high_last = p.tl.high_last; tb.high_feat = p.tl.high_feat;
    const uint32_t *lut0 = reinterpret_cast<const uint32_t *>(tableS + p.tl.lut3);
    const uint32_t *lut1 = lut0 + LUT_ENTRIES, *lut2 = lut1 + LUT_ENTRIES;
    const uint32_t *lutv = reinterpret_cast<const uint32_t *>(tableS + p.tl.lutv);

    uint32_t phase = 0;
    const bool want_feats = (p.what & 4u) != 0u, want_matrix = (p.what & 8u) != 0u;
    const bool want_spans = (p.what & 2u) != 0u, want_splits = (p.what & 1u) != 0u;

    for (;;) {
        if (tid == 0) sc.tile = (long long)(atomicAdd(p.ticket, 1ull) - p.ticket_base);
        __syncthreads();  // (A) also fences shared-memory reuse across tiles
        const long long tile = sc.tile;
        if (tile >= p.ntiles) break;

        // ------------------------------------------------------------------ load window (TMA bulk copy)
        const long long w0 = tile * (long long)TILE - LHALO;
        const long long lo = w0 < 0 ? 0 : w0;
        long long hi = w0 + WINB;
        const long long full16 = p.n_bytes & ~15LL;
        if (hi > full16) hi = full16;
        const int tma_bytes = hi > lo ? int(hi - lo) : 0;
        if (tid == 0 && tma_bytes > 0) {
            fence_proxy_async();
            mbar_expect_tx(mbar, (uint32_t)tma_bytes);
            tma_load_1d(tileS + (lo - w0), p.in + lo, (uint32_t)tma_bytes, mbar);
        }
        {
            const int a_end = int(lo - w0);
            const int b_beg = a_end + tma_bytes;
            for (int i = tid; i < a_end; i += NT) tileS[i] = 0;
            for (int i = b_beg + tid; i < WINB + 16; i += NT) {
                long long g = w0 + i;
                tileS[i] = (g >= 0 && g < p.n_bytes) ? p.in[g] : (uint8_t)0;
            }
        }
        startbits[tid] = 0;
        if (tid == 0) { sc.need_walk = 0; sc.far = 0; sc.open_has = 0; sc.slow_vals = 0; }
        __syncthreads();  // (B)
        {
            const long long wend = w0 + WINB;
            for (long long s = p.tile_first_str[tile] + tid; s <= p.n_strings; s += NT) {
                long long o = p.offsets[s];
                if (o >= wend) break;
                int wb = int(o - w0);
                atomicOr(&startbits[wb >> 5], 1u << (wb & 31));
            }
        }
        if (tma_bytes > 0) { mbar_wait(mbar, phase); phase ^= 1u; }
        __syncthreads();  // (C)

        // ------------------------------------------------------------------ phase 1: bytes -> bit-planes
        const int wb0 = tid * 32;
        const uint32_t sb = startbits[tid];
        uint32_t lead, mbl;          // lead bytes / lead bytes of multi-byte characters (byte positions)
        int vhi;                     // number of valid bytes at the low end of this thread's range
        uint32_t acc0[4], acc1[4], acc2[4];
        {
            const uint4 q0 = *reinterpret_cast<const uint4 *>(tileS + wb0);
            const uint4 q1 = *reinterpret_cast<const uint4 *>(tileS + wb0 + 16);
            const uint32_t wds[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t leadbits = 0, hib = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t t = (wds[j] & 0xC0C0C0C0u) ^ 0x80808080u;   // byte == 0  <=>  continuation byte
                uint32_t m = (t | (t << 1)) & 0x80808080u;
                leadbits |= (((m >> 7) * 0x10204080u) >> 28) << (4 * j);
                hib |= ((((wds[j] & 0x80808080u) >> 7) * 0x10204080u) >> 28) << (4 * j);
            }
            const long long g0 = w0 + wb0;
            const int vlo = g0 < 0 ? int(-g0 < 32 ? -g0 : 32) : 0;
            const long long rem = p.n_bytes - g0;
            vhi = rem <= 0 ? 0 : (rem >= 32 ? 32 : int(rem));
            const uint32_t valid = vhi > vlo ? (mask_lt(vhi) & ~mask_lt(vlo)) : 0u;
            lead = (leadbits & valid) | sb;
            mbl = leadbits & hib & valid;     // bytes >= 0xC0 inside the data
            // every byte through the feature LUT (bytes >= 0x80 and padding map to "no features")
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t b = (wds[2 * g + (j >> 2)] >> ((j & 3) * 8)) & 0xFFu;
                    a0 += lut0[b] << j; a1 += lut1[b] << j; a2 += lut2[b] << j;
                }
                acc0[g] = a0; acc1[g] = a1; acc2[g] = a2;
            }
        }
        // patch in the multi-byte characters (decode + two-stage class table)
        {
            uint32_t mm = mbl;
            while (mm) {
                const int k = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t e = mb_entry(tileS + wb0 + k, tb, p.tl.high_class);
                const uint32_t x0 = lut0[e] << (k & 7), x1 = lut1[e] << (k & 7), x2 = lut2[e] << (k & 7);
                const int g = k >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (g == q) { acc0[q] |= x0; acc1[q] |= x1; acc2[q] |= x2; }
            }
        }
        uint32_t P[12];
        planes4(acc0, P[0], P[1], P[2], P[3]);
        planes4(acc1, P[4], P[5], P[6], P[7]);
        planes4(acc2, P[8], P[9], P[10], P[11]);
        uint32_t Fm = sb;
        // byte space -> character space: squeeze out the continuation-byte positions
        {
            uint32_t del = lead ? (~lead & mask_lt(vhi)) : 0u;
#pragma unroll
            for (int f = 0; f < 12; ++f) P[f] &= lead;
            while (del) {
                const int top = 31 - __clz(del);
                const int r = __clz(~(del << (31 - top)));       // run length of deleted bytes ending at `top`
                const int c = top - r;                            // last kept position below the run (may be -1)
                const uint32_t keep = mask_lt(c + 1);
#pragma unroll
                for (int f = 0; f < 12; ++f) P[f] = (P[f] & keep) | ((P[f] >> r) & ~keep);
                Fm = (Fm & keep) | ((Fm >> r) & ~keep);
                del &= keep;
            }
        }
        const int n = __popc(lead);                     // characters of this thread (incl. the end-of-data terminator)
        int c_end;
        const int c0 = block_excl_sum(n, scratch, c_end, lane, warp);
        leadmaskS[tid] = lead;
        cprefS[tid] = c0;
        if (tid == 0) cprefS[NT] = c_end;

        // ------------------------------------------------------------------ phase 2a: context planes
        // prev: last character of the previous thread; next/after-next: first two characters of the next thread
        const uint32_t myLB = n > 0 ? ((((P[PL_A] >> (n - 1)) & 1u)) | (((P[PL_N] >> (n - 1)) & 1u) << 1) |
                                       (((P[PL_LO] >> (n - 1)) & 1u) << 2) | (((P[PL_SP] >> (n - 1)) & 1u) << 3) |
                                       (((P[PL_SY] >> (n - 1)) & 1u) << 4) | 32u)
                                    : 0u;
        uint32_t LB = __shfl_up_sync(FULL, myLB, 1);
        uint32_t XA = __shfl_down_sync(FULL, P[PL_A], 1), XN = __shfl_down_sync(FULL, P[PL_N], 1);
        uint32_t XLO = __shfl_down_sync(FULL, P[PL_LO], 1), XSP = __shfl_down_sync(FULL, P[PL_SP], 1);
        uint32_t XAT = __shfl_down_sync(FULL, P[PL_AT], 1), XSL = __shfl_down_sync(FULL, P[PL_SL], 1);
        uint32_t XF = __shfl_down_sync(FULL, Fm, 1);
        if (lane == 0) {
            uint32_t *e = edgeS + warp * 16;
            e[0] = P[PL_A]; e[1] = P[PL_N]; e[2] = P[PL_LO]; e[3] = P[PL_SP]; e[4] = P[PL_AT]; e[5] = P[PL_SL]; e[6] = Fm;
        }
        if (lane == 31) edgeS[warp * 16 + 8] = myLB;
        __syncthreads();  // (D) edges + cpref/leadmask visible
        if (lane == 31) {
            if (warp + 1 < NWARP) {
                const uint32_t *e = edgeS + (warp + 1) * 16;
                XA = e[0]; XN = e[1]; XLO = e[2]; XSP = e[3]; XAT = e[4]; XSL = e[5]; XF = e[6];
            } else { XA = XN = XLO = XSP = XAT = XSL = XF = 0; }
        }
        if (lane == 0) LB = warp > 0 ? edgeS[(warp - 1) * 16 + 8] : 0u;

        auto cidx = [&](int wb) -> int {  // characters starting at window bytes < wb
            int t = wb >> 5;
            if (t >= NT) return cprefS[NT];
            return cprefS[t] + __popc(leadmaskS[t] & mask_lt(wb & 31));
        };
        const int c_lo = cprefS[FIRST_OWNED_THREAD];
        long long own_end_g = (tile + 1) * (long long)TILE;
        if (own_end_g > p.n_bytes) own_end_g = p.n_bytes;
        const int c_hi = cidx(int(own_end_g - w0));
        const bool term_in_win = p.n_bytes < w0 + WINB;
        const int n_own = c_hi - c_lo;
        // thread-local masks: real characters (not the terminator), owned, active (trusted forward context)
        const long long g0 = w0 + wb0;
        const bool has_term = p.n_bytes >= g0 && p.n_bytes < g0 + 32;
        const uint32_t REAL = mask_lt(n - (has_term ? 1 : 0));
        const uint32_t OWN = (tid >= FIRST_OWNED_THREAD && tid < END_OWNED_THREAD) ? REAL : 0u;
        uint32_t ACT = tid >= FIRST_OWNED_THREAD ? REAL : 0u;
        if (tid == NT - 1 && !term_in_win) ACT &= mask_lt(__popc(lead & mask_lt(32 - TRUST_MARGIN)));

        // 64-bit view "this thread's characters followed by the next thread's": shift right by 1 / 2
        auto next1 = [&](uint32_t X, uint32_t Xn) -> uint32_t {
            const uint32_t l = X | __funnelshift_lc(0u, Xn, n), h = __funnelshift_lc(Xn, 0u, n);
            return __funnelshift_r(l, h, 1);
        };
        auto next2 = [&](uint32_t X, uint32_t Xn) -> uint32_t {
            const uint32_t l = X | __funnelshift_lc(0u, Xn, n), h = __funnelshift_lc(Xn, 0u, n);
            return __funnelshift_r(l, h, 2);
        };
        const uint32_t Lm_raw = next1(Fm, XF);           // character ends a string (next one starts a string)
        const uint32_t L2m = next2(Fm, XF);
        const uint32_t nF = ~Lm_raw, aF = ~(Lm_raw | L2m), pF = ~Fm;
        uint32_t full[25];
#pragma unroll
        for (int f = 0; f < 12; ++f) full[f] = P[f];
        full[12] = ((P[PL_A] << 1) | (LB & 1u)) & pF;                 // PREV_ALPHA
        full[13] = next1(P[PL_A], XA) & nF;                           // NEXT_ALPHA
        full[14] = ((P[PL_N] << 1) | ((LB >> 1) & 1u)) & pF;          // PREV_ALPHA_NUM
        full[15] = next1(P[PL_N], XN) & nF;                           // NEXT_ALPHA_NUM
        full[16] = ((P[PL_LO] << 1) | ((LB >> 2) & 1u)) & pF;         // PREV_LOWER
        full[17] = next1(P[PL_LO], XLO) & nF;                         // NEXT_LOWER
        full[18] = ((P[PL_SP] << 1) | ((LB >> 3) & 1u)) | Fm;         // PREV_SPACE  (start of string = space)
        full[19] = next1(P[PL_SP], XSP) | Lm_raw;                     // NEXT_SPACE  (end of string = space)
        full[20] = ((P[PL_SY] << 1) | ((LB >> 4) & 1u)) & pF;         // PREV_SYMBOL
        full[21] = next1(P[PL_AT], XAT) & nF;                         // NEXT_AT
        full[22] = next1(P[PL_SL], XSL) & nF;                         // NEXT_SLASH
        full[23] = next2(P[PL_A], XA) & aF;                           // AFTER_NEXT_ALPHA
        full[24] = next2(P[PL_SL], XSL) & aF;                         // AFTER_NEXT_SLASH

        // ---- rules: split count, mark, sym (combine_matrix_rows 2-D, latok.c:318-341) as bit-sliced counters
        uint32_t CNT[4], SYC[4], Mm;
        if (kDefault) {
            // C_SPLIT: SPACE + SYMBOL + PREV_SYMBOL + UPPER*NEXT_LOWER + UPPER*PREV_LOWER (default_tokenizer.py:49-55)
            const uint32_t t1 = full[5], t2 = full[6], t3 = full[20], t4 = full[4] & full[17], t5 = full[4] & full[16];
            const uint32_t s1 = t1 ^ t2 ^ t3, c1 = (t1 & t2) | (t3 & (t1 ^ t2));
            const uint32_t s2 = t4 ^ t5, c2 = t4 & t5;
            CNT[0] = s1 ^ s2;
            const uint32_t c3 = s1 & s2;
            CNT[1] = c1 ^ c2 ^ c3;
            CNT[2] = (c1 & c2) | (c3 & (c1 ^ c2));
            CNT[3] = 0;
            // C_MASK (default_tokenizer.py:80-91)
            Mm = (full[7] & full[18] & full[13]) | (full[11] & full[18] & full[21] & full[23]) |
                 (full[8] & full[14] & full[15]) | (full[9] & full[22] & full[24] & full[12]);
            // C_SYM: SYMBOL*NEXT_SPACE (default_tokenizer.py:100-102)
            SYC[0] = full[6] & full[19]; SYC[1] = SYC[2] = SYC[3] = 0;
        } else {
            auto term = [&](uint32_t mask) -> uint32_t {
                uint32_t a = 0xFFFFFFFFu;
#pragma unroll
                for (int f = 0; f < NFEAT; ++f) a &= full[f] | (((mask >> f) & 1u) - 1u);
                return a;
            };
            auto add1 = [&](uint32_t c[4], uint32_t t) {
#pragma unroll
                for (int b = 0; b < 4; ++b) { const uint32_t k = c[b] & t; c[b] ^= t; t = k; }
            };
#pragma unroll
            for (int b = 0; b < 4; ++b) { CNT[b] = 0; SYC[b] = 0; }
            Mm = 0;
            for (int i = 0; i < p.rules.n_split; ++i) add1(CNT, term(p.rules.split[i]));
            for (int i = 0; i < p.rules.n_mask; ++i) Mm |= term(p.rules.mask[i]);
            for (int i = 0; i < p.rules.n_sym; ++i) add1(SYC, term(p.rules.sym[i]));
        }
        const uint32_t Sraw = P[PL_SP];
        const uint32_t S = Sraw & ACT;
        Mm &= ACT;
        const uint32_t FmA = Fm & ACT, Lm = Lm_raw & ACT;

        if (kWords) {
            // per-character 25-bit words (+ FIRST / LAST flags) for the token-feature and matrix emitters
            uint32_t a[32];
#pragma unroll
            for (int f = 0; f < 25; ++f) a[f] = full[f];
            a[25] = Fm; a[26] = Lm_raw;
#pragma unroll
            for (int f = 27; f < 32; ++f) a[f] = 0;
            transpose32(a);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < n) wordS[widx(c0 + j)] = a[j];
        }

        // local backlog function: +1 per mark, max(x-1,0) per space, reset at a string start
        Fn f_act;
        {
            const uint32_t ev0 = Mm | FmA;
            if (ev0 == 0u) { f_act.u = -__popc(S); f_act.v = 0; }
            else {
                f_act = fn_id();
                uint32_t ev = ev0 | S;
                while (ev) {
                    const uint32_t b = ev & (0u - ev); ev &= ev - 1;
                    if (FmA & b) { f_act.u = NEG; f_act.v = 0; }
                    if (Mm & b) { f_act.u = max(f_act.u + 1, NEG); f_act.v = f_act.v + 1; }
                    if (S & b) { f_act.u = max(f_act.u - 1, NEG); f_act.v = max(f_act.v - 1, 0); }
                }
            }
        }
        Fn excl, total;
        block_excl_fn(f_act, scratch, excl, total, lane, warp);
        if (tid == END_OWNED_THREAD) { sc.agg_u = excl.u; sc.agg_v = excl.v; }  // composition over the owned threads
        // last string start among the owned characters (tile-relative), and for each thread the latest one before it
        int lf_excl;
        {
            const uint32_t FO = Fm & OWN;
            const int mine = FO ? c0 + 31 - __clz(FO) : -1;
            const unsigned has = __ballot_sync(FULL, FO != 0u);
            const unsigned below = has & mask_lt(lane);
            const int src = below ? 31 - __clz(below) : 0;
            const int got = __shfl_sync(FULL, mine, src);
            lf_excl = below ? got : -1;
            const int wlast = __shfl_sync(FULL, mine, has ? 31 - __clz(has) : 0);
            if (lane == 0) scratch[128 + warp] = has ? wlast : -1;
        }
        __syncthreads();  // (E)
        {
            int lfw = -1, lft = -1;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) { const int v = scratch[128 + w]; if (w < warp) lfw = max(lfw, v); lft = max(lft, v); }
            lf_excl = max(lf_excl, lfw);
            if (tid == 0) sc.lf_tile = lft;
        }

        // ------------------------------------------------------------------ chain 1
        if (warp == 0) {
            int lft = -1;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) lft = max(lft, scratch[128 + w]);
            Chain1 a;
            a.n = (unsigned long long)n_own;
            a.has = lft >= 0 ? 1u : 0u;
            a.lf = a.has ? (unsigned long long)(lft - c_lo) : 0ull;
            a.u = sc.agg_u; a.v = sc.agg_v; a.reset = 0;
            if (lane == 0) publish(p.agg1 + tile, p.status1 + tile, a, p.epoch, 1u);
            Chain1 pre = lookback<Chain1>(tile, p.status1, p.agg1, p.inc1, p.epoch, p.result, lane);
            if (lane == 0) {
                Chain1 inc = combine1(pre, a);
                inc.reset = 1;
                publish(p.inc1 + tile, p.status1 + tile, inc, p.epoch, 2u);
                sc.G_in = pre.n; sc.base_in = pre.lf; sc.x_in = pre.v;
                sc.x_end = fn_apply(total, pre.v);
            }
        }
        __syncthreads();  // (F)
        const unsigned long long G_in = sc.G_in;
        const int x_t = fn_apply(excl, sc.x_in);

        // ------------------------------------------------------------------ phase 2b: block mask
        const uint32_t CL = S | Lm;        // characters that close a whitespace chunk
        uint32_t HOT = 0;                  // closers whose chunk is blanked (backlog >= 1 at the closer)
        if (x_t != 0 || Mm != 0u) {
            int x = x_t; uint32_t ev = Mm | S | FmA | Lm;
            while (ev) {
                const uint32_t b = ev & (0u - ev); ev &= ev - 1;
                if (FmA & b) x = 0;
                if (Mm & b) ++x;
                if ((CL & b) && x >= 1) HOT |= b;
                if (S & b) x = max(x - 1, 0);
            }
        }
        uint32_t Zm = HOT;
        {
            uint32_t pr = ~CL;
            Zm |= pr & (Zm >> 1); pr &= pr >> 1;
            Zm |= pr & (Zm >> 2); pr &= pr >> 2;
            Zm |= pr & (Zm >> 4); pr &= pr >> 4;
            Zm |= pr & (Zm >> 8); pr &= pr >> 8;
            Zm |= pr & (Zm >> 16);
        }
        const bool hasCL = CL != 0u;
        const bool firstHot = hasCL && (HOT & (CL & (0u - CL))) != 0u;
        int cin;
        {
            const unsigned H = __ballot_sync(FULL, hasCL), FH = __ballot_sync(FULL, firstHot);
            if (lane == 0) { scratch[64 + 2 * warp] = H != 0u; scratch[64 + 2 * warp + 1] = H ? ((FH >> (__ffs(H) - 1)) & 1u) : 0u; }
            __syncthreads();  // (G)
            const unsigned above = lane == 31 ? 0u : (H & (0xFFFFFFFFu << (lane + 1)));
            if (above) cin = (FH >> (__ffs(above) - 1)) & 1u;
            else {
                cin = 2;
                for (int w2 = warp + 1; w2 < NWARP; ++w2)
                    if (scratch[64 + 2 * w2]) { cin = scratch[64 + 2 * w2 + 1]; break; }
            }
        }
        // the whitespace chunk open at the end of the owned range: closed inside the halo, or walk ahead
        if (tid == END_OWNED_THREAD - 1 && !term_in_win) {
            const int nr = __popc(ACT);
            const bool open = nr > 0 && ((CL >> (nr - 1)) & 1u) == 0u && cin == 2;
            if (open) sc.need_walk = 1;
        }
        __syncthreads();  // (H)
        if (sc.need_walk) {
            if (sc.x_end >= 1) { if (tid == 0) sc.far = 1; }
            else if (warp == 0) {
                bool any = walk_ahead(p, tb, w0 + WINB - TRUST_MARGIN, lane);
                if (lane == 0) { sc.far = any ? 1 : 0; atomicAdd(&p.result->walks, 1ull); }
            }
            __syncthreads();
        }
        if (cin == 1 || (cin == 2 && sc.far)) {
            const uint32_t top = hasCL ? ~((2u << (31 - __clz(CL))) - 1u) : 0xFFFFFFFFu;
            Zm |= top;
        }

        // ------------------------------------------------------------------ phase 3: split values, token flags
        // splits = split_cnt * block_mask + sym; splits[0] = 1   (default_tokenizer.py:121-132), bit-sliced
        const uint32_t keepm = ~Zm | Sraw;             // block mask = 1
        uint32_t V[5];
        {
            uint32_t carry = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t x = CNT[b] & keepm, y = SYC[b];
                V[b] = x ^ y ^ carry;
                carry = (x & y) | (carry & (x ^ y));
            }
            V[4] = carry;
            V[0] |= Fm;
#pragma unroll
            for (int b = 1; b < 5; ++b) V[b] &= ~Fm;
        }
        const uint32_t SPLIT = V[0] | V[1] | V[2] | V[3] | V[4];
        const uint32_t PS = (Sraw << 1) | ((LB >> 3) & 1u);                // previous character is a space
        const uint32_t E = ((SPLIT & ~Sraw) | (~SPLIT & PS & ~Fm)) & OWN;  // a token is counted at this character
        const uint32_t EW = SPLIT & ~Fm & ~PS & OWN;                       // this split ends the previous token
        const uint32_t EL = Lm & ~Sraw & OWN;                              // end of string ends the last token
        int ntok_tile;
        const int tp = block_excl_sum(__popc(E), scratch, ntok_tile, lane, warp);
        emitS[tid] = E; tokprefS[tid] = tp; splitS[tid] = SPLIT;
        if (tid == 0) tokprefS[NT] = ntok_tile;

        // split values -> bytes -> staging buffer at (tile-relative character index + G_in % 16), so that the
        // staging buffer and the global split mask share their 16-byte alignment
        const int a16 = int(G_in & 15ull);
        if (want_splits) {
            uint32_t W[8];
            if (kDefault) {
                const uint32_t qlo = (V[0] & 0x0F0F0F0Fu) | ((V[1] & 0x0F0F0F0Fu) << 4);
                const uint32_t qhi = ((V[0] >> 4) & 0x0F0F0F0Fu) | (V[1] & 0xF0F0F0F0u);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    W[2 * g] = lutv[(qlo >> (8 * g)) & 0xFFu];
                    W[2 * g + 1] = lutv[(qhi >> (8 * g)) & 0xFFu];
                }
                if (V[2]) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) W[g] += spread4((V[2] >> (4 * g)) & 15u) << 2;
                }
            } else {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t w = 0;
#pragma unroll
                    for (int b = 0; b < 5; ++b) w += spread4((V[b] >> (4 * g)) & 15u) << b;
                    W[g] = w;
                }
            }
            // the 4 bytes before this thread's first character (tail of the previous thread)
            uint32_t tailw = 0;
            {
                const int nr = n;   // terminator value bytes are harmless: they are never copied out
                if (nr >= 4) {
                    const int sft = nr - 4;
                    uint32_t w = 0;
#pragma unroll
                    for (int b = 0; b < 5; ++b) w += spread4((V[b] >> sft) & 15u) << b;
                    tailw = w;
                }
            }
            uint32_t headw = __shfl_up_sync(FULL, tailw, 1);
            if (lane == 31) edgeS[warp * 16 + 9] = tailw;
            const bool shortn = (n < 4) && (tid >= FIRST_OWNED_THREAD - 1) && (tid <= END_OWNED_THREAD) && g0 < p.n_bytes && g0 + 32 > 0;
            const int slow = __syncthreads_or(shortn ? 1 : 0);  // (I0) also publishes edgeS[..+9]
            if (lane == 0) headw = warp > 0 ? edgeS[(warp - 1) * 16 + 9] : 0u;
            const int o = c0 - c_lo + a16;                 // staging offset of this thread's first character
            if (tid >= FIRST_OWNED_THREAD && tid <= END_OWNED_THREAD && o >= 0) {
                if (!slow) {
                    // words are written by the thread that owns their LAST byte: no partial words, no races
                    const int s = o & 3;
                    const int cnt = (s + n) >> 2;
                    uint32_t *dst = reinterpret_cast<uint32_t *>(valS) + (o >> 2);
                    const int sh = 32 - 8 * s;             // stream = [headw, W0..W7]; word i starts s bytes before W[i-1]'s end
                    uint32_t prev = headw;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t w = s ? __funnelshift_r(prev, W[i], sh) : W[i];
                        if (i < cnt) dst[i] = w;
                        prev = W[i];
                    }
                    if (8 < cnt) dst[8] = __funnelshift_r(prev, 0u, sh);
                    // the thread that holds the end-of-data terminator flushes the trailing partial word byte by byte
                    // (bytes before this thread's first character come from the previous thread's tail word)
                    if (has_term) {
                        const int end = o + n - 1;               // staging position just after the last real character
                        for (int q = end & ~3; q < end; ++q) {
                            const int j = q - o;
                            const uint32_t src = j >= 0 ? W[j >> 2] >> ((j & 3) * 8) : headw >> ((4 + j) * 8);
                            if (q >= 0) valS[q] = (uint8_t)(src & 0xFFu);
                        }
                    }
                } else {
                    for (int j = 0; j < n; ++j) {
                        uint32_t v = 0;
#pragma unroll
                        for (int b = 0; b < 5; ++b) v |= ((V[b] >> j) & 1u) << b;
                        valS[o + j] = (uint8_t)v;
                    }
                }
            }
        }
        __syncthreads();  // (I)

        auto is_split = [&](int c) -> bool {   // c = tile character index; thread-local planes live in splitS
            // find the thread that holds character c: threads hold at least 8 characters each in valid UTF-8
            int t = c >> 5;
            while (t + 1 < NT && cprefS[t + 1] <= c) ++t;
            return (splitS[t] >> (c - cprefS[t])) & 1u;
        };
        // feature sums of characters c, c-1, ... down to the token's first character (a split) or c_lo
        auto walk_back = [&](int c, unsigned acc[7], bool &hit) {
            hit = false;
            for (; c >= c_lo; --c) {
                const uint32_t w = wordS[widx(c)] & FEATMASK;
#pragma unroll
                for (int g = 0; g < 7; ++g) acc[g] = __vadd4(acc[g], spread4((w >> (4 * g)) & 15u));
                if (is_split(c)) { hit = true; break; }
            }
        };
        if (kWords && want_feats) {
            if (n_own > 0 && tid == END_OWNED_THREAD) {
                unsigned acc[7] = {0, 0, 0, 0, 0, 0, 0}; bool hit;
                walk_back(c_hi - 1, acc, hit);
                sc.open_has = hit ? 1u : 0u;
                for (int g = 0; g < 7; ++g) sc.open_sums[g] = acc[g];
                sc.open_sums[7] = 0;
            }
            __syncthreads();
        }

        // ------------------------------------------------------------------ chain 2
        if (warp == 0) {
            Chain2 a;
            a.k = (unsigned long long)ntok_tile; a.reset = 0;
            a.has_split = (kWords && want_feats) ? sc.open_has : 0u;
            for (int g = 0; g < 8; ++g) a.sums[g] = (kWords && want_feats && n_own > 0) ? sc.open_sums[g] : 0u;
            if (lane == 0) publish(p.agg2 + tile, p.status2 + tile, a, p.epoch, 1u);
            Chain2 pre = lookback<Chain2>(tile, p.status2, p.agg2, p.inc2, p.epoch, p.result, lane);
            if (lane == 0) {
                Chain2 inc = combine2(pre, a);
                inc.reset = 1;
                publish(p.inc2 + tile, p.status2 + tile, inc, p.epoch, 2u);
                sc.K_in = pre.k;
                for (int g = 0; g < 8; ++g) sc.carry_sums[g] = pre.sums[g];
            }
        }
        // split mask out while warp 0 looks back: 16-byte chunks, byte-wise at the two ragged ends
        if (want_splits && n_own > 0) {
            int8_t *dst = p.splits + G_in;                  // dst[j] <-> valS[a16 + j]
            const int head = (16 - a16) & 15;               // bytes before the first 16-byte boundary
            const int nh = head < n_own ? head : n_own;
            if (tid < nh) dst[tid] = (int8_t)valS[a16 + tid];
            const int nchunks = (n_own - nh) >> 4;
            const uint4 *src = reinterpret_cast<const uint4 *>(valS + a16 + nh);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst + nh);
            for (int i = tid; i < nchunks; i += NT) d4[i] = src[i];
            const int done = nh + (nchunks << 4);
            if (tid < n_own - done) dst[done + tid] = (int8_t)valS[a16 + done + tid];
        }
        if (tid == NT - 1 && ntok_tile >= 1 && ntok_tile <= SPAN_STAGE - 1) spanS[2 * (ntok_tile - 1) + 3] = -1;  // "still open"
        __syncthreads();  // (J)
        const unsigned long long K_in = sc.K_in;

        // ------------------------------------------------------------------ phase 4: emission
        if (kWords && want_matrix) {
            int8_t *dst = p.matrix + G_in * NFEAT;
            const int nb = n_own * NFEAT;
            for (int j = tid; j < nb; j += NT) {
                const int c = j / NFEAT, f = j - c * NFEAT;
                dst[j] = (int8_t)((wordS[widx(c_lo + c)] >> f) & 1u);
            }
        }
        const bool stage = want_spans && ntok_tile <= SPAN_STAGE - 1 && K_in + (unsigned long long)ntok_tile <= (unsigned long long)p.cap_tokens;
        if (want_spans || want_feats) {
            // string-relative index of character c: (G_in + c - c_lo) - (global index of its string's first character)
            unsigned long long gbase = lf_excl >= 0 ? G_in + (unsigned long long)(lf_excl - c_lo) : sc.base_in;
            uint32_t ev = (E | EW | EL | Fm) & OWN;
            int rank = 0;  // tokens counted at earlier characters of this thread
            const bool over = K_in + (unsigned long long)ntok_tile > (unsigned long long)p.cap_tokens;
            if (over && tid == 0) atomicOr(&p.result->error, 4u);
            while (ev) {
                const int i = __ffs(ev) - 1; const uint32_t b = 1u << i; ev &= ev - 1;
                const int c = c0 + i;
                const unsigned long long g = G_in + (unsigned long long)(c - c_lo);
                if (Fm & b) gbase = g;
                const int idx = (int)(g - gbase);
                const int lord = tp + rank;                        // tile-local ordinal of the token counted here
                const long long ordx = (long long)K_in + lord;     // tokens counted before this character
                if (EW & b) {  // previous token [.., idx)
                    const long long k = ordx - 1;
                    if (k >= 0 && k < p.cap_tokens) {
                        if (want_spans) { if (stage && lord >= 1) spanS[2 * (lord - 1) + 3] = idx; else p.spans[2 * k + 1] = idx; }
                        if (kWords && want_feats) {
                            unsigned acc[7] = {0, 0, 0, 0, 0, 0, 0}; bool hit;
                            walk_back(c - 1, acc, hit);
                            if (!hit) for (int q = 0; q < 7; ++q) acc[q] = __vadd4(acc[q], sc.carry_sums[q]);
                            int8_t *row = p.feats + k * NFEAT;
                            for (int f = 0; f < NFEAT; ++f) row[f] = (int8_t)((acc[f >> 2] >> ((f & 3) * 8)) & 0xFFu);
                        }
                    }
                }
                if (E & b) {
                    const int sidx = (SPLIT & b) ? idx : idx - 1;
                    if (want_spans && ordx < p.cap_tokens) { if (stage) spanS[2 * lord + 2] = sidx; else p.spans[2 * ordx] = sidx; }
                    ++rank;
                }
                if (EL & b) {  // last token of the string [.., idx + 1)
                    const int ll = tp + rank - 1;
                    const long long k = (long long)K_in + ll;
                    if (k >= 0 && k < p.cap_tokens) {
                        if (want_spans) { if (stage && ll >= 0) spanS[2 * ll + 3] = idx + 1; else p.spans[2 * k + 1] = idx + 1; }
                        if (kWords && want_feats) {
                            unsigned acc[7] = {0, 0, 0, 0, 0, 0, 0}; bool hit;
                            walk_back(c, acc, hit);
                            if (!hit) for (int q = 0; q < 7; ++q) acc[q] = __vadd4(acc[q], sc.carry_sums[q]);
                            int8_t *row = p.feats + k * NFEAT;
                            for (int f = 0; f < NFEAT; ++f) row[f] = (int8_t)((acc[f >> 2] >> ((f & 3) * 8)) & 0xFFu);
                        }
                    }
                }
            }
        }
        // per-string CSR offsets for the strings that start in the owned byte range
        {
            const long long s_end = p.tile_first_str[tile + 1];
            for (long long s = p.tile_first_str[tile] + tid; s < s_end; s += NT) {
                const int wb = int(p.offsets[s] - w0);
                const int c = cidx(wb);
                p.char_off[s] = (long long)(G_in + (unsigned long long)(c - c_lo));
                const int t = wb >> 5;
                p.tok_off[s] = (long long)K_in + tokprefS[t] + __popc(emitS[t] & mask_lt(c - cprefS[t]));
            }
        }
        if (tile == p.ntiles - 1 && tid == 0) {
            p.result->n_chars = G_in + (unsigned long long)n_own;
            p.result->n_tokens = K_in + (unsigned long long)ntok_tile;
        }
        if (stage) {
            // staged spans -> global, 8-byte pairs (the end of the token open at the tile end is written by a later tile;
            // the end of the token open at the tile start, staged in slot -1, goes to token K_in - 1)
            __syncthreads();
            int2 *dst = reinterpret_cast<int2 *>(p.spans) + K_in;
            const int2 *src = reinterpret_cast<const int2 *>(spanS) + 1;
            // the last staged token may still be open (its end is written by a later tile): write only its start
            for (int i = tid; i < ntok_tile; i += NT) {
                const int2 v = src[i];
                if (v.y >= 0) dst[i] = v;
                else p.spans[2 * (K_in + i)] = v.x;
            }
        }
    }
}

template <bool kDefault, bool kWords>
static cudaError_t launch_one(const Params &p, int grid, cudaStream_t s)
{
    const size_t smem = tokenize_smem_bytes(p.tl, kWords);
    static size_t configured = 0;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(tokenize_kernel<kDefault, kWords>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    tokenize_kernel<kDefault, kWords><<<grid, NT, smem, s>>>(p);
    return cudaGetLastError();
}

int tokenize_ctas_per_sm(const TableLayout &tl, bool is_default, bool want_words)
{
    int nb = 0;
    const size_t smem = tokenize_smem_bytes(tl, want_words);
    cudaError_t e;
    if (is_default && !want_words) { cudaFuncSetAttribute(tokenize_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tokenize_kernel<true, false>, NT, smem); }
    else if (is_default) { cudaFuncSetAttribute(tokenize_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tokenize_kernel<true, true>, NT, smem); }
    else if (!want_words) { cudaFuncSetAttribute(tokenize_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tokenize_kernel<false, false>, NT, smem); }
    else { cudaFuncSetAttribute(tokenize_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, tokenize_kernel<false, true>, NT, smem); }
    if (e != cudaSuccess) { cudaGetLastError(); return 1; }
    return nb < 1 ? 1 : nb;
}

cudaError_t launch_tokenize(const Params &p, int grid, cudaStream_t s)
{
    const bool words = (p.what & (4u | 8u)) != 0u;
    if (p.rules.is_default) return words ? launch_one<true, true>(p, grid, s) : launch_one<true, false>(p, grid, s);
    return words ? launch_one<false, true>(p, grid, s) : launch_one<false, false>(p, grid, s);
}

// =====================================================================================================
// tile_first_str[t] = first string whose byte offset is >= t * TILE  (t = 0 .. ntiles; [ntiles] = S + 1).
// Also validates the offsets array.
__global__ void tile_index_kernel(const long long *offsets, long long n_strings, long long n_bytes,
                                  long long *first_str, long long ntiles, Result *result)
{
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    long long t_begin = 0, t_end = -1;
    if (s <= n_strings) {
        const long long cur = offsets[s];
        const long long prev = s ? offsets[s - 1] : -1;
        bool bad = cur < 0 || cur > n_bytes || cur < prev || (s == 0 && cur != 0) || (s == n_strings && cur != n_bytes);
        if (bad) atomicOr(&result->error, 2u);
        else {
            t_begin = prev < 0 ? 0 : prev / TILE + 1;
            t_end = cur / TILE;
            if (t_end > ntiles - 1) t_end = ntiles - 1;
        }
        if (s == n_strings) first_str[ntiles] = n_strings + 1;
    }
    // short ranges: each lane writes its own; long ranges: the whole warp helps
    const long long len = t_end - t_begin + 1;
    if (len > 0 && len <= 4)
        for (long long t = t_begin; t <= t_end; ++t) first_str[t] = s;
    unsigned long_mask = __ballot_sync(0xFFFFFFFFu, len > 4);
    while (long_mask) {
        const int src = __ffs(long_mask) - 1; long_mask &= long_mask - 1;
        const long long b = __shfl_sync(0xFFFFFFFFu, t_begin, src), e = __shfl_sync(0xFFFFFFFFu, t_end, src);
        const long long ss = __shfl_sync(0xFFFFFFFFu, s, src);
        for (long long t = b + lane; t <= e; t += 32) first_str[t] = ss;
    }
}

cudaError_t launch_tile_index(const long long *offsets, long long n_strings, long long n_bytes,
                              long long *tile_first_str, long long ntiles, Result *result, cudaStream_t s)
{
    const int bs = 256;
    const long long n = n_strings + 1;
    const unsigned grid = (unsigned)((n + bs - 1) / bs);
    tile_index_kernel<<<grid, bs, 0, s>>>(offsets, n_strings, n_bytes, tile_first_str, ntiles, result);
    return cudaGetLastError();
}

// =====================================================================================================
// Stand-alone _gen_block_mask(a1, a2) (latok.c:140-258) on caller arrays: one CTA, a forward sweep
// (backlog at every space) and a backward sweep (blank the chunks whose closing space is "hot").
constexpr int BM_NT = 1024;
__global__ void __launch_bounds__(BM_NT, 1)
block_mask_kernel(const int8_t *a1, long long s1, const int8_t *a2, long long s2, long long n, int8_t *out,
                  unsigned char *hot)
{
    __shared__ int scr[3 * 32 + 8];
    __shared__ int s_any_mark, s_any_space, s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_any_mark = 0; s_any_space = 0; s_carry = 0; }
    __syncthreads();
    bool am = false, as = false;
    for (long long i = tid; i < n; i += BM_NT) { am |= a1[i * s1] != 0; as |= a2[i * s2] != 0; }
    if (am) s_any_mark = 1;
    if (as) s_any_space = 1;
    __syncthreads();
    if (!s_any_mark || !s_any_space) {  // latok.c:191-196 / :211-216
        const int8_t v = s_any_mark ? 0 : 1;
        for (long long i = tid; i < n; i += BM_NT) out[i] = v;
        return;
    }
    const long long nchunks = (n + BM_NT - 1) / BM_NT;
    // forward: x = backlog; hot[i] = space i closes a blanked chunk
    for (long long ch = 0; ch < nchunks; ++ch) {
        const long long i = ch * BM_NT + tid;
        const bool mk = i < n && a1[i * s1] != 0, spc = i < n && a2[i * s2] != 0;
        Fn f = fn_id();
        if (mk) { f.u = 1; f.v = NEG + 1; }
        if (spc) { f.u = max(f.u - 1, NEG); f.v = max(f.v - 1, 0); }
        Fn inc = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int ou = __shfl_up_sync(0xFFFFFFFFu, inc.u, d), ov = __shfl_up_sync(0xFFFFFFFFu, inc.v, d);
            if (lane >= d) inc = fn_compose(Fn{ou, ov}, inc);
        }
        int eu = __shfl_up_sync(0xFFFFFFFFu, inc.u, 1), ev = __shfl_up_sync(0xFFFFFFFFu, inc.v, 1);
        Fn wex = lane ? Fn{eu, ev} : fn_id();
        if (lane == 31) { scr[2 * warp] = inc.u; scr[2 * warp + 1] = inc.v; }
        __syncthreads();
        Fn base = fn_id(), tot = fn_id();
        for (int w = 0; w < BM_NT / 32; ++w) {
            Fn t = Fn{scr[2 * w], scr[2 * w + 1]};
            if (w < warp) base = fn_compose(base, t);
            tot = fn_compose(tot, t);
        }
        const int carry = s_carry;
        const int x = fn_apply(fn_compose(base, wex), carry) + (mk ? 1 : 0);
        if (i < n) hot[i] = (spc && x >= 1) ? 1 : 0;
        __syncthreads();
        if (tid == 0) s_carry = fn_apply(tot, carry);
        __syncthreads();
    }
    // backward: every non-space position takes the hot flag of the next space (or of the virtual end)
    if (tid == 0) s_carry = s_carry >= 1 ? 1 : 0;  // latok.c:239-244: marks left after the last space
    __syncthreads();
    for (long long ch = nchunks - 1; ch >= 0; --ch) {
        const long long i = ch * BM_NT + tid;
        const bool spc = i < n && a2[i * s2] != 0;
        const bool h = spc && hot[i];
        const unsigned H = __ballot_sync(0xFFFFFFFFu, spc), FH = __ballot_sync(0xFFFFFFFFu, h);
        if (lane == 0) { scr[2 * warp] = H != 0u; scr[2 * warp + 1] = H ? ((FH >> (__ffs(H) - 1)) & 1u) : 0u; }
        __syncthreads();
        const int carry = s_carry;
        int cin;
        const unsigned above = lane == 31 ? 0u : (H & (0xFFFFFFFFu << (lane + 1)));
        if (above) cin = (FH >> (__ffs(above) - 1)) & 1u;
        else {
            cin = carry;
            for (int w2 = warp + 1; w2 < BM_NT / 32; ++w2)
                if (scr[2 * w2]) { cin = scr[2 * w2 + 1]; break; }
        }
        if (i < n) out[i] = (spc || i == 0) ? 1 : (cin ? 0 : 1);
        int first = carry;
        for (int w2 = 0; w2 < BM_NT / 32; ++w2)
            if (scr[2 * w2]) { first = scr[2 * w2 + 1]; break; }
        __syncthreads();
        if (tid == 0) s_carry = first;
        __syncthreads();
    }
}

cudaError_t launch_block_mask(const int8_t *a1, long long s1, const int8_t *a2, long long s2, long long n,
                              int8_t *out, unsigned char *scratch, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    block_mask_kernel<<<1, BM_NT, 0, s>>>(a1, s1, a2, s2, n, out, scratch);
    return cudaGetLastError();
}

// =====================================================================================================
// Stand-alone _combine_matrix_rows(m, idxs) (latok.c:275-370): one thread per output column.
__global__ void combine_rows_kernel(const int8_t *m, long long m_rows, long long m_cols, long long sr, long long sc,
                                    const int8_t *idx, int idx_rows, int idx_cols, int8_t *out)
{
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= m_cols) return;
    const unsigned char *mu = reinterpret_cast<const unsigned char *>(m);
    unsigned char result = 0;
    if (idx_cols > 0) {          // "and" within a row, "or" across rows (:318-341)
        unsigned char row = 0;
        for (int i = 0; i < idx_rows; ++i) {
            for (int j = 0; j < idx_cols; ++j) {
                const unsigned char r = (unsigned char)idx[i * idx_cols + j];
                if (r < 255 && r < m_rows) {
                    const unsigned char v = mu[r * sr + k * sc];
                    row = j == 0 ? v : (unsigned char)(row * v);
                }
            }
            result = (unsigned char)(result + row);
        }
    } else {                     // 1-D: plain sum of the listed rows (:342-354)
        for (int j = 0; j < idx_rows; ++j) {
            const unsigned char r = (unsigned char)idx[j];
            if (r < 255 && r < m_rows) result = (unsigned char)(result + mu[r * sr + k * sc]);
        }
    }
    out[k] = (int8_t)result;
}

cudaError_t launch_combine_rows(const int8_t *m, long long m_rows, long long m_cols, long long stride_r,
                                long long stride_c, const int8_t *idx, int idx_rows, int idx_cols,
                                int8_t *out, cudaStream_t s)
{
    if (m_cols <= 0) return cudaSuccess;
    const int bs = 256;
    combine_rows_kernel<<<(unsigned)((m_cols + bs - 1) / bs), bs, 0, s>>>(m, m_rows, m_cols, stride_r, stride_c, idx,
                                                                         idx_rows, idx_cols, out);
    return cudaGetLastError();
}

}  // namespace latok
