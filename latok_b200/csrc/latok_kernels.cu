// latok_kernels.cu -- sm_100a kernels for LaTok's tokenization hot path.
//
// One persistent kernel (tokenize_kernel) makes a single pass over the flat UTF-8 buffer:
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
constexpr int MAXC = WINB;                // at most one character per window byte
constexpr int WORDS_LEN = 1 + MAXC + (MAXC >> 5) + 1 + 40;
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
struct SmemPlan {
    int mbar, scal, tile, table, words, vals, startbits, leadmask, cpref, emit, tokpref, split, scratch, total;
};
__host__ __device__ inline SmemPlan smem_plan(int table_bytes)
{
    SmemPlan s; int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    s.mbar = take(16);
    s.scal = take(256);
    s.tile = take(WINB + 16);
    s.table = take(table_bytes);
    s.words = take(WORDS_LEN * 4);
    s.vals = take(WINB + 32);
    s.startbits = take(NT * 4);
    s.leadmask = take(NT * 4);
    s.cpref = take((NT + 1) * 4);
    s.emit = take(NT * 4);
    s.tokpref = take((NT + 1) * 4);
    s.split = take(NT * 4);
    s.scratch = take(1024);
    s.total = o;
    return s;
}
size_t tokenize_smem_bytes(const TableLayout &tl) { return (size_t)smem_plan(tl.total).total; }

struct Scalars {          // block-shared scalars
    long long tile;
    unsigned long long G_in, base_in, K_in;
    int x_in, x_end;
    int agg_u, agg_v;
    int need_walk, far;
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

// exclusive scan of backlog functions (composition) and of "last string start" (max)
__device__ __forceinline__ void block_excl_fn(Fn f, int lf, int *scratch, Fn &excl, int &lf_excl, Fn &total, int &lf_total,
                                              int lane, int warp)
{
    Fn inc = f; int lfi = lf;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int ou = __shfl_up_sync(0xFFFFFFFFu, inc.u, d), ov = __shfl_up_sync(0xFFFFFFFFu, inc.v, d);
        int ol = __shfl_up_sync(0xFFFFFFFFu, lfi, d);
        if (lane >= d) { inc = fn_compose(Fn{ou, ov}, inc); lfi = max(lfi, ol); }
    }
    // exclusive within warp
    int eu = __shfl_up_sync(0xFFFFFFFFu, inc.u, 1), ev = __shfl_up_sync(0xFFFFFFFFu, inc.v, 1);
    int el = __shfl_up_sync(0xFFFFFFFFu, lfi, 1);
    Fn wex = lane ? Fn{eu, ev} : fn_id();
    int lex = lane ? el : -1;
    if (lane == 31) { scratch[3 * warp] = inc.u; scratch[3 * warp + 1] = inc.v; scratch[3 * warp + 2] = lfi; }
    __syncthreads();
    Fn base = fn_id(), tot = fn_id(); int lb = -1, lt = -1;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
        Fn t = Fn{scratch[3 * w], scratch[3 * w + 1]}; int l = scratch[3 * w + 2];
        if (w < warp) { base = fn_compose(base, t); lb = max(lb, l); }
        tot = fn_compose(tot, t); lt = max(lt, l);
    }
    __syncthreads();
    excl = fn_compose(base, wex);
    lf_excl = max(lb, lex);
    total = tot; lf_total = lt;
}

__global__ void __launch_bounds__(NT, 1) tokenize_kernel(const Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemPlan sp = smem_plan(p.tl.total);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + sp.mbar);
    Scalars &sc = *reinterpret_cast<Scalars *>(smem + sp.scal);
    uint8_t *tileS = smem + sp.tile;
    uint8_t *tableS = smem + sp.table;
    uint32_t *wordS = reinterpret_cast<uint32_t *>(smem + sp.words);
    uint8_t *valS = smem + sp.vals;
    uint32_t *startbits = reinterpret_cast<uint32_t *>(smem + sp.startbits);
    uint32_t *leadmaskS = reinterpret_cast<uint32_t *>(smem + sp.leadmask);
    int *cprefS = reinterpret_cast<int *>(smem + sp.cpref);
    uint32_t *emitS = reinterpret_cast<uint32_t *>(smem + sp.emit);
    int *tokprefS = reinterpret_cast<int *>(smem + sp.tokpref);
    uint32_t *splitS = reinterpret_cast<uint32_t *>(smem + sp.split);
    int *scratch = reinterpret_cast<int *>(smem + sp.scratch);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (ld_volatile_u32(&p.result->error) & 2u) return;  // offsets failed validation in tile_index_kernel

    // one-time per CTA: class table into shared memory, mbarrier init
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

    uint32_t phase = 0;
    const bool want_feats = (p.what & 4u) != 0u, want_matrix = (p.what & 8u) != 0u;
    const bool want_spans = (p.what & 2u) != 0u, want_splits = (p.what & 1u) != 0u;

    for (;;) {
        if (tid == 0) sc.tile = (long long)(atomicAdd(p.ticket, 1ull) - p.ticket_base);
        __syncthreads();  // (A) also fences shared-memory reuse across tiles
        const long long tile = sc.tile;
        if (tile >= p.ntiles) break;

        // ------------------------------------------------------------------ load window
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
        if (tid == 0) { sc.need_walk = 0; sc.far = 0; sc.open_has = 0; }
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

        // ------------------------------------------------------------------ phase 1: bytes -> characters
        const int wb0 = tid * 32;
        uint32_t lead;
        const uint32_t sb = startbits[tid];
        {
            const uint4 q0 = *reinterpret_cast<const uint4 *>(tileS + wb0);
            const uint4 q1 = *reinterpret_cast<const uint4 *>(tileS + wb0 + 16);
            const uint32_t wds[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t leadbits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t t = (wds[j] & 0xC0C0C0C0u) ^ 0x80808080u;   // byte == 0  <=>  continuation byte
                uint32_t m = (t | (t << 1)) & 0x80808080u;
                leadbits |= (((m >> 7) * 0x10204080u) >> 28) << (4 * j);
            }
            const long long g0 = w0 + wb0;
            const int vlo = g0 < 0 ? int(-g0 < 32 ? -g0 : 32) : 0;
            const long long rem = p.n_bytes - g0;
            const int vhi = rem <= 0 ? 0 : (rem >= 32 ? 32 : int(rem));
            const uint32_t valid = vhi > vlo ? (mask_lt(vhi) & ~mask_lt(vlo)) : 0u;
            lead = (leadbits & valid) | sb;
        }
        int c_end;
        const int c0 = block_excl_sum(__popc(lead), scratch, c_end, lane, warp);
        leadmaskS[tid] = lead;
        cprefS[tid] = c0;
        if (tid == 0) { cprefS[NT] = c_end; wordS[widx(-1)] = 0; }
        {
            uint32_t mm = lead; int c = c0;
            while (mm) {
                int k = __ffs(mm) - 1; mm &= mm - 1;
                uint32_t f = classify_at(tileS + wb0 + k, tb);
                if ((sb >> k) & 1u) f |= FIRSTBIT;
                wordS[widx(c)] = f; ++c;
            }
        }
        if (tid < 36) wordS[widx(c_end + tid)] = 0;
        __syncthreads();  // (D)

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
        const int c_trust = term_in_win ? cidx(int(p.n_bytes - w0)) : cidx(WINB - TRUST_MARGIN);
        const int n_own = c_hi - c_lo;

        // ------------------------------------------------------------------ phase 2a: context + rules
        const int cb = tid * 32;
        const uint32_t ACT = range_mask(cb, c_lo, c_trust), OWN = range_mask(cb, c_lo, c_hi);
        uint32_t r[36];
#pragma unroll
        for (int j = 0; j < 36; ++j) r[j] = wordS[widx(cb - 1 + j)];
        uint32_t Sraw = 0, Mm = 0, Fm = 0, Lm = 0;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t w = r[i + 1];
            const bool F = (w & FIRSTBIT) != 0u, L = (r[i + 2] & FIRSTBIT) != 0u, L2 = (r[i + 3] & FIRSTBIT) != 0u;
            uint32_t full = make_word(r[i], w, r[i + 2], r[i + 3], F, L, L2);
            uint32_t cnt, sy; bool mk;
            eval_rules(p.rules, full, cnt, mk, sy);
            const uint32_t bit = 1u << i;
            if (full & (1u << 5)) Sraw |= bit;
            if (mk) Mm |= bit;
            if (F) Fm |= bit;
            if (L) Lm |= bit;
            pk[i >> 2] |= ((cnt & 15u) | ((sy & 15u) << 4)) << ((i & 3) * 8);
            r[i + 1] = full | (F ? FIRSTBIT : 0u) | (L ? LASTBIT : 0u);
        }
        const uint32_t S = Sraw & ACT;
        Mm &= ACT; Fm &= ACT; Lm &= ACT;

        // local backlog function: +1 per mark, max(x-1,0) per space, reset at a string start
        auto build_fn = [&](uint32_t ev) -> Fn {
            Fn f = fn_id();
            while (ev) {
                const uint32_t b = ev & (0u - ev); ev &= ev - 1;
                if (Fm & b) { f.u = NEG; f.v = 0; }
                if (Mm & b) { f.u = max(f.u + 1, NEG); f.v = f.v + 1; }
                if (S & b) { f.u = max(f.u - 1, NEG); f.v = max(f.v - 1, 0); }
            }
            return f;
        };
        const uint32_t EV = Mm | S | Fm;
        const Fn f_own = build_fn(EV & OWN);
        const Fn f_act = fn_compose(f_own, build_fn(EV & ~OWN));
        const uint32_t FO = Fm & OWN;
        const int lf_own = FO ? cb + 31 - __clz(FO) : -1;
        Fn excl, total; int lf_excl, lf_total;
        block_excl_fn(f_act, lf_own, scratch, excl, lf_excl, total, lf_total, lane, warp);
        if (n_own > 0 && tid == ((c_hi - 1) >> 5)) { Fn a = fn_compose(excl, f_own); sc.agg_u = a.u; sc.agg_v = a.v; }
        if (n_own == 0 && tid == 0) { sc.agg_u = 0; sc.agg_v = NEG; }
        if (want_feats || want_matrix) {
#pragma unroll
            for (int i = 0; i < 32; ++i) wordS[widx(cb + i)] = r[i + 1];
        }
        __syncthreads();  // (E)

        // ------------------------------------------------------------------ chain 1
        if (warp == 0) {
            Chain1 a;
            a.n = (unsigned long long)n_own;
            a.has = lf_total >= 0 ? 1u : 0u;
            a.lf = a.has ? (unsigned long long)(lf_total - c_lo) : 0ull;
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
            int x = x_t; uint32_t ev = EV | Lm;
            while (ev) {
                const uint32_t b = ev & (0u - ev); ev &= ev - 1;
                if (Fm & b) x = 0;
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
            const unsigned H = __ballot_sync(0xFFFFFFFFu, hasCL), FH = __ballot_sync(0xFFFFFFFFu, firstHot);
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
        if (n_own > 0 && tid == ((c_hi - 1) >> 5)) {
            const bool open = (CL & (0xFFFFFFFFu << ((c_hi - 1) & 31))) == 0u && cin == 2;
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
        uint32_t SPLIT = 0;
        {
            uint32_t vals[8];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t bit = 1u << i;
                const uint32_t e8 = (pk[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
                const uint32_t cnt = e8 & 15u, sy = e8 >> 4;
                // splits = split_cnt * block_mask + sym; splits[0] = 1   (default_tokenizer.py:121-132)
                const bool blank = (Zm & bit) && !(Sraw & bit);
                uint32_t v = ((blank ? 0u : cnt) + sy) & 0xFFu;
                if (Fm & bit) v = 1u;
                if (v) SPLIT |= bit;
                if ((i & 3) == 0) vals[i >> 2] = 0;
                vals[i >> 2] |= v << ((i & 3) * 8);
            }
            *reinterpret_cast<uint4 *>(valS + cb) = make_uint4(vals[0], vals[1], vals[2], vals[3]);
            *reinterpret_cast<uint4 *>(valS + cb + 16) = make_uint4(vals[4], vals[5], vals[6], vals[7]);
        }
        const uint32_t PS = (Sraw << 1) | ((r[0] >> 5) & 1u);            // previous character is a space
        const uint32_t E = ((SPLIT & ~Sraw) | (~SPLIT & PS & ~Fm)) & OWN;  // a token is counted at this character
        const uint32_t EW = SPLIT & ~Fm & ~PS & OWN;                       // this split ends the previous token
        const uint32_t EL = Lm & ~Sraw & OWN;                              // end of string ends the last token
        int ntok_tile;
        const int tp = block_excl_sum(__popc(E), scratch, ntok_tile, lane, warp);
        emitS[tid] = E; tokprefS[tid] = tp; splitS[tid] = SPLIT;
        if (tid == 0) tokprefS[NT] = ntok_tile;
        __syncthreads();  // (I)

        auto is_split = [&](int c) -> bool { return (splitS[c >> 5] >> (c & 31)) & 1u; };
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
        if (want_feats && n_own > 0 && tid == ((c_hi - 1) >> 5)) {
            unsigned acc[7] = {0, 0, 0, 0, 0, 0, 0}; bool hit;
            walk_back(c_hi - 1, acc, hit);
            sc.open_has = hit ? 1u : 0u;
            for (int g = 0; g < 7; ++g) sc.open_sums[g] = acc[g];
            sc.open_sums[7] = 0;
        }
        if (want_feats) __syncthreads();

        // ------------------------------------------------------------------ chain 2
        if (warp == 0) {
            Chain2 a;
            a.k = (unsigned long long)ntok_tile; a.reset = 0;
            a.has_split = want_feats ? sc.open_has : 0u;
            for (int g = 0; g < 8; ++g) a.sums[g] = (want_feats && n_own > 0) ? sc.open_sums[g] : 0u;
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
        __syncthreads();  // (J)
        const unsigned long long K_in = sc.K_in;

        // ------------------------------------------------------------------ phase 4: emission
        if (want_splits) {
            int8_t *dst = p.splits + G_in;
            for (int j = tid; j < n_own; j += NT) dst[j] = (int8_t)valS[c_lo + j];
        }
        if (want_matrix) {
            int8_t *dst = p.matrix + G_in * NFEAT;
            const int nb = n_own * NFEAT;
            for (int j = tid; j < nb; j += NT) {
                const int c = j / NFEAT, f = j - c * NFEAT;
                dst[j] = (int8_t)((wordS[widx(c_lo + c)] >> f) & 1u);
            }
        }
        if (want_spans || want_feats) {
            // string-relative index of character c: (G_in + c - c_lo) - (global index of its string's first character)
            unsigned long long gbase = lf_excl >= 0 ? G_in + (unsigned long long)(lf_excl - c_lo) : sc.base_in;
            uint32_t ev = (E | EW | EL | Fm) & OWN;
            int rank = 0;  // tokens counted at earlier characters of this thread
            const bool over = K_in + (unsigned long long)ntok_tile > (unsigned long long)p.cap_tokens;
            if (over && tid == 0) atomicOr(&p.result->error, 4u);
            while (ev) {
                const int i = __ffs(ev) - 1; const uint32_t b = 1u << i; ev &= ev - 1;
                const int c = cb + i;
                const unsigned long long g = G_in + (unsigned long long)(c - c_lo);
                if (Fm & b) gbase = g;
                const int idx = (int)(g - gbase);
                const long long ordx = (long long)K_in + tp + rank;   // tokens counted before this character
                if (EW & b) {  // previous token [.., idx)
                    const long long k = ordx - 1;
                    if (k >= 0 && k < p.cap_tokens) {
                        if (want_spans) p.spans[2 * k + 1] = idx;
                        if (want_feats) {
                            unsigned acc[7] = {0, 0, 0, 0, 0, 0, 0}; bool hit;
                            walk_back(c - 1, acc, hit);
                            if (!hit) for (int q = 0; q < 7; ++q) acc[q] = __vadd4(acc[q], sc.carry_sums[q]);
                            int8_t *row = p.feats + k * NFEAT;
                            for (int f = 0; f < NFEAT; ++f) row[f] = (int8_t)((acc[f >> 2] >> ((f & 3) * 8)) & 0xFFu);
                        }
                    }
                }
                if (E & b) {
                    if (want_spans && ordx < p.cap_tokens) p.spans[2 * ordx] = (SPLIT & b) ? idx : idx - 1;
                    ++rank;
                }
                if (EL & b) {  // last token of the string [.., idx + 1)
                    const long long k = (long long)K_in + tp + rank - 1;
                    if (k >= 0 && k < p.cap_tokens) {
                        if (want_spans) p.spans[2 * k + 1] = idx + 1;
                        if (want_feats) {
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
                const int c = cidx(int(p.offsets[s] - w0));
                p.char_off[s] = (long long)(G_in + (unsigned long long)(c - c_lo));
                p.tok_off[s] = (long long)K_in + tokprefS[c >> 5] + __popc(emitS[c >> 5] & mask_lt(c & 31));
            }
        }
        if (tile == p.ntiles - 1 && tid == 0) {
            p.result->n_chars = G_in + (unsigned long long)n_own;
            p.result->n_tokens = K_in + (unsigned long long)ntok_tile;
        }
    }
}

cudaError_t launch_tokenize(const Params &p, int grid, cudaStream_t s)
{
    const size_t smem = tokenize_smem_bytes(p.tl);
    static size_t configured = 0;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(tokenize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    tokenize_kernel<<<grid, NT, smem, s>>>(p);
    return cudaGetLastError();
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
