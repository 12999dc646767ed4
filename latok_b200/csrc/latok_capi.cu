// latok_capi.cu -- host side of liblatok_b200.so: the C ABI declared in include/latok_b200.h.
//
// Owns device / pinned memory, streams and the packed Unicode class table, validates arguments
// (the reference reports errors with PyErr_SetString(PyExc_ValueError, ...), latok.c:40-50,
// 151-171, 292-312; here they become status codes + latok_b200_last_error()) and launches the
// kernels of latok_kernels.cu.  There is no CPU implementation of the path in this library.
#include "../../include/latok_b200.h"
#include "latok_internal.h"
#include "_gen/latok_tables.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace latok;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? LATOK_B200_ENOMEM : LATOK_B200_ECUDA,          \
                        "%s failed: %s", #call, cudaGetErrorString(_e));                                 \
    } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    int ensure(size_t n, bool zero = false)
    {
        if (n <= cap) return 0;
        size_t want = n + n / 8 + 64;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        CU(cudaMalloc((void **)&p, want * sizeof(T)));
        if (zero) { CU(cudaMemset(p, 0, want * sizeof(T))); CU(cudaDeviceSynchronize()); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        size_t want = n + n / 8 + 64;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        CU(cudaMallocHost((void **)&p, want * sizeof(T)));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

uint32_t row_mask(const int8_t *row, int cols, bool &ok)
{
    uint32_t m = 0;
    for (int j = 0; j < cols; ++j) {
        int v = row[j];
        if (v == -1) continue;
        if (v < 0 || v >= NFEAT) { ok = false; return 0; }
        m |= 1u << v;
    }
    if (m == 0) ok = false;  // a row of only -1 has no defined value in the reference (latok.c:325-337)
    return m;
}

// default rules = C_SPLIT / C_MASK / C_SYM of default_tokenizer.py:49-55, 80-91, 100-102 (column
// numbers from offsets.py:24-48)
RuleSet default_rules()
{
    RuleSet r;
    memset(&r, 0, sizeof r);
    r.is_default = 1;
    const uint32_t split[] = {1u << 5, 1u << 6, 1u << 20, (1u << 4) | (1u << 17), (1u << 4) | (1u << 16)};
    const uint32_t mask[] = {(1u << 7) | (1u << 18) | (1u << 13), (1u << 11) | (1u << 18) | (1u << 21) | (1u << 23),
                             (1u << 8) | (1u << 14) | (1u << 15), (1u << 9) | (1u << 22) | (1u << 24) | (1u << 12)};
    const uint32_t sym[] = {(1u << 6) | (1u << 19)};
    r.n_split = 5; r.n_mask = 4; r.n_sym = 1;
    memcpy(r.split, split, sizeof split);
    memcpy(r.mask, mask, sizeof mask);
    memcpy(r.sym, sym, sizeof sym);
    return r;
}

bool same_rows(const uint32_t *a, int na, const uint32_t *b, int nb)
{
    if (na != nb) return false;
    std::vector<uint32_t> x(a, a + na), y(b, b + nb);
    for (auto v : x) {
        bool found = false;
        for (auto &w : y) if (w == v) { w = 0xFFFFFFFFu; found = true; break; }
        if (!found) return false;
    }
    return true;
}

}  // namespace

struct latok_b200_engine {
    int device = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr, aux = nullptr;   // aux: string index of the next batch while the previous one is tokenized
    cudaEvent_t ev_in = nullptr, ev_index[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    int slot = 0;                                   // which of the two index / result sets the current batch uses
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_b0 = nullptr, ev_b1 = nullptr;
    TableLayout tl{};
    RuleSet rules{};
    DevBuf<uint8_t> d_table, d_in, d_scratch, d_scratch2, d_scratch3;
    DevBuf<long long> d_off, d_first[2], d_char_off, d_tok_off;
    DevBuf<int8_t> d_splits, d_feats, d_matrix;
    DevBuf<int32_t> d_spans;
    DevBuf<AggRec> agg;
    DevBuf<IncRec> inc;
    DevBuf<OpenSums> osum;
    DevBuf<uint32_t> d_planes;
    DevBuf<unsigned long long> span_scratch;
    DevBuf<long long> d_tok_bytes;                  // token byte ranges (latok_b200_fetch_token_bytes)
    DevBuf<unsigned> d_blk_local, d_group_tot;
    DevBuf<unsigned long long> d_group_pref;
    float last_token_bytes_ms = 0.f;
    DevBuf<Result> d_result[2];
    PinBuf<uint8_t> h_in;
    PinBuf<long long> h_off;
    PinBuf<Result> h_result[2];
    unsigned epoch = 0;
    long long launches = 0;
    // current batch
    bool submitted = false, sized = false, inputs_on_stream = false;
    const uint8_t *cur_in = nullptr;
    const long long *cur_off = nullptr;
    long long n_strings = 0, n_bytes = 0;
    uint32_t what = 0;
    long long n_chars = 0, n_tokens = 0, walks = 0;
    float last_kernel_ms = 0.f;
};

// for the other translation units of the library (latok_reader.cpp)
extern "C" int latok_b200_set_error_(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

static int set_device(latok_b200_engine *e)
{
    CU(cudaSetDevice(e->device));
    return 0;
}

extern "C" {

int latok_b200_abi_version(void) { return LATOK_B200_ABI_VERSION; }
const char *latok_b200_last_error(void) { return g_err; }

int latok_b200_device_count(int *count)
{
    if (!count) return fail(LATOK_B200_EINVAL, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return LATOK_B200_OK;
}

static int build_table(latok_b200_engine *e)
{
    TableLayout &tl = e->tl;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    tl.lut3 = take(3 * LUT_ENTRIES * 4);
    tl.lutv = take(256 * 4);
    tl.ascii_feat = take(128 * 2);
    tl.class_feat = take(16 * 2);
    tl.stage1 = take(LATOK_TBL_STAGE1_LEN * (int)sizeof(latok_stage1_t));
    tl.stage2 = take(LATOK_TBL_STAGE2_LEN);
    tl.total = o;
    tl.stage1_len = LATOK_TBL_STAGE1_LEN;
    tl.stage2_len = LATOK_TBL_STAGE2_LEN;
    tl.low_limit = LATOK_TBL_LOW_LIMIT;
    static_assert(LATOK_TBL_NHIGH <= 1, "kernel supports one run above the two-stage table");
    static_assert(LATOK_TBL_SHIFT == 7, "kernel assumes 128-code-point blocks");
    tl.high_first = LATOK_HIGH_RUNS[0][0];
    tl.high_last = LATOK_HIGH_RUNS[0][1];
    tl.high_feat = LATOK_HIGH_RUNS[0][2];
    std::vector<uint8_t> blob((size_t)tl.total, 0);
    tl.high_class = 0;
    for (uint32_t c = 0; c < 16; ++c) if (LATOK_CLASS_FEAT[c] == tl.high_feat) tl.high_class = c;
    {
        uint32_t *lut3 = reinterpret_cast<uint32_t *>(blob.data() + tl.lut3);
        for (int e = 0; e < LUT_ENTRIES; ++e) {
            const uint32_t f = e < 128 ? LATOK_ASCII_FEAT[e] : (e < 256 ? 0u : LATOK_CLASS_FEAT[e - 256]);
            for (int k = 0; k < 3; ++k) {
                uint32_t w = 0;
                for (int b = 0; b < 4; ++b) w |= ((f >> (4 * k + b)) & 1u) << (8 * b);
                lut3[k * LUT_ENTRIES + e] = w;
            }
        }
        uint32_t *lutv = reinterpret_cast<uint32_t *>(blob.data() + tl.lutv);
        for (int i = 0; i < 256; ++i) {
            uint32_t w = 0;
            for (int b = 0; b < 4; ++b) w |= ((((uint32_t)i >> b) & 1u) + 2u * (((uint32_t)i >> (4 + b)) & 1u)) << (8 * b);
            lutv[i] = w;
        }
    }
    memcpy(blob.data() + tl.ascii_feat, LATOK_ASCII_FEAT, sizeof LATOK_ASCII_FEAT);
    memcpy(blob.data() + tl.class_feat, LATOK_CLASS_FEAT, sizeof LATOK_CLASS_FEAT);
    memcpy(blob.data() + tl.stage1, LATOK_STAGE1, sizeof LATOK_STAGE1);
    memcpy(blob.data() + tl.stage2, LATOK_STAGE2, sizeof LATOK_STAGE2);
    if (int rc = e->d_table.ensure(blob.size())) return rc;
    CU(cudaMemcpy(e->d_table.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    return 0;
}

int latok_b200_create(int device, size_t max_batch_bytes, int64_t max_strings, latok_b200_engine **out)
{
    if (!out) return fail(LATOK_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (max_strings < 0) return fail(LATOK_B200_EINVAL, "max_strings must be >= 0");
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(LATOK_B200_ECUDA, "no CUDA device available (%s); latok_b200 has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(LATOK_B200_EINVAL, "device %d out of range [0,%d)", device, n);
    latok_b200_engine *e = new (std::nothrow) latok_b200_engine();
    if (!e) return fail(LATOK_B200_ENOMEM, "out of host memory");
    e->device = device;
    int rc = [&]() -> int {
        CU(cudaSetDevice(device));
        CU(cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device));
        CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking));
        CU(cudaEventCreate(&e->ev_k0)); CU(cudaEventCreate(&e->ev_k1));
        CU(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&e->ev_index[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreate(&e->ev_t0)); CU(cudaEventCreate(&e->ev_t1));
        CU(cudaEventCreate(&e->ev_b0)); CU(cudaEventCreate(&e->ev_b1));
        if (int r = build_table(e)) return r;
        e->rules = default_rules();
        for (int i = 0; i < 2; ++i) {
            if (int r = e->d_result[i].ensure(1, true)) return r;
            if (int r = e->h_result[i].ensure(1)) return r;
        }
        if (max_batch_bytes) {
            if (int r = e->d_in.ensure(max_batch_bytes + 64)) return r;
            if (int r = e->d_splits.ensure(max_batch_bytes + 64)) return r;
            if (int r = e->d_spans.ensure(2 * (max_batch_bytes / 3 + 1024))) return r;
        }
        if (max_strings) {
            if (int r = e->d_off.ensure((size_t)max_strings + 1)) return r;
            if (int r = e->d_char_off.ensure((size_t)max_strings + 1)) return r;
            if (int r = e->d_tok_off.ensure((size_t)max_strings + 1)) return r;
        }
        return 0;
    }();
    if (rc) { latok_b200_destroy(e); return rc; }
    *out = e;
    return LATOK_B200_OK;
}

int latok_b200_destroy(latok_b200_engine *e)
{
    if (!e) return LATOK_B200_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->aux) cudaStreamSynchronize(e->aux);
    e->d_table.release(); e->d_in.release(); e->d_scratch.release(); e->d_scratch2.release(); e->d_scratch3.release();
    e->d_off.release(); e->d_first[0].release(); e->d_first[1].release(); e->d_char_off.release(); e->d_tok_off.release();
    e->d_splits.release(); e->d_feats.release(); e->d_matrix.release(); e->d_spans.release();
    e->agg.release(); e->inc.release(); e->osum.release(); e->d_planes.release(); e->span_scratch.release();
    e->d_tok_bytes.release(); e->d_blk_local.release(); e->d_group_tot.release(); e->d_group_pref.release();
    e->d_result[0].release(); e->d_result[1].release();
    e->h_in.release(); e->h_off.release(); e->h_result[0].release(); e->h_result[1].release();
    if (e->ev_in) cudaEventDestroy(e->ev_in);
    for (int i = 0; i < 2; ++i) { if (e->ev_index[i]) cudaEventDestroy(e->ev_index[i]); if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]); }
    if (e->aux) cudaStreamDestroy(e->aux);
    if (e->ev_k0) cudaEventDestroy(e->ev_k0);
    if (e->ev_k1) cudaEventDestroy(e->ev_k1);
    if (e->ev_t0) cudaEventDestroy(e->ev_t0);
    if (e->ev_b0) cudaEventDestroy(e->ev_b0);
    if (e->ev_b1) cudaEventDestroy(e->ev_b1);
    if (e->ev_t1) cudaEventDestroy(e->ev_t1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return LATOK_B200_OK;
}

int latok_b200_set_rules(latok_b200_engine *e, const int8_t *split, int sr, int sc, const int8_t *mask, int mr,
                         int mc, const int8_t *sym, int yr, int yc)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (!split && !mask && !sym) { e->rules = default_rules(); return LATOK_B200_OK; }
    if (!split || !mask || !sym) return fail(LATOK_B200_EINVAL, "must specify split, mask and sym combo matrices");
    if (sr < 1 || mr < 0 || yr < 0 || sr > MAX_RULE_ROWS || mr > MAX_RULE_ROWS || yr > MAX_RULE_ROWS)
        return fail(LATOK_B200_EINVAL, "combo matrices must have between 1 and %d rows", MAX_RULE_ROWS);
    if (sc < 1 || (mr && mc < 1) || (yr && yc < 1)) return fail(LATOK_B200_EINVAL, "combo matrices must have >= 1 column");
    RuleSet r;
    memset(&r, 0, sizeof r);
    bool ok = true, has_space = false;
    for (int i = 0; i < sr; ++i) { r.split[i] = row_mask(split + i * sc, sc, ok); has_space |= r.split[i] == (1u << 5); }
    for (int i = 0; i < mr; ++i) r.mask[i] = row_mask(mask + i * mc, mc, ok);
    for (int i = 0; i < yr; ++i) r.sym[i] = row_mask(sym + i * yc, yc, ok);
    if (!ok) return fail(LATOK_B200_EINVAL, "combo matrix entries must be feature indices in [0,%d) or -1, with at least one index per row", NFEAT);
    if (!has_space)
        return fail(LATOK_B200_EINVAL, "the split combo matrix must contain the row [SPACE_IDX] (see latok_b200.h)");
    r.n_split = sr; r.n_mask = mr; r.n_sym = yr;
    const RuleSet d = default_rules();
    r.is_default = same_rows(r.split, sr, d.split, d.n_split) && same_rows(r.mask, mr, d.mask, d.n_mask) &&
                   same_rows(r.sym, yr, d.sym, d.n_sym);
    e->rules = r;
    return LATOK_B200_OK;
}

static int run_device(latok_b200_engine *e)
{
    // token-feature / matrix modes run the v4 kernel (one CTA per 7 936-byte tile); split mask + spans run v5 (one warp
    // per 3 968-byte range, V5_NW ranges per tile)
    const bool words = (e->what & LATOK_B200_MATRIX) != 0, feats = (e->what & LATOK_B200_FEATS) != 0;
    const bool use5 = !words && !getenv("LATOK_B200_FORCE_V4");
    const int unit = use5 ? V5_RANGE : TILE;
    const long long nunits = e->n_bytes / unit + 1;
    const long long ntiles = use5 ? (nunits + V5_NW - 1) / V5_NW : nunits;
    const int slot = e->slot ^= 1;
    if (int r = e->d_first[slot].ensure((size_t)nunits + 1)) return r;
    // the status word of every aggregate record carries the launch epoch, so records are zeroed only when (re)allocated
    if ((size_t)ntiles > e->agg.cap) { if (int r = e->agg.ensure((size_t)ntiles, true)) return r; }
    if (int r = e->inc.ensure((size_t)ntiles)) return r;
    if (feats) { if (int r = e->osum.ensure((size_t)(use5 ? nunits : ntiles), true)) return r; }
    if (feats && use5) { if (int r = e->d_planes.ensure(tokenize5_plane_words(nunits))) return r; }
    if (int r = e->d_splits.ensure((size_t)e->n_bytes + 64)) return r;
    if (int r = e->d_char_off.ensure((size_t)e->n_strings + 1)) return r;
    if (int r = e->d_tok_off.ensure((size_t)e->n_strings + 1)) return r;
    if (e->d_spans.cap < 2048) { if (int r = e->d_spans.ensure(2 * ((size_t)e->n_bytes / 3 + 1024))) return r; }
    if (e->what & LATOK_B200_FEATS) { if (int r = e->d_feats.ensure((e->d_spans.cap / 2) * NFEAT)) return r; }
    if (e->what & LATOK_B200_MATRIX) { if (int r = e->d_matrix.ensure(((size_t)e->n_bytes + 64) * NFEAT)) return r; }

    e->epoch = (e->epoch + 1) & 0x3FFFFFFFu;
    if (e->epoch == 0) {  // wrapped: clear stale status words
        CU(cudaMemsetAsync(e->agg.p, 0, e->agg.cap * sizeof(AggRec), e->stream));
        e->epoch = 1;
    }
    Params p;
    memset(&p, 0, sizeof p);
    p.in = e->cur_in; p.n_bytes = e->n_bytes; p.offsets = e->cur_off; p.n_strings = e->n_strings;
    p.tile_first_str = e->d_first[slot].p; p.ntiles = ntiles; p.nranges = nunits;
    p.splits = e->d_splits.p; p.char_off = e->d_char_off.p; p.spans = e->d_spans.p; p.tok_off = e->d_tok_off.p;
    p.feats = e->d_feats.p; p.matrix = e->d_matrix.p;
    p.cap_tokens = (long long)(e->d_spans.cap / 2);
    p.what = e->what;
    p.agg = e->agg.p; p.inc = e->inc.p; p.osum = e->osum.p; p.planes = e->d_planes.p; p.span_scratch = e->span_scratch.p; p.epoch = e->epoch;
    p.ticket = &e->d_result[slot].p->ticket; p.ticket_base = 0;
    p.result = e->d_result[slot].p;
    p.table_blob = e->d_table.p; p.tl = e->tl; p.rules = e->rules;
    int grid = e->n_sm * (use5 ? tokenize5_ctas_per_sm(e->tl, e->rules.is_default != 0, feats)
                               : tokenize_ctas_per_sm(e->tl, e->rules.is_default != 0, words, feats));
    if ((long long)grid > ntiles) grid = (int)ntiles;
    if (!use5) { if (int r = e->span_scratch.ensure((size_t)grid * 2 * SPAN_SCRATCH)) return r; }
    p.span_scratch = e->span_scratch.p;

    // The string index of this batch is built on the aux stream, so that with back-to-back submits it overlaps the
    // tokenize kernel of the previous batch (two index / result sets alternate).
    if (e->inputs_on_stream) {                       // host submit: the offsets arrive by a copy on the main stream
        CU(cudaEventRecord(e->ev_in, e->stream));
        CU(cudaStreamWaitEvent(e->aux, e->ev_in, 0));
    }
    CU(cudaStreamWaitEvent(e->aux, e->ev_done[slot], 0));     // this set's previous batch has been tokenized and read back
    CU(cudaMemsetAsync(e->d_result[slot].p, 0, sizeof(Result), e->aux));
    CU(launch_tile_index(e->cur_off, e->n_strings, e->n_bytes, e->d_first[slot].p, nunits, unit, e->d_result[slot].p, e->aux));
    CU(cudaEventRecord(e->ev_index[slot], e->aux));
    CU(cudaStreamWaitEvent(e->stream, e->ev_index[slot], 0));
    CU(cudaEventRecord(e->ev_k0, e->stream));
    CU(use5 ? launch_tokenize5(p, grid, e->stream) : launch_tokenize(p, grid, e->stream));
    CU(cudaEventRecord(e->ev_k1, e->stream));
    CU(cudaMemcpyAsync(e->h_result[slot].p, e->d_result[slot].p, sizeof(Result), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaEventRecord(e->ev_done[slot], e->stream));
    e->launches += 2;
    e->submitted = true;
    e->sized = false;
    return LATOK_B200_OK;
}

static int check_what(uint32_t what)
{
    if (what == 0 || (what & ~15u)) return fail(LATOK_B200_EINVAL, "`what` must be a non-empty OR of LATOK_B200_SPLITS|SPANS|FEATS|MATRIX");
    return 0;
}

int latok_b200_submit(latok_b200_engine *e, const uint8_t *utf8, const int64_t *offsets, int64_t n_strings, uint32_t what)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (n_strings < 0) return fail(LATOK_B200_EINVAL, "n_strings must be >= 0");
    if (!offsets) return fail(LATOK_B200_EINVAL, "must specify the offsets array (n_strings + 1 entries)");
    if (int r = check_what(what)) return r;
    if (offsets[0] != 0) return fail(LATOK_B200_EINVAL, "offsets[0] must be 0");
    const long long n_bytes = offsets[n_strings];
    if (n_bytes < 0) return fail(LATOK_B200_EINVAL, "offsets must be non-negative");
    if (n_bytes > 0 && !utf8) return fail(LATOK_B200_EINVAL, "must specify the UTF-8 buffer");
    if (int r = set_device(e)) return r;
    e->submitted = false;
    if (int r = e->d_in.ensure((size_t)n_bytes + 64)) return r;
    if (int r = e->d_off.ensure((size_t)n_strings + 1)) return r;
    // stage through pinned memory unless the caller's buffers already are pinned
    auto is_pinned = [](const void *ptr) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const uint8_t *src_b = utf8;
    const long long *src_o = (const long long *)offsets;
    if (n_bytes > 0 && !is_pinned(utf8)) {
        if (int r = e->h_in.ensure((size_t)n_bytes)) return r;
        CU(cudaStreamSynchronize(e->stream));  // staging buffer may still feed a previous copy
        memcpy(e->h_in.p, utf8, (size_t)n_bytes);
        src_b = e->h_in.p;
    }
    if (!is_pinned(offsets)) {
        if (int r = e->h_off.ensure((size_t)n_strings + 1)) return r;
        CU(cudaStreamSynchronize(e->stream));
        memcpy(e->h_off.p, offsets, sizeof(long long) * ((size_t)n_strings + 1));
        src_o = e->h_off.p;
    }
    if (n_bytes > 0) CU(cudaMemcpyAsync(e->d_in.p, src_b, (size_t)n_bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_off.p, src_o, sizeof(long long) * ((size_t)n_strings + 1), cudaMemcpyHostToDevice, e->stream));
    e->cur_in = e->d_in.p; e->cur_off = e->d_off.p;
    e->n_strings = n_strings; e->n_bytes = n_bytes; e->what = what;
    e->inputs_on_stream = true;
    return run_device(e);
}

int latok_b200_submit_device(latok_b200_engine *e, const uint8_t *d_utf8, const int64_t *d_offsets, int64_t n_strings,
                             int64_t n_bytes, uint32_t what)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (n_strings < 0 || n_bytes < 0) return fail(LATOK_B200_EINVAL, "n_strings and n_bytes must be >= 0");
    if (!d_offsets || (n_bytes > 0 && !d_utf8)) return fail(LATOK_B200_EINVAL, "must specify device buffers");
    if (((uintptr_t)d_utf8 & 15u) != 0) return fail(LATOK_B200_EINVAL, "d_utf8 must be 16-byte aligned");
    if (int r = check_what(what)) return r;
    if (int r = set_device(e)) return r;
    e->submitted = false;
    e->cur_in = d_utf8; e->cur_off = (const long long *)d_offsets;
    e->n_strings = n_strings; e->n_bytes = n_bytes; e->what = what;
    e->inputs_on_stream = false;
    return run_device(e);
}

int latok_b200_sizes(latok_b200_engine *e, int64_t *n_chars, int64_t *n_tokens)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (!e->submitted) return fail(LATOK_B200_ESTATE, "no batch submitted");
    if (int r = set_device(e)) return r;
    for (int attempt = 0; attempt < 4 && !e->sized; ++attempt) {
        CU(cudaStreamSynchronize(e->stream));
        const Result res = *e->h_result[e->slot].p;
        if (res.error & 2u) { e->submitted = false; return fail(LATOK_B200_EINVAL, "offsets must start at 0, be non-decreasing and end at the buffer length"); }
        if (res.error & 1u) { e->submitted = false; return fail(LATOK_B200_EINTERNAL, "device look-back watchdog tripped"); }
        if (res.error & 8u) {
            e->submitted = false;
            return fail(LATOK_B200_EINTERNAL, "device consistency check failed: tile %llu G_in %llu K_in %llu n_own %lld ntok %lld c_lo %lld c_hi %lld",
                        res.prof[8], res.prof[9], res.prof[10], (long long)res.prof[11], (long long)res.prof[12], (long long)res.prof[13], (long long)res.prof[14]);
        }
        if ((res.error & 4u) || (long long)res.n_tokens > (long long)(e->d_spans.cap / 2)) {
            // token buffers too small: grow to the exact count and run the batch again
            const size_t need = (size_t)res.n_tokens + 1024;
            if (int r = e->d_spans.ensure(2 * need)) return r;
            if (int r = run_device(e)) return r;
            continue;
        }
        if (getenv("LATOK_B200_PRINT_PROF")) {
            fprintf(stderr, "[latok prof]");
            for (int i = 0; i < 16; ++i) fprintf(stderr, " %llu", res.prof[i]);
            fprintf(stderr, "\n");
        }
        e->n_chars = (long long)res.n_chars; e->n_tokens = (long long)res.n_tokens; e->walks = (long long)res.walks;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e->ev_k0, e->ev_k1) == cudaSuccess) e->last_kernel_ms = ms; else cudaGetLastError();
        e->sized = true;
    }
    if (!e->sized) return fail(LATOK_B200_EINTERNAL, "token capacity did not converge");
    if (n_chars) *n_chars = e->n_chars;
    if (n_tokens) *n_tokens = e->n_tokens;
    return LATOK_B200_OK;
}

int latok_b200_fetch(latok_b200_engine *e, int8_t *splits, int64_t *char_offsets, int32_t *spans, int64_t *tok_offsets,
                     int8_t *tok_feats, int8_t *matrix)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (int r = latok_b200_sizes(e, nullptr, nullptr)) return r;
    const uint32_t w = e->what;
    if (splits && !(w & LATOK_B200_SPLITS)) return fail(LATOK_B200_ESTATE, "split mask was not requested at submit");
    if (spans && !(w & LATOK_B200_SPANS)) return fail(LATOK_B200_ESTATE, "spans were not requested at submit");
    if (tok_feats && !(w & LATOK_B200_FEATS)) return fail(LATOK_B200_ESTATE, "token features were not requested at submit");
    if (matrix && !(w & LATOK_B200_MATRIX)) return fail(LATOK_B200_ESTATE, "feature matrix was not requested at submit");
    const size_t C = (size_t)e->n_chars, T = (size_t)e->n_tokens, S1 = (size_t)e->n_strings + 1;
    cudaStream_t s = e->stream;
    if (splits && C) CU(cudaMemcpyAsync(splits, e->d_splits.p, C, cudaMemcpyDeviceToHost, s));
    if (char_offsets) CU(cudaMemcpyAsync(char_offsets, e->d_char_off.p, S1 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    if (spans && T) CU(cudaMemcpyAsync(spans, e->d_spans.p, T * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (tok_offsets) CU(cudaMemcpyAsync(tok_offsets, e->d_tok_off.p, S1 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    if (tok_feats && T) CU(cudaMemcpyAsync(tok_feats, e->d_feats.p, T * NFEAT, cudaMemcpyDeviceToHost, s));
    if (matrix && C) CU(cudaMemcpyAsync(matrix, e->d_matrix.p, C * NFEAT, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return LATOK_B200_OK;
}

int latok_b200_fetch_token_bytes(latok_b200_engine *e, int64_t *byte_spans, int on_device)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (int r = latok_b200_sizes(e, nullptr, nullptr)) return r;
    if (!(e->what & LATOK_B200_SPANS)) return fail(LATOK_B200_ESTATE, "spans were not requested at submit");
    const size_t T = (size_t)e->n_tokens;
    if (T && !byte_spans) return fail(LATOK_B200_EINVAL, "must specify the byte_spans array (2 * n_tokens entries)");
    if (!T) return LATOK_B200_OK;
    if (on_device && ((uintptr_t)byte_spans & 15u) != 0) return fail(LATOK_B200_EINVAL, "a device byte_spans array must be 16-byte aligned");
    long long *d_out = on_device ? (long long *)byte_spans : nullptr;
    if (!on_device) { if (int r = e->d_tok_bytes.ensure(2 * T)) return r; d_out = e->d_tok_bytes.p; }
    if (int r = e->d_blk_local.ensure((size_t)token_bytes_words(e->n_bytes))) return r;
    if (int r = e->d_group_tot.ensure((size_t)token_bytes_groups(e->n_bytes))) return r;
    if (int r = e->d_group_pref.ensure((size_t)token_bytes_groups(e->n_bytes))) return r;
    cudaStream_t s = e->stream;
    CU(cudaEventRecord(e->ev_b0, s));
    CU(launch_token_bytes(e->cur_in, e->n_bytes, e->cur_off, e->d_char_off.p, e->d_tok_off.p, e->n_strings, e->n_tokens, e->d_spans.p,
                          d_out, e->d_blk_local.p, e->d_group_tot.p, e->d_group_pref.p, e->d_table.p, e->tl, e->n_sm, s));
    CU(cudaEventRecord(e->ev_b1, s));
    e->launches += 3;
    if (!on_device) CU(cudaMemcpyAsync(byte_spans, d_out, T * 2 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->ev_b0, e->ev_b1) == cudaSuccess) e->last_token_bytes_ms = ms; else cudaGetLastError();
    return LATOK_B200_OK;
}

int latok_b200_token_bytes_ms(latok_b200_engine *e, float *ms)
{
    if (!e || !ms) return fail(LATOK_B200_EINVAL, "engine or ms is NULL");
    *ms = e->last_token_bytes_ms;
    return LATOK_B200_OK;
}

int latok_b200_device_results(latok_b200_engine *e, const int8_t **splits, const int64_t **char_offsets,
                              const int32_t **spans, const int64_t **tok_offsets, const int8_t **tok_feats,
                              const int8_t **matrix)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (!e->submitted) return fail(LATOK_B200_ESTATE, "no batch submitted");
    if (splits) *splits = e->d_splits.p;
    if (char_offsets) *char_offsets = (const int64_t *)e->d_char_off.p;
    if (spans) *spans = e->d_spans.p;
    if (tok_offsets) *tok_offsets = (const int64_t *)e->d_tok_off.p;
    if (tok_feats) *tok_feats = e->d_feats.p;
    if (matrix) *matrix = e->d_matrix.p;
    return LATOK_B200_OK;
}

int latok_b200_timer_begin(latok_b200_engine *e)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (int r = set_device(e)) return r;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaEventRecord(e->ev_t0, e->stream));
    return LATOK_B200_OK;
}

int latok_b200_timer_end(latok_b200_engine *e, float *elapsed_ms)
{
    if (!e || !elapsed_ms) return fail(LATOK_B200_EINVAL, "engine or elapsed_ms is NULL");
    if (int r = set_device(e)) return r;
    CU(cudaEventRecord(e->ev_t1, e->stream));
    CU(cudaEventSynchronize(e->ev_t1));
    CU(cudaEventElapsedTime(elapsed_ms, e->ev_t0, e->ev_t1));
    return LATOK_B200_OK;
}

int latok_b200_launch_count(latok_b200_engine *e, int64_t *launches)
{
    if (!e || !launches) return fail(LATOK_B200_EINVAL, "engine or launches is NULL");
    *launches = e->launches;
    return LATOK_B200_OK;
}

int latok_b200_last_stats(latok_b200_engine *e, float *tokenize_kernel_ms, int64_t *lookahead_walks)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (int r = latok_b200_sizes(e, nullptr, nullptr)) return r;
    if (tokenize_kernel_ms) *tokenize_kernel_ms = e->last_kernel_ms;
    if (lookahead_walks) *lookahead_walks = e->walks;
    return LATOK_B200_OK;
}

/* debugging aid (not part of the public header): copy the per-tile inclusive prefixes of the last run */
extern "C" __attribute__((visibility("default"))) int latok_b200_debug_chain(latok_b200_engine *e, void *inc_out, void *agg_out, int64_t max_tiles)
{
    if (!e) return 1;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    const int64_t nt = e->n_bytes / TILE + 1;
    const int64_t n = nt < max_tiles ? nt : max_tiles;
    cudaMemcpy(inc_out, e->inc.p, sizeof(IncRec) * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(agg_out, e->agg.p, sizeof(AggRec) * n, cudaMemcpyDeviceToHost);
    return (int)n;
}

int latok_b200_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(LATOK_B200_EINVAL, "ptr is NULL");
    CU(cudaMallocHost(ptr, bytes ? bytes : 1));
    return LATOK_B200_OK;
}

int latok_b200_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return LATOK_B200_OK;
}

int latok_b200_gen_parse_matrix(latok_b200_engine *e, const uint8_t *utf8, int64_t n_bytes, int64_t *n_chars, int8_t *out)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (n_bytes < 0 || (n_bytes > 0 && !utf8)) return fail(LATOK_B200_EINVAL, "must specify string to generate the parse matrix for");
    const int64_t offs[2] = {0, n_bytes};
    if (int r = latok_b200_submit(e, utf8, offs, 1, LATOK_B200_MATRIX)) return r;
    int64_t C = 0;
    if (int r = latok_b200_sizes(e, &C, nullptr)) return r;
    if (n_chars) *n_chars = C;
    if (out) return latok_b200_fetch(e, nullptr, nullptr, nullptr, nullptr, nullptr, out);
    return LATOK_B200_OK;
}

int latok_b200_gen_block_mask(latok_b200_engine *e, const int8_t *a1, int64_t stride1, const int8_t *a2, int64_t stride2,
                              int64_t n, int8_t *out)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (n < 0) return fail(LATOK_B200_EINVAL, "must specify 1d numpy arrays of matching length");
    if (n == 0) return LATOK_B200_OK;
    if (!a1 || !a2 || !out) return fail(LATOK_B200_EINVAL, "must specify two aligning 1d numpy array args");
    if (int r = set_device(e)) return r;
    if (int r = e->d_scratch.ensure((size_t)n * 2)) return r;
    if (int r = e->d_scratch2.ensure((size_t)n)) return r;
    if (int r = e->d_scratch3.ensure((size_t)n)) return r;
    std::vector<int8_t> h((size_t)n * 2);
    for (int64_t i = 0; i < n; ++i) { h[(size_t)i] = a1[i * stride1]; h[(size_t)(n + i)] = a2[i * stride2]; }
    CU(cudaMemcpyAsync(e->d_scratch.p, h.data(), (size_t)n * 2, cudaMemcpyHostToDevice, e->stream));
    CU(launch_block_mask((const int8_t *)e->d_scratch.p, 1, (const int8_t *)e->d_scratch.p + n, 1, n,
                         (int8_t *)e->d_scratch2.p, e->d_scratch3.p, e->stream));
    e->launches += 1;
    CU(cudaMemcpyAsync(out, e->d_scratch2.p, (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LATOK_B200_OK;
}

int latok_b200_combine_matrix_rows(latok_b200_engine *e, const int8_t *m, int64_t m_rows, int64_t m_cols, int64_t stride_row,
                                   int64_t stride_col, const int8_t *idx, int idx_rows, int idx_cols, int8_t *out)
{
    if (!e) return fail(LATOK_B200_EINVAL, "engine is NULL");
    if (m_rows < 0 || m_cols < 0 || idx_rows < 0 || idx_cols < 0) return fail(LATOK_B200_EINVAL, "must specify 2d numpy array args");
    if (stride_row < 0 || stride_col < 0) return fail(LATOK_B200_EINVAL, "negative strides are not supported");
    if (m_cols == 0) return LATOK_B200_OK;
    if (!m || !out || (idx_rows && !idx)) return fail(LATOK_B200_EINVAL, "must specify 2d m and idxs matrices");
    if (int r = set_device(e)) return r;
    const size_t extent = m_rows ? (size_t)((m_rows - 1) * stride_row + (m_cols - 1) * stride_col + 1) : 0;
    const size_t n_idx = (size_t)idx_rows * (size_t)(idx_cols ? idx_cols : 1);
    if (int r = e->d_scratch.ensure(extent + 16)) return r;
    if (int r = e->d_scratch2.ensure((size_t)m_cols)) return r;
    if (int r = e->d_scratch3.ensure(n_idx + 16)) return r;
    if (extent) CU(cudaMemcpyAsync(e->d_scratch.p, m, extent, cudaMemcpyHostToDevice, e->stream));
    if (n_idx) CU(cudaMemcpyAsync(e->d_scratch3.p, idx, n_idx, cudaMemcpyHostToDevice, e->stream));
    CU(launch_combine_rows((const int8_t *)e->d_scratch.p, m_rows, m_cols, stride_row, stride_col,
                           (const int8_t *)e->d_scratch3.p, idx_rows, idx_cols, (int8_t *)e->d_scratch2.p, e->stream));
    e->launches += 1;
    CU(cudaMemcpyAsync(out, e->d_scratch2.p, (size_t)m_cols, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return LATOK_B200_OK;
}

}  // extern "C"
