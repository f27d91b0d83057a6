// ldpc_b200.cu — C ABI of libldpc_b200.so (see include/ldpc_b200.h) and the host-side
// orchestration of the decode: ingest -> [cn_sweep, book, vn_sweep] x iterations -> emit.
// Built for sm_100a only; there is no CPU path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "channel_gen.cuh"
#include "common.cuh"
#include "io_kernels.cuh"
#include "res_layout.h"
#include "resident_bp.cuh"
#include "resident_vp.cuh"
#include "resident_vd.cuh"
#include "resident_vq.cuh"
#include "resident_bec.cuh"
#include "stream_bec.cuh"
#include "stream_bp.cuh"
#include "stream_bp_tma.cuh"

using namespace ldpc;

namespace {

std::string g_create_error;

int fail(ldpc_t *h, int code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                           \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail((h), LDPC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
    } while (0)

// A failed launch stops the call at once: nothing that depends on it is enqueued behind it.
#define LAUNCH(h, kern, grid, block, stream, ...)                                                   \
    do {                                                                                            \
        kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                        \
        (h)->launches++;                                                                            \
        if (cudaPeekAtLastError() != cudaSuccess) return check_launch((h), #kern);                  \
    } while (0)

// Word-wise tiles need 4-byte aligned symbol rows (io_kernels.cuh).
inline bool rows_word_aligned(const void *base, int n) { return (n % 4) == 0 && (reinterpret_cast<uintptr_t>(base) & 3u) == 0; }
inline dim3 tile_grid(int n, int wpr) { return dim3((unsigned)((n + kTileVars - 1) / kTileVars), (unsigned)((wpr + 7) / 8), 1); }

// cudaGetLastError, not Peek: the error is reported ONCE, by the call that caused it, and does not stick to the
// thread (a stale launch error would otherwise be blamed on every later ldpc_* call and on the caller's own CUDA work).
int check_launch(ldpc_t *h, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, LDPC_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return LDPC_OK;
}

// Entry of every ABI call that launches: select the device and drop whatever error an EARLIER call (ours or the
// caller's) left on this thread, so that check_launch reports this call's launches only.
#define ENTER(h)                                                                                    \
    do {                                                                                            \
        CUDA_TRY((h), cudaSetDevice((h)->device));                                                  \
        (void)cudaGetLastError();                                                                   \
    } while (0)

// Optional per-launch timing of the two sweeps (bench.py's roofline): events on the launching stream.
ProfEvent *prof_begin(ldpc_t *h, int kind, cudaStream_t s)
{
    if (!h->prof) return nullptr;
    if (h->prof_used == h->prof_ev.size()) {
        ProfEvent ev;
        if (cudaEventCreate(&ev.t0) != cudaSuccess || cudaEventCreate(&ev.t1) != cudaSuccess) return nullptr;
        h->prof_ev.push_back(ev);
    }
    ProfEvent *ev = &h->prof_ev[h->prof_used++];
    ev->kind = kind;
    cudaEventRecord(ev->t0, s);
    return ev;
}
void prof_end(ProfEvent *ev, cudaStream_t s)
{
    if (ev) cudaEventRecord(ev->t1, s);
}

constexpr int kMaxFramesPerCall = 65535 * 32 - 480;      // Bp / 32 (grid.y of the layout kernels) must stay <= 65535 after padding to 512

template <typename T> struct Fpt;
template <> struct Fpt<float> { static constexpr int value = 4; };
template <> struct Fpt<double> { static constexpr int value = 2; };

inline int frames_per_group(int dtype) { return dtype == LDPC_F64 ? 64 : 128; }     // one warp of the register sweeps
inline int frames_per_tile(int dtype) { return dtype == LDPC_F64 ? 256 : 512; }     // one CTA tile of cn_sweep_tma (2 KB rows)
inline size_t elem_size(int dtype) { return dtype == LDPC_F64 ? 8 : 4; }
inline size_t in_elem_size(int y_dtype) { return y_dtype == LDPC_F64 ? 8 : (y_dtype == LDPC_F16 ? 2 : 4); }   // received rows may be binary16
inline bool in_dtype_ok(int y_dtype) { return y_dtype == LDPC_F32 || y_dtype == LDPC_F64 || y_dtype == LDPC_F16; }

// Grid for a sweep: x = frame tiles of 8 groups, y = chunks of checks / variables, >= ~16 CTAs per SM.
dim3 sweep_grid(const ldpc_t *h, int ngroups, int items, int *per_cta)
{
    const int gx = (ngroups + kCtaWarps - 1) / kCtaWarps;
    const int target = h->sm_count * 16;
    int gy = std::max(1, std::min(items, target / std::max(1, gx)));
    gy = std::min(gy, 65535);
    int per = (items + gy - 1) / gy;
    gy = (items + per - 1) / per;
    *per_cta = per;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

// ------------------------------------------------------------------------------------------------
// workspace layouts
// ------------------------------------------------------------------------------------------------
struct BpLayout {
    int Bp, wpr, ngroups;
    size_t bytes;
};

BpLayout bp_layout(const Tables &t, int dtype, int B, bool want_marg)
{
    BpLayout L;
    const int G = frames_per_group(dtype), TF = frames_per_tile(dtype);
    L.Bp = (B + TF - 1) / TF * TF;
    L.wpr = L.Bp / 32;
    L.ngroups = L.Bp / G;
    const size_t es = elem_size(dtype);
    size_t b = 0;
    b += align_up((size_t)t.E * L.Bp * es, 256);                // msg
    b += align_up((size_t)t.n * L.Bp * es, 256);                // prior
    if (want_marg) b += align_up((size_t)t.n * L.Bp * es, 256); // marg
    b += align_up((size_t)t.n * L.wpr * 4, 256);                // xbits
    b += 2 * align_up((size_t)L.wpr * 4, 256);                  // act, unsat
    b += align_up((size_t)L.Bp * 4, 256);                       // iters
    b += 256;                                                   // any_active, live count
    b += 2 * align_up((size_t)L.Bp * 4, 256);                   // src_of, orig (active-frame compaction)
    L.bytes = b + 256;
    return L;
}

struct BecLayout {
    int Bp, wpr;
    size_t bytes;
};

BecLayout bec_layout(const Tables &t, int B)
{
    BecLayout L;
    L.Bp = (B + 127) / 128 * 128;                               // wpr % 4 == 0: the sweeps may take four words per thread
    L.wpr = L.Bp / 32;
    size_t b = 0;
    b += 2 * align_up((size_t)t.E * L.wpr * 4, 256);            // mnz, mpos
    b += 4 * align_up((size_t)t.n * L.wpr * 4, 256);            // pnz, ppos, xe, xv
    b += 4 * align_up((size_t)L.wpr * 4, 256);                  // act, changed, haser, stopped
    b += align_up((size_t)L.Bp * 4, 256);                       // iters
    b += 256;
    L.bytes = b + 256;
    return L;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch by degree profile
// ------------------------------------------------------------------------------------------------
// The dynamic shared-memory limit of a kernel instance is an attribute of (device, function) and cudaFuncSetAttribute
// SETS it (a smaller request lowers it), while a process may hold several handles on the same device (one per code)
// and handles on several devices.  So the opted-in size lives in one process-wide table keyed by (device, kernel) and
// only ever grows; the occupancy that goes with the current size is cached next to it.
struct SmemOpt {
    size_t bytes = 0;
    int per_sm = 0;
};
std::mutex g_smem_mu;
std::map<std::pair<int, const void *>, SmemOpt> g_smem_opted;

template <typename KernelT> int opt_in_smem(ldpc_t *h, KernelT kern, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_smem_mu);
    SmemOpt &slot = g_smem_opted[{h->device, reinterpret_cast<const void *>(kern)}];
    if (bytes <= slot.bytes) return LDPC_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(h, LDPC_ECUDA, std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(e));
    slot.bytes = bytes;
    slot.per_sm = 0;
    return LDPC_OK;
}

// Opt-in + CTAs per SM of an on-chip kernel at this launch geometry (cached for the opted-in size).
template <typename KernelT> int resident_occupancy(ldpc_t *h, KernelT kern, int threads, size_t smem, int *per_sm)
{
    int rc = opt_in_smem(h, kern, smem);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_smem_mu);
    SmemOpt &slot = g_smem_opted[{h->device, reinterpret_cast<const void *>(kern)}];
    if (smem == slot.bytes && slot.per_sm > 0) { *per_sm = slot.per_sm; return LDPC_OK; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int n = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n < 1) n = 1;
    if (smem == slot.bytes) slot.per_sm = n;                       // smaller launches than the opted-in size are not cached
    *per_sm = n;
    return LDPC_OK;
}

// Check-node sweep: bulk-async staged kernel for check degrees <= 8, register kernel otherwise / on request.
template <typename T, int ALGO>
int launch_cn(ldpc_t *h, BpParams<T> p, bool tma, cudaStream_t s, int extent)
{
    constexpr int FPT = Fpt<T>::value;
    const Tables &t = h->t;
    if (tma && t.max_dc <= 8) {
        const int gx = extent / (kTmaThreads * FPT);                // frame tiles in the extent (p.Bp stays the row pitch)
        int gy = std::max(1, std::min(t.m, (h->sm_count * 8) / std::max(1, gx)));
        p.per_cta = (t.m + gy - 1) / gy;
        gy = (t.m + p.per_cta - 1) / p.per_cta;
        const dim3 grid((unsigned)gx, (unsigned)gy, 1);
#define CN_TMA(DC, UNI)                                                                             \
        do {                                                                                        \
            auto kern = cn_sweep_tma<T, FPT, ALGO, DC, UNI>;                                        \
            { int rc_ = opt_in_smem(h, kern, tma_smem_bytes<DC>()); if (rc_) return rc_; }          \
            kern<<<grid, kTmaThreads, tma_smem_bytes<DC>(), s>>>(p);                                \
            h->launches++;                                                                          \
        } while (0)
        if (t.uni_dc == 4) CN_TMA(4, true);
        else if (t.uni_dc == 6) CN_TMA(6, true);
        else if (t.uni_dc == 8) CN_TMA(8, true);
        else CN_TMA(8, false);
#undef CN_TMA
        return LDPC_OK;
    }
    const dim3 grid = sweep_grid(h, p.ngroups, t.m, &p.per_cta);
#define CN_CASE(DC, UNI) LAUNCH(h, (cn_sweep<T, FPT, ALGO, DC, UNI>), grid, kCtaThreads, s, p)
    if (t.uni_dc == 4) CN_CASE(4, true);
    else if (t.uni_dc == 6) CN_CASE(6, true);
    else if (t.uni_dc == 8) CN_CASE(8, true);
    else if (t.max_dc <= 8) CN_CASE(8, false);
    else if (t.max_dc <= 16) CN_CASE(16, false);
    else if (t.max_dc <= 32) CN_CASE(32, false);
    else return fail(h, LDPC_EUNSUPPORTED, "check degree > 32 is not supported by the streaming kernels");
#undef CN_CASE
    return LDPC_OK;
}

template <typename T>
int launch_vn(ldpc_t *h, const BpParams<T> &p, dim3 grid, cudaStream_t s)
{
    constexpr int FPT = Fpt<T>::value;
    const Tables &t = h->t;
#define VN_CASE(DV, UNI) LAUNCH(h, (vn_sweep<T, FPT, DV, UNI>), grid, kCtaThreads, s, p)
    if (t.uni_dv == 3) VN_CASE(3, true);
    else if (t.uni_dv == 4) VN_CASE(4, true);
    else if (t.max_dv <= 4) VN_CASE(4, false);
    else if (t.max_dv <= 8) VN_CASE(8, false);
    else if (t.max_dv <= 16) VN_CASE(16, false);
    else if (t.max_dv <= 32) VN_CASE(32, false);
    else return fail(h, LDPC_EUNSUPPORTED, "variable degree > 32 is not supported by the streaming kernels");
#undef VN_CASE
    return LDPC_OK;
}

template <typename T>
int launch_cn_algo(ldpc_t *h, int algo, const BpParams<T> &p, bool tma, cudaStream_t s, int extent);
template <>
int launch_cn_algo<float>(ldpc_t *h, int algo, const BpParams<float> &p, bool tma, cudaStream_t s, int extent)
{
    return algo == LDPC_MSA ? launch_cn<float, ALGO_MSA>(h, p, tma, s, extent) : launch_cn<float, ALGO_SPA_PHI>(h, p, tma, s, extent);
}
template <>
int launch_cn_algo<double>(ldpc_t *h, int algo, const BpParams<double> &p, bool tma, cudaStream_t s, int extent)
{
    return algo == LDPC_MSA ? launch_cn<double, ALGO_MSA>(h, p, tma, s, extent) : launch_cn<double, ALGO_SPA_REF>(h, p, tma, s, extent);
}

// ------------------------------------------------------------------------------------------------
// input description shared by ldpc_decode and ldpc_decode_host
// ------------------------------------------------------------------------------------------------
struct InSpec {
    int channel;            // LDPC_CH_*
    int in_dtype;           // element type of src for PRIORS / BIAWGN
    const void *src;        // device [B][n]
    const uint8_t *y_hard;  // device [B][n] or NULL (PRIORS only)
    double param;
};

template <typename T>
int ingest_bp(ldpc_t *h, const InSpec &in, T *prior, uint32_t *xbits, int B, const BpLayout &L, cudaStream_t s,
              bool *have_hard)
{
    const Tables &t = h->t;
    const dim3 pgrid((t.n + 31) / 32, L.Bp / 32), block(32, 8);                       // pack_hard: one 32 x 32 tile per CTA
    const dim3 grid((t.n + 32 * kIngestTiles - 1) / (32 * kIngestTiles), L.Bp / 32);  // ingest_priors: kIngestTiles tiles per CTA
    *have_hard = false;
    switch (in.channel) {
    case LDPC_CH_PRIORS:
        if (in.in_dtype == LDPC_F64)
            LAUNCH(h, (ingest_priors<double, T, IN_COPY>), grid, block, s, (const double *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, 0.0);
        else if (in.in_dtype == LDPC_F16)
            LAUNCH(h, (ingest_priors<__half, T, IN_COPY>), grid, block, s, (const __half *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, 0.0);
        else
            LAUNCH(h, (ingest_priors<float, T, IN_COPY>), grid, block, s, (const float *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, 0.0);
        if (in.y_hard) {
            LAUNCH(h, pack_hard, pgrid, block, s, in.y_hard, xbits, B, t.n, L.wpr);
            *have_hard = true;
        }
        break;
    case LDPC_CH_BSC:
        LAUNCH(h, (ingest_priors<uint8_t, T, IN_BSC>), grid, block, s, (const uint8_t *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, in.param);
        *have_hard = true;
        break;
    case LDPC_CH_BIAWGN:
        if (in.in_dtype == LDPC_F64)
            LAUNCH(h, (ingest_priors<double, T, IN_BIAWGN>), grid, block, s, (const double *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, in.param);
        else if (in.in_dtype == LDPC_F16)
            LAUNCH(h, (ingest_priors<__half, T, IN_BIAWGN>), grid, block, s, (const __half *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, in.param);
        else
            LAUNCH(h, (ingest_priors<float, T, IN_BIAWGN>), grid, block, s, (const float *)in.src, prior, xbits, B, t.n, L.Bp, L.wpr, in.param);
        break;
    default:
        return fail(h, LDPC_EINVAL, "bad channel for MSA/SPA");
    }
    if (!*have_hard) {
        cudaError_t e = cudaMemsetAsync(xbits, 0, (size_t)t.n * L.wpr * 4, s);
        if (e != cudaSuccess) return fail(h, LDPC_ECUDA, std::string("memset xbits: ") + cudaGetErrorString(e));
    }
    return check_launch(h, "ingest");
}

struct PollState {
    int *d_flag;            // device
    int *h_flag;            // pinned host
};

int poll_any_active(ldpc_t *h, const PollState &ps, cudaStream_t s, bool *active)
{
    CUDA_TRY(h, cudaMemcpyAsync(ps.h_flag, ps.d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    *active = (*ps.h_flag != 0);
    return LDPC_OK;
}

}  // namespace

struct HostSlot {
    cudaStream_t stream = nullptr;
    void *d_y = nullptr; size_t y_bytes = 0;
    uint8_t *d_x = nullptr; size_t x_bytes = 0;
    uint8_t *d_yp = nullptr; size_t yp_bytes = 0;       // bit-packed input rows as they arrive (LDPC_IN_PACKED)
    uint8_t *d_xp = nullptr; size_t xp_bytes = 0;       // bit-packed words before they leave (LDPC_OUT_PACKED)
    int32_t *d_it = nullptr; uint8_t *d_rs = nullptr; size_t f_cap = 0;
    void *ws = nullptr; size_t ws_bytes = 0;
};

struct HostStage {
    static constexpr int kSlots = 3;
    HostSlot slot[kSlots];
    int next = 0;               // round-robin position, carried across calls so that back-to-back batches interleave
    int *h_flag = nullptr;      // pinned, used by the unlimited-iteration poll
};

namespace {

int ensure_stage(ldpc_t *h)
{
    if (h->stage) return LDPC_OK;
    HostStage *st = new (std::nothrow) HostStage();
    if (!st) return fail(h, LDPC_ENOMEM, "host stage");
    for (auto &sl : st->slot) {
        cudaError_t e = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete st; return fail(h, LDPC_ECUDA, std::string("stream create: ") + cudaGetErrorString(e)); }
    }
    cudaError_t e = cudaMallocHost((void **)&st->h_flag, 64);
    if (e != cudaSuccess) { delete st; return fail(h, LDPC_ECUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); }
    h->stage = st;
    return LDPC_OK;
}

template <typename U> int grow(ldpc_t *h, U **ptr, size_t *cap, size_t need)
{
    if (*cap >= need) return LDPC_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void **)ptr, need);
    if (e != cudaSuccess) return fail(h, LDPC_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cap = need;
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------------
// streaming BP decode
// ------------------------------------------------------------------------------------------------
template <typename T>
int decode_bp_stream(ldpc_t *h, int algo, const InSpec &in, int B, int max_iter, int iter_cap,
                     uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *marg_out,
                     void *ws, size_t ws_bytes, unsigned flags, cudaStream_t s)
{
    const Tables &t = h->t;
    const int dtype = sizeof(T) == 8 ? LDPC_F64 : LDPC_F32;
    const bool tma = (flags & LDPC_CN_REGISTER) == 0;
    const BpLayout L = bp_layout(t, dtype, B, marg_out != nullptr);
    if (ws_bytes < L.bytes) return fail(h, LDPC_EWORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(h, LDPC_EWORKSPACE, "workspace must be 256-byte aligned");
    const int limit = max_iter > 0 ? max_iter : iter_cap;
    if (limit <= 0) return fail(h, LDPC_EINVAL, "max_iter <= 0 (unlimited in the reference) needs iter_cap > 0");

    Carver cv(ws);
    BpParams<T> p;
    p.n = t.n; p.m = t.m; p.E = t.E;
    p.Bp = L.Bp; p.wpr = L.wpr; p.ngroups = L.ngroups;
    p.chk_ptr = t.chk_ptr; p.edge_var = t.edge_var; p.var_ptr = t.var_ptr; p.var_edges = t.var_edges;
    p.msg = cv.take<T>((size_t)t.E * L.Bp);
    T *prior = cv.take<T>((size_t)t.n * L.Bp);
    p.prior = prior;
    p.marg = marg_out ? cv.take<T>((size_t)t.n * L.Bp) : nullptr;
    p.xbits = cv.take<uint32_t>((size_t)t.n * L.wpr);
    p.act = cv.take<uint32_t>(L.wpr);
    p.unsat = cv.take<uint32_t>(L.wpr);
    p.iters = cv.take<int>(L.Bp);
    int *any_active = cv.take<int>(2);                              // [0] some frame runs, [1] how many
    int *src_of = cv.take<int>(L.Bp);
    int *orig = cv.take<int>(L.Bp);

    bool have_hard = false;
    int rc = ingest_bp<T>(h, in, prior, p.xbits, B, L, s, &have_hard);
    if (rc) return rc;
    if (p.marg) {                                                   // a frame that never runs reports marginal = prior
        CUDA_TRY(h, cudaMemcpyAsync(p.marg, prior, (size_t)t.n * L.Bp * sizeof(T), cudaMemcpyDeviceToDevice, s));
    }
    LAUNCH(h, init_flags, (L.Bp + 255) / 256, 256, s, p.act, p.unsat, p.iters, B, L.Bp, L.wpr, have_hard ? 0 : 1);

    // Long runs (iteration bound above 32, or "unlimited") read two words back every 4 iterations: whether any frame
    // still runs (to stop launching sweeps) and how many (to compact the live columns once half of them are done).
    const int TF = frames_per_tile(dtype), G = frames_per_group(dtype);
    const bool poll = (max_iter <= 0) || (limit > 32);
    const bool may_compact = poll && marg_out == nullptr && getenv("LDPC_NO_COMPACTION") == nullptr;
    // a read-back drains the stream (~20 us): every 2 iterations when an iteration moves >= 64 MB, every 4 otherwise
    const int poll_every = ((size_t)t.E * L.Bp * sizeof(T) >= ((size_t)64 << 20)) ? 2 : 4;
    int extent = L.Bp;                                              // frame columns the sweeps cover (a multiple of TF)
    bool compacted = false;
    PollState ps{any_active, nullptr};
    if (poll) {
        rc = ensure_stage(h);
        if (rc) return rc;
        ps.h_flag = h->stage->h_flag;
    }
    auto emit_all = [&]() -> int {                                  // words, iteration counts, reasons of the columns in `extent`
        const int wx = extent / 32;
        const int *og = compacted ? orig : nullptr;
        const dim3 egrid((t.n + 31) / 32, wx), eblock(32, 8);
        if (rows_word_aligned(x_hat, t.n)) LAUNCH(h, emit_words_tiled, tile_grid(t.n, wx), 256, s, p.xbits, (const uint32_t *)nullptr, x_hat, B, t.n, L.wpr, og, extent);
        else LAUNCH(h, emit_words, egrid, eblock, s, p.xbits, (const uint32_t *)nullptr, x_hat, B, t.n, L.wpr, og, extent);
        const int cols = compacted ? extent : B;
        LAUNCH(h, emit_status, (cols + 255) / 256, 256, s, p.iters, p.act, (const uint32_t *)nullptr, iters, reason, B,
               max_iter > 0 ? LDPC_REASON_MAXIMUM : LDPC_REASON_CAP, og, compacted ? extent : 0x7fffffff);
        return LDPC_OK;
    };

    for (int it = 0; it < limit; ++it) {
        p.first = (it == 0);
        p.skip_syn = (it == 0 && !have_hard);
        p.Bp = L.Bp; p.wpr = L.wpr;                                 // pitches stay; the extent shrinks with compaction
        p.ngroups = extent / G;
        ProfEvent *pe = prof_begin(h, 0, s);
        rc = launch_cn_algo<T>(h, algo, p, tma, s, extent);
        prof_end(pe, s);
        if (rc) return rc;
        const bool poll_now = poll && it >= 4 && (it % poll_every) == 0;
        if (poll_now) CUDA_TRY(h, cudaMemsetAsync(any_active, 0, 2 * sizeof(int), s));
        LAUNCH(h, bp_book, (extent + 255) / 256, 256, s, p.act, p.unsat, p.iters, extent / 32, any_active);
        int live = -1;
        if (poll_now) {
            CUDA_TRY(h, cudaMemcpyAsync(ps.h_flag, ps.d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(h, cudaStreamSynchronize(s));
            if (ps.h_flag[0] == 0) break;
            live = ps.h_flag[1];
        }
        int vpc = 1;
        const dim3 vgrid = sweep_grid(h, p.ngroups, t.n, &vpc);
        p.per_cta = vpc;
        pe = prof_begin(h, 1, s);
        rc = launch_vn<T>(h, p, vgrid, s);
        prof_end(pe, s);
        if (rc) return rc;
        // packing costs ~0.4 of an iteration's traffic and every later iteration saves 1 - live / extent of its own:
        // worth it once <= 70 % of the columns are live (and the tile-rounded extent really shrinks)
        if (may_compact && live > 0 && 10 * (long long)live <= 7 * (long long)extent && (live + TF - 1) / TF * TF < extent && it + 1 < limit) {
            // ---- active-frame compaction: retire what is done, pack the live columns to the front of every row
            if (!compacted) LAUNCH(h, iota_kernel, (L.Bp + 255) / 256, 256, s, orig, L.Bp);
            compacted = true;
            if ((rc = emit_all()) != 0) return rc;                  // finished frames leave now (live ones are rewritten at the end)
            LAUNCH(h, compact_plan, 1, 1024, s, p.act, extent / 32, src_of);
            LAUNCH(h, (compact_rows<T>), t.E, 256, s, p.msg, (size_t)L.Bp, src_of, live);
            LAUNCH(h, (compact_rows<T>), t.n, 256, s, prior, (size_t)L.Bp, src_of, live);
            LAUNCH(h, (compact_rows<int>), 1, 256, s, p.iters, (size_t)L.Bp, src_of, live);
            LAUNCH(h, (compact_rows<int>), 1, 256, s, orig, (size_t)L.Bp, src_of, live);
            LAUNCH(h, compact_bits, t.n, 128, s, p.xbits, L.wpr, src_of, live);
            extent = (live + TF - 1) / TF * TF;
            LAUNCH(h, compact_flags, (extent / 32 + 255) / 256, 256, s, p.act, p.unsat, orig, extent / 32, live);
        }
    }

    if ((rc = emit_all()) != 0) return rc;
    if (marg_out) {
        const dim3 tgrid((B + 31) / 32, (t.n + 31) / 32), eblock(32, 8);
        LAUNCH(h, (transpose_tile<T>), tgrid, eblock, s, p.marg, (T *)marg_out, t.n, B, L.Bp, t.n);
    }
    return check_launch(h, "decode_bp_stream");
}

// ------------------------------------------------------------------------------------------------
// streaming BEC decode
// ------------------------------------------------------------------------------------------------
int decode_bec_stream(ldpc_t *h, const uint8_t *y, int B, int max_iter, int iter_cap,
                      uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *ws, size_t ws_bytes, cudaStream_t s)
{
    const Tables &t = h->t;
    const BecLayout L = bec_layout(t, B);
    if (ws_bytes < L.bytes) return fail(h, LDPC_EWORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(h, LDPC_EWORKSPACE, "workspace must be 256-byte aligned");
    if (t.max_dv > 126) return fail(h, LDPC_EUNSUPPORTED, "variable degree > 126 is not supported by the BEC kernels");
    // Peeling ends by itself ('decoded' or 'stopping') after at most n rounds; that bounds "unlimited".
    const int limit = max_iter > 0 ? max_iter : (iter_cap > 0 ? iter_cap : t.n + 1);

    Carver cv(ws);
    BecParams p;
    p.n = t.n; p.m = t.m; p.E = t.E; p.wpr = L.wpr;
    p.chk_ptr = t.chk_ptr; p.edge_var = t.edge_var; p.var_ptr = t.var_ptr; p.var_edges = t.var_edges;
    p.mnz = cv.take<uint32_t>((size_t)t.E * L.wpr);
    p.mpos = cv.take<uint32_t>((size_t)t.E * L.wpr);
    uint32_t *pnz = cv.take<uint32_t>((size_t)t.n * L.wpr);
    uint32_t *ppos = cv.take<uint32_t>((size_t)t.n * L.wpr);
    p.pnz = pnz; p.ppos = ppos;
    p.xe = cv.take<uint32_t>((size_t)t.n * L.wpr);
    p.xv = cv.take<uint32_t>((size_t)t.n * L.wpr);
    p.act = cv.take<uint32_t>(L.wpr);
    p.changed = cv.take<uint32_t>(L.wpr);
    p.haser = cv.take<uint32_t>(L.wpr);
    p.stopped = cv.take<uint32_t>(L.wpr);
    p.iters = cv.take<int>(L.Bp);
    int *any_active = cv.take<int>(1);

    CUDA_TRY(h, cudaMemsetAsync(p.changed, 0, (size_t)((char *)p.iters - (char *)p.changed), s));   // changed, haser, stopped
    const dim3 igrid((t.n + 31) / 32, L.wpr), iblock(32, 8);
    if (rows_word_aligned(y, t.n)) LAUNCH(h, ingest_bec_tiled, tile_grid(t.n, L.wpr), 256, s, y, pnz, ppos, p.xe, p.xv, p.haser, B, t.n, L.wpr);
    else LAUNCH(h, ingest_bec, igrid, iblock, s, y, pnz, ppos, p.xe, p.xv, p.haser, B, t.n, L.wpr);
    LAUNCH(h, init_flags, (L.Bp + 255) / 256, 256, s, p.act, (uint32_t *)nullptr, p.iters, B, L.Bp, L.wpr, 0);

    // four words (128 frames) per thread once that still gives every SM two full loads of threads, else one word
    // (LDPC(1200,3,6): 96 against 80 M frames/s at 131072 frames, but 43 against 72 M at 32768)
    const bool wide = (long long)(L.wpr / 4) * std::min(t.m, t.n) >= (long long)h->sm_count * 4096;
    const int wx = wide ? L.wpr / 4 : L.wpr;
    const int gx = (wx + 127) / 128;
    const int target = h->sm_count * 32;
    auto grid_for = [&](int items, int *per) {
        int gy = std::max(1, std::min(items, target / gx));
        gy = std::min(gy, 65535);
        *per = (items + gy - 1) / gy;
        gy = (items + *per - 1) / *per;
        return dim3((unsigned)gx, (unsigned)gy, 1);
    };
    int cpc = 1, vpc = 1;
    const dim3 cgrid = grid_for(t.m, &cpc), vgrid = grid_for(t.n, &vpc);
    const int book_blocks = (L.wpr * 32 + 255) / 256;
    const bool poll = (max_iter <= 0) || (limit > 32);
    PollState ps{any_active, nullptr};
    if (poll) {
        int rc = ensure_stage(h);
        if (rc) return rc;
        ps.h_flag = h->stage->h_flag;
    }

    int it = 0;
    for (; it < limit; ++it) {
        const bool poll_now = poll && it >= 8 && (it % 8) == 0;
        if (poll_now) CUDA_TRY(h, cudaMemsetAsync(any_active, 0, sizeof(int), s));
        LAUNCH(h, bec_book, book_blocks, 256, s, p.act, p.changed, p.haser, p.stopped, p.iters, L.wpr, it == 0 ? 1 : 0, 0, any_active);
        if (poll_now) {
            bool active = true;
            int rc = poll_any_active(h, ps, s, &active);
            if (rc) return rc;
            if (!active) break;
        }
        p.first = (it == 0);
        p.per_cta = cpc;
        ProfEvent *pe = prof_begin(h, 0, s);
        if (wide) LAUNCH(h, bec_cn<W4>, cgrid, 128, s, p);
        else LAUNCH(h, bec_cn<uint32_t>, cgrid, 128, s, p);
        prof_end(pe, s);
        p.per_cta = vpc;
        pe = prof_begin(h, 1, s);
        if (wide) {
            if (t.max_dv <= 14) LAUNCH(h, (bec_vn<5, W4>), vgrid, 128, s, p);
            else LAUNCH(h, (bec_vn<8, W4>), vgrid, 128, s, p);
        } else {
            if (t.max_dv <= 14) LAUNCH(h, (bec_vn<5, uint32_t>), vgrid, 128, s, p);
            else LAUNCH(h, (bec_vn<8, uint32_t>), vgrid, 128, s, p);
        }
        prof_end(pe, s);
    }
    if (it == limit)    // account the last round; frames still active after it hit the bound
        LAUNCH(h, bec_book, book_blocks, 256, s, p.act, p.changed, p.haser, p.stopped, p.iters, L.wpr, 0, 1, any_active);

    if (rows_word_aligned(x_hat, t.n)) LAUNCH(h, emit_words_tiled, tile_grid(t.n, L.wpr), 256, s, p.xv, p.xe, x_hat, B, t.n, L.wpr);
    else LAUNCH(h, emit_words, igrid, iblock, s, p.xv, p.xe, x_hat, B, t.n, L.wpr);
    LAUNCH(h, emit_status, (B + 255) / 256, 256, s, p.iters, p.act, p.stopped, iters, reason, B,
           max_iter > 0 ? LDPC_REASON_MAXIMUM : LDPC_REASON_CAP);
    return check_launch(h, "decode_bec_stream");
}

// ------------------------------------------------------------------------------------------------
// resident (on-chip) BP decode for short codes
// ------------------------------------------------------------------------------------------------
struct ResLaunch {
    int grid, threads;
    size_t smem;
};

template <int Q, int ALGO, int DCP, bool UDC, int DVP, bool UDV, bool REGC, bool VSM>
int launch_resident_t(ldpc_t *h, const ResParams &rp, ResLaunch lc, int max_grid, cudaStream_t s)
{
    auto kern = resident_bp<Q, ALGO, DCP, UDC, DVP, UDV, REGC, VSM>;
    int per_sm = 1;
    int rc = resident_occupancy(h, kern, lc.threads, lc.smem, &per_sm);
    if (rc) return rc;
    const int grid = std::max(1, std::min(max_grid, h->sm_count * per_sm));
    kern<<<grid, lc.threads, lc.smem, s>>>(rp);
    h->launches++;
    return LDPC_OK;
}

template <int ALGO, int TT, int NPC, bool IRR = false, int MAXT = 320>
int launch_resident_vp(ldpc_t *h, const ResParams &rp, ResLaunch lc, int max_grid, cudaStream_t s)
{
    auto kern = resident_vp<ALGO, 6, IRR ? 8 : 3, TT, NPC, IRR, MAXT>;
    int per_sm = 1;
    int rc = resident_occupancy(h, kern, lc.threads, lc.smem, &per_sm);
    if (rc) return rc;
    const int grid = std::max(1, std::min(max_grid, h->sm_count * per_sm));
    kern<<<grid, lc.threads, lc.smem, s>>>(rp);
    h->launches++;
    return LDPC_OK;
}

template <int ALGO, int TT, int NPC, int INMODE = -1, int INES = -1, bool IRR = false, typename T = float, int MAXT = 320>
int launch_resident_vq(ldpc_t *h, const ResParams &rp, ResLaunch lc, int max_grid, cudaStream_t s)
{
    auto kern = resident_vq<ALGO, 6, IRR ? 8 : 3, TT, NPC, INMODE, INES, IRR, T, MAXT>;
    int per_sm = 1;
    int rc = resident_occupancy(h, kern, lc.threads, lc.smem, &per_sm);
    if (rc) return rc;
    const int grid = std::max(1, std::min(max_grid, h->sm_count * per_sm));
    kern<<<grid, lc.threads, lc.smem, s>>>(rp);
    h->launches++;
    return LDPC_OK;
}

template <int TT, int NPC, bool IRR>
int launch_resident_vd(ldpc_t *h, const ResParams &rp, ResLaunch lc, int max_grid, cudaStream_t s)
{
    auto kern = resident_vd<6, IRR ? 8 : 3, TT, NPC, IRR>;
    int per_sm = 1;
    int rc = resident_occupancy(h, kern, lc.threads, lc.smem, &per_sm);
    if (rc) return rc;
    const int grid = std::max(1, std::min(max_grid, h->sm_count * per_sm));
    kern<<<grid, lc.threads, lc.smem, s>>>(rp);
    h->launches++;
    return LDPC_OK;
}

// Shared memory one resident CTA may use: two CTAs share an SM (1 KB each is reserved by the system).
size_t resident_budget(const ldpc_t *h)
{
    return h->smem_per_sm / 2 > 4096 ? std::min(h->smem_optin, h->smem_per_sm / 2 - 1536) : 0;
}

// float32: MSA / SPA on any on-chip kernel; float64 (the reference's own arithmetic): min-sum on the two-CTA
// variable-plane geometries, regular or irregular (resident_vd.cuh).
bool resident_eligible(const ldpc_t *h, int algo, int dtype, const void *marg_out)
{
    if (!h->res.ok || marg_out != nullptr) return false;
    if (dtype == LDPC_F32) return algo == LDPC_MSA || algo == LDPC_SPA;
    return dtype == LDPC_F64 && algo == LDPC_MSA && ((h->res.vp && !h->res.vp_big) || h->res.vx);
}

int decode_bp_resident(ldpc_t *h, int algo, int dtype, const InSpec &in, int B, int max_iter, int iter_cap,
                       uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *ws, size_t ws_bytes,
                       unsigned flags, cudaStream_t s)
{
    const Tables &t = h->t;
    const ResidentInfo &r = h->res;
    constexpr int Q = 1;
    if (ws_bytes < 256) return fail(h, LDPC_EWORKSPACE, "workspace too small");
    const int limit = max_iter > 0 ? max_iter : iter_cap;
    if (limit <= 0) return fail(h, LDPC_EINVAL, "max_iter <= 0 (unlimited in the reference) needs iter_cap > 0");
    const int tb = (algo == LDPC_MSA) ? 0 : 1;              // sum-product keeps the natural edge order inside a check
    ResParams rp;
    rp.n = r.np; rp.m = r.mp; rp.nref = t.n; rp.planes = r.planes;
    rp.cvar = r.cvar[tb]; rp.vrow = r.vrow[tb]; rp.cdeg = r.cdeg; rp.vdeg = r.vdeg;
    rp.vposmap = r.vposmap; rp.vinvmap = r.vinvmap;
    rp.cw = nullptr;
    rp.cwx = nullptr;
    rp.plane_cells = 0;
    for (int k = 0; k < 8; ++k) { rp.pcnt[k] = 0; rp.pbase[k] = 0; }
    if (r.vp) { rp.cw = r.vp_cw[tb]; rp.vposmap = r.vp_vposmap[tb]; rp.vinvmap = r.vp_vinvmap[tb]; }
    if (r.vx) {
        rp.cwx = r.vx_cwx[tb]; rp.vposmap = r.vp_vposmap[tb]; rp.vinvmap = r.vp_vinvmap[tb];
        rp.plane_cells = r.vx_cells[tb];
        for (int k = 0; k < 8; ++k) { rp.pcnt[k] = r.vx_pcnt[tb][k]; rp.pbase[k] = r.vx_pbase[tb][k]; }
    }
    rp.cn_items = r.mp * Q; rp.vn_items = r.np * Q;
    rp.src = in.src;
    rp.y_hard = in.y_hard;
    rp.param = in.param;
    rp.inv_param = (in.param != 0.0) ? 1.0 / in.param : 0.0;
    switch (in.channel) {
    case LDPC_CH_PRIORS: rp.in_mode = IN_COPY; rp.in_es = (int)in_elem_size(in.in_dtype); break;
    case LDPC_CH_BSC: rp.in_mode = IN_BSC; rp.in_es = 1; break;
    case LDPC_CH_BIAWGN: rp.in_mode = IN_BIAWGN; rp.in_es = (int)in_elem_size(in.in_dtype); break;
    default: return fail(h, LDPC_EINVAL, "bad channel for MSA/SPA");
    }
    rp.B = B; rp.limit = limit;
    rp.bound_reason = max_iter > 0 ? LDPC_REASON_MAXIMUM : LDPC_REASON_CAP;
    rp.sat_llr = (flags & LDPC_SPA_ROBUST) ? INFINITY : kSpaSatLlr;
    rp.x_hat = x_hat; rp.iters = iters; rp.reason = reason;
    rp.counter = static_cast<int *>(ws);

    // Ring of received rows landed by the bulk-copy engine: needs 16-byte aligned rows; as deep as shared memory allows.
    const size_t row_bytes = (size_t)t.n * rp.in_es;
    const size_t stride = align_up(row_bytes, 16);
    const size_t budget = r.vp_big ? h->smem_optin - 1024 : resident_budget(h);
    const int vtw = r.regular36 ? 2 : 0;                     // (3,6) variant keeps the variable-edge table in shared memory
    const size_t state = r.vx ? vx_smem_layout(r.np, r.vx_cells[tb], 0, 0).total
                       : r.vp ? vp_smem_layout(r.np, 3, 0, 0, !r.vp_big).total : resident_smem_layout(Q, r.np, r.mp, r.planes, vtw, 0, 0).total;
    int ring = 0;
    if (row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(in.src) & 15u) == 0 && budget > state)
        ring = (int)std::min<size_t>(kResRingMax, (budget - state) / stride);
    // regular codes, two-CTA geometry, float32, rows the bulk copy can stage, no separate hard input: the kernel with the
    // frame hand-over fused into its variable phase (LDPC_RESIDENT_VP=1 keeps resident_vp, for A/B runs)
    // (float64 min-sum runs the same kernel on double2 cells; float64 ROWS leave room for one ring entry only, which is
    // enough as long as one frame leaves per iteration - a second one waits for the next variable phase)
    // (the one-CTA-per-SM geometry of codes up to n ~ 2850 has room for one ring entry too)
    const bool use_vq = ((r.vp && r.regular36 && (!r.vp_big || dtype == LDPC_F32)) || r.vx) &&
                        ring >= ((dtype == LDPC_F64 || r.vp_big) ? 1 : 2) && in.y_hard == nullptr && getenv("LDPC_RESIDENT_VP") == nullptr;
    if (use_vq) ring = std::min(ring, kVqRing);
    rp.ring = ring;
    rp.stage_stride = (int)stride;
    ResLaunch lc;
    lc.threads = r.threads;
    lc.smem = r.vx ? vx_smem_layout(r.np, r.vx_cells[tb], ring, (int)stride).total
            : r.vp ? vp_smem_layout(r.np, 3, ring, (int)stride, !r.vp_big).total
                   : resident_smem_layout(Q, r.np, r.mp, r.planes, vtw, ring, (int)stride).total;
    const int slots = (dtype == LDPC_F64) ? 2 : 4 * Q;           // frames a CTA holds
    const int max_grid = (B + slots - 1) / slots;

    CUDA_TRY(h, cudaMemsetAsync(rp.counter, 0, sizeof(int), s));
    ProfEvent *pe = prof_begin(h, 0, s);
    int rc;
    if (dtype == LDPC_F64 && use_vq) {                              // float64 min-sum with the fused hand-over
        const bool ens = lc.threads == 320 && r.np == 1200 && r.mp == 600 && t.n == 1200;
        if (r.vx) rc = ens ? launch_resident_vq<ALGO_MSA, 320, 1200, -1, -1, true, double>(h, rp, lc, max_grid, s)
                           : launch_resident_vq<ALGO_MSA, 0, 0, -1, -1, true, double>(h, rp, lc, max_grid, s);
        else if (ens && rp.in_mode == IN_BIAWGN && rp.in_es == 8) rc = launch_resident_vq<ALGO_MSA, 320, 1200, IN_BIAWGN, 8, false, double>(h, rp, lc, max_grid, s);
        else if (ens && rp.in_mode == IN_BIAWGN && rp.in_es == 4) rc = launch_resident_vq<ALGO_MSA, 320, 1200, IN_BIAWGN, 4, false, double>(h, rp, lc, max_grid, s);
        else rc = ens ? launch_resident_vq<ALGO_MSA, 320, 1200, -1, -1, false, double>(h, rp, lc, max_grid, s)
                      : launch_resident_vq<ALGO_MSA, 0, 0, -1, -1, false, double>(h, rp, lc, max_grid, s);
    } else if (dtype == LDPC_F64) {                                 // resident_eligible: min-sum, r.vp, two CTAs per SM
        const bool ens = lc.threads == 320 && r.np == 1200 && r.mp == 600 && t.n == 1200;
        if (r.vx) rc = ens ? launch_resident_vd<320, 1200, true>(h, rp, lc, max_grid, s) : launch_resident_vd<0, 0, true>(h, rp, lc, max_grid, s);
        else rc = ens ? launch_resident_vd<320, 1200, false>(h, rp, lc, max_grid, s) : launch_resident_vd<0, 0, false>(h, rp, lc, max_grid, s);
    } else if (r.vx && use_vq) {                                      // irregular instance, hand-over fused (resident_vq.cuh)
        const bool ens = lc.threads == 320 && r.np == 1200 && r.mp == 600 && t.n == 1200;
        if (ens && rp.in_mode == IN_BSC)                                           // config 4: the irregular ensemble on the BSC
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 320, 1200, IN_BSC, 1, true>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 320, 1200, IN_BSC, 1, true>(h, rp, lc, max_grid, s);
        else if (ens)
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 320, 1200, -1, -1, true>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 320, 1200, -1, -1, true>(h, rp, lc, max_grid, s);
        else
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 0, 0, -1, -1, true>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 0, 0, -1, -1, true>(h, rp, lc, max_grid, s);
    } else if (r.vx) {
        if (lc.threads == 320 && r.np == 1200 && r.mp == 600 && t.n == 1200)      // the reference's irregular n = 1200 ensemble
            rc = (algo == LDPC_MSA) ? launch_resident_vp<ALGO_MSA, 320, 1200, true>(h, rp, lc, max_grid, s)
                                    : launch_resident_vp<ALGO_SPA_PHI, 320, 1200, true>(h, rp, lc, max_grid, s);
        else
            rc = (algo == LDPC_MSA) ? launch_resident_vp<ALGO_MSA, 0, 0, true>(h, rp, lc, max_grid, s)
                                    : launch_resident_vp<ALGO_SPA_PHI, 0, 0, true>(h, rp, lc, max_grid, s);
    } else if (r.vp_big && use_vq) {
        rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 0, 0, -1, -1, false, float, kVpBigThreads>(h, rp, lc, max_grid, s)
                                : launch_resident_vq<ALGO_SPA_PHI, 0, 0, -1, -1, false, float, kVpBigThreads>(h, rp, lc, max_grid, s);
    } else if (r.vp_big) {
        rc = (algo == LDPC_MSA) ? launch_resident_vp<ALGO_MSA, 0, 0, false, kVpBigThreads>(h, rp, lc, max_grid, s)
                                : launch_resident_vp<ALGO_SPA_PHI, 0, 0, false, kVpBigThreads>(h, rp, lc, max_grid, s);
    } else if (r.vp && use_vq) {                                      // frame hand-over fused into the variable phase (resident_vq.cuh)
        const bool ens = lc.threads == 320 && r.np == 1200 && r.mp == 600 && t.n == 1200;      // the reference's (1200,3,6) ensemble
        if (ens && rp.in_mode == IN_BIAWGN && rp.in_es == 4)                                   // ... on float32 BIAWGN rows: the headline
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 320, 1200, IN_BIAWGN, 4>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 320, 1200, IN_BIAWGN, 4>(h, rp, lc, max_grid, s);
        else if (ens && rp.in_mode == IN_BSC)
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 320, 1200, IN_BSC, 1>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 320, 1200, IN_BSC, 1>(h, rp, lc, max_grid, s);
        else if (ens)
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 320, 1200>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 320, 1200>(h, rp, lc, max_grid, s);
        else
            rc = (algo == LDPC_MSA) ? launch_resident_vq<ALGO_MSA, 0, 0>(h, rp, lc, max_grid, s)
                                    : launch_resident_vq<ALGO_SPA_PHI, 0, 0>(h, rp, lc, max_grid, s);
    } else if (r.vp) {
        if (lc.threads == 320 && r.np == 1200 && r.mp == 600 && t.n == 1200)      // the reference's (1200,3,6) ensemble
            rc = (algo == LDPC_MSA) ? launch_resident_vp<ALGO_MSA, 320, 1200>(h, rp, lc, max_grid, s)
                                    : launch_resident_vp<ALGO_SPA_PHI, 320, 1200>(h, rp, lc, max_grid, s);
        else if (lc.threads == 320)
            rc = (algo == LDPC_MSA) ? launch_resident_vp<ALGO_MSA, 320, 0>(h, rp, lc, max_grid, s)
                                    : launch_resident_vp<ALGO_SPA_PHI, 320, 0>(h, rp, lc, max_grid, s);
        else
            rc = (algo == LDPC_MSA) ? launch_resident_vp<ALGO_MSA, 0, 0>(h, rp, lc, max_grid, s)
                                    : launch_resident_vp<ALGO_SPA_PHI, 0, 0>(h, rp, lc, max_grid, s);
    } else if (r.regular36)
        rc = (algo == LDPC_MSA) ? launch_resident_t<Q, ALGO_MSA, 6, true, 3, true, true, true>(h, rp, lc, max_grid, s)
                                : launch_resident_t<Q, ALGO_SPA_PHI, 6, true, 3, true, true, true>(h, rp, lc, max_grid, s);
    else
        rc = (algo == LDPC_MSA) ? launch_resident_t<Q, ALGO_MSA, 8, false, 8, false, false, false>(h, rp, lc, max_grid, s)
                                : launch_resident_t<Q, ALGO_SPA_PHI, 8, false, 8, false, false, false>(h, rp, lc, max_grid, s);
    prof_end(pe, s);
    if (rc) return rc;
    return check_launch(h, "decode_bp_resident");
}

// ------------------------------------------------------------------------------------------------
// resident (on-chip) BEC decode for short codes (resident_bec.cuh)
// ------------------------------------------------------------------------------------------------
// The erasure kernel shares the variable-plane tables of resident_vp (two-CTA geometry): regular (3,6) codes and the
// irregular instance (check degrees 2..6, variable degrees 0..8).
bool bec_resident_eligible(const ldpc_t *h)
{
    const ResidentInfo &r = h->res;
    if (!r.ok || !((r.vp && !r.vp_big) || r.vx)) return false;
    const BecSmem L = bec_smem_layout(r.np, r.vx ? r.vx_cells[0] : 0, r.vx, h->t.n);
    return L.total <= resident_budget(h) && (!r.vx || r.vx_vdeg != nullptr);
}

int decode_bec_resident(ldpc_t *h, const uint8_t *y, int B, int max_iter, int iter_cap,
                        uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *ws, size_t ws_bytes, cudaStream_t s)
{
    const Tables &t = h->t;
    const ResidentInfo &r = h->res;
    if (ws_bytes < 256) return fail(h, LDPC_EWORKSPACE, "workspace too small");
    BecResParams bp;
    bp.np = r.np; bp.mp = r.mp; bp.nref = t.n;
    bp.cw = r.vp ? r.vp_cw[0] : nullptr;
    bp.cwx = r.vx ? r.vx_cwx[0] : nullptr;
    bp.vposmap = r.vp_vposmap[0];
    bp.vdeg = r.vx_vdeg;
    bp.plane_cells = r.vx ? r.vx_cells[0] : 0;
    for (int k = 0; k < 8; ++k) { bp.pcnt[k] = r.vx ? r.vx_pcnt[0][k] : 0; bp.pbase[k] = r.vx ? r.vx_pbase[0][k] : 0; }
    bp.y = y; bp.B = B;
    // peeling ends by itself ('decoded' or 'stopping') after at most n rounds; that bounds "unlimited"
    bp.limit = max_iter > 0 ? max_iter : (iter_cap > 0 ? iter_cap : t.n + 1);
    bp.bound_reason = max_iter > 0 ? LDPC_REASON_MAXIMUM : LDPC_REASON_CAP;
    bp.x_hat = x_hat; bp.iters = iters; bp.reason = reason;
    bp.counter = static_cast<int *>(ws);
    const BecSmem L = bec_smem_layout(r.np, bp.plane_cells, r.vx, t.n);
    const int tiles = (B + 63) / 64;
    CUDA_TRY(h, cudaMemsetAsync(bp.counter, 0, sizeof(int), s));
    ProfEvent *pe = prof_begin(h, 0, s);
    int per_sm = 1, rc;
    // wide geometry: one check and two variables per thread when the code allows it (n = 1200: 608 threads)
    const int wide_t = (std::max(r.mp, (r.np + 1) / 2) + 31) / 32 * 32;
    // measured (r2): 318 M frames/s wide against 356 M for resident_vp's own geometry on config 2 — 48 registers starve the
    // transposes and add instructions; kept behind LDPC_BEC_WIDE=1 for A/B runs (regular codes only: the irregular instance spills)
    const bool wide = !r.vx && wide_t <= 608 && wide_t > r.threads && getenv("LDPC_BEC_WIDE") != nullptr;
    const int threads = wide ? wide_t : r.threads;
#define BEC_LAUNCH(...)                                                                             \
    do {                                                                                            \
        auto kern = resident_bec<__VA_ARGS__>;                                                      \
        if ((rc = resident_occupancy(h, kern, threads, L.total, &per_sm)) != 0) return rc;          \
        kern<<<std::max(1, std::min(tiles, h->sm_count * per_sm)), threads, L.total, s>>>(bp);      \
    } while (0)
    if (r.vx) BEC_LAUNCH(true);
    else if (wide) BEC_LAUNCH(false, 1, 2, 608);
    else BEC_LAUNCH(false);
#undef BEC_LAUNCH
    h->launches++;
    prof_end(pe, s);
    return check_launch(h, "decode_bec_resident");
}

// Places the graph in shared memory (res_layout.h), builds the position-indexed uint16 tables and the geometry of
// the resident kernel; leaves res.ok = false when the code does not fit.
int build_resident(ldpc_t *h, const int32_t *chk_ptr, const int32_t *edge_var, const int32_t *var_ptr, const int32_t *var_edges)
{
    const Tables &t = h->t;
    ResidentInfo &r = h->res;
    constexpr int Q = 1;
    if (t.max_dc > 8 || t.max_dv > 8 || t.max_dc < 1) return LDPC_OK;
    const int G = 8 / Q;
    const int mp = (t.m + G - 1) / G * G, np = (t.n + G - 1) / G * G;
    r.planes = t.max_dc;
    r.regular36 = (t.uni_dc == 6 && t.uni_dv == 3 && mp == t.m && np == t.n);
    const int vtw = r.regular36 ? 2 : 0;
    const char *lay = getenv("LDPC_RESIDENT_LAYOUT");
    const bool lay_check = lay && std::string(lay) == "check";
    // Regular (3,6) codes too long for half an SM (Margulis, n = 2640): ONE variable-plane CTA per SM, all of its shared memory
    const bool big = r.regular36 && !lay_check && np % 4 == 0 && np * 16 <= 0xfff0 &&
                     (vp_smem_layout(np, 3, 0, 0).total > resident_budget(h) ||                      // too much state for half an SM,
                      mp > kResCnPasses * res_max_threads(Q) || np > kResVnPasses * res_max_threads(Q)) &&   // or too many items for 320 threads

                     vp_smem_layout(np, 3, 0, 0, false).total <= h->smem_optin - 1024 &&
                     mp <= (kResCnPasses + 1) * kVpBigThreads && np <= kVpBigVnPasses * kVpBigThreads;
    // the check-major kernel (resident_bp) needs more shared memory than the variable-plane layouts (its edge table):
    // it is only the fallback, so its limits must not keep a code off the variable-plane kernels (n = 1280 fits those only)
    const bool bp_fits = (long long)r.planes * mp * Q + 1 <= 65535 &&                 // c2v float4 index must fit 16 bits
                         ((long long)np * Q + 1) * 16 <= 65535 &&                     // marg byte offset must fit 16 bits
                         resident_smem_layout(Q, np, mp, r.planes, vtw, 0, 0).total <= resident_budget(h);
    const bool vp_cand = r.regular36 && !lay_check && np % 4 == 0 && np * 16 <= 0xfff0 &&
                         (big || vp_smem_layout(np, 3, 0, 0).total <= resident_budget(h));
    const bool vx_cand = !r.regular36 && !lay_check && t.max_dc <= 6 && t.max_dv <= 8;
    if (!bp_fits && !vp_cand && !vx_cand) return LDPC_OK;
    // threads: check items (mp * Q) in at most 2 passes, variable items (np * Q) in at most 4
    const int citems = mp * Q, vitems = np * Q, maxT = big ? kVpBigThreads : res_max_threads(Q);
    int T = std::max((citems + kResCnPasses - 1) / kResCnPasses, (vitems + kResVnPasses - 1) / kResVnPasses);
    if (citems <= maxT && vitems <= 2 * maxT) T = std::max(citems, (vitems + 1) / 2);
    T = std::max(64, (T + 31) / 32 * 32);
    if (big) T = kVpBigThreads;                                   // third check pass / fifth variable pass cover the rest
    if (T > maxT) return LDPC_OK;
    r.threads = T; r.Q = Q; r.np = np; r.mp = mp;

    std::vector<int> edge_chk((size_t)t.E);
    for (int c = 0; c < t.m; ++c)
        for (int e = chk_ptr[c]; e < chk_ptr[c + 1]; ++e) edge_chk[e] = c;
    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaSuccess;

    // ---- variable-plane variant for regular (3,6) codes: its own placement per edge order (res_layout.h, vn_contiguous)
    if (vp_cand) {
        std::vector<int> slot((size_t)t.E, 0);
        for (int v = 0; v < t.n; ++v)
            for (int p0 = var_ptr[v], k = 0; p0 < var_ptr[v + 1]; ++p0, ++k) slot[var_edges[p0]] = k;
        for (int tb = 0; tb < 2 && e == cudaSuccess; ++tb) {
            ResPlanner pl2(t.n, t.m, t.E, chk_ptr, edge_var, var_ptr, var_edges, G);
            const ResLayout V = pl2.plan(12345u, h->plan_effort, true, tb == 1);
            std::vector<uint16_t> cw((size_t)mp * 8, 0), vpm((size_t)t.n), vim((size_t)np, 0xffffu);
            for (int ed = 0; ed < t.E; ++ed)
                cw[(size_t)V.cpos[edge_chk[ed]] * 8 + V.eord[ed]] = (uint16_t)((V.vpos[edge_var[ed]] << 4) | (slot[ed] + 1));
            for (int v = 0; v < t.n; ++v) { vpm[v] = (uint16_t)V.vpos[v]; vim[V.vpos[v]] = (uint16_t)v; }
            if ((e = up((void **)&r.vp_cw[tb], cw.data(), cw.size() * 2)) != cudaSuccess) break;
            if ((e = up((void **)&r.vp_vposmap[tb], vpm.data(), vpm.size() * 2)) != cudaSuccess) break;
            if ((e = up((void **)&r.vp_vinvmap[tb], vim.data(), vim.size() * 2)) != cudaSuccess) break;
            if (tb == 0) { r.plan[0] = V.cn_ideal; r.plan[1] = V.cn_file; r.plan[3] = V.cn_plan; r.plan[4] = V.vn_ideal; r.plan[5] = V.vn_file; r.plan[6] = V.vn_plan; }
            else r.plan[2] = V.cn_plan;
        }
        if (e != cudaSuccess) return fail(nullptr, LDPC_ECUDA, std::string("resident table upload: ") + cudaGetErrorString(e));
        r.vp = true;
        r.vp_big = big;
        r.ok = true;
        return LDPC_OK;
    }

    // ---- the same layout for irregular codes (resident_vp IRR = true): check degrees 2..6, variable degrees 0..8
    if (vx_cand) {
        bool fits = true;
        for (int tb = 0; tb < 2 && fits && e == cudaSuccess; ++tb) {
            ResPlanner pl2(t.n, t.m, t.E, chk_ptr, edge_var, var_ptr, var_edges, G);
            const ResLayout V = pl2.plan(12345u, h->plan_effort, true, tb == 1);
            VxTables X;
            if (!build_vx_tables(V, t.n, t.m, t.E, chk_ptr, edge_var, var_ptr, var_edges, 6, &X) ||
                vx_smem_layout(np, X.plane_cells, 0, 0).total > resident_budget(h)) { fits = false; break; }
            if ((e = up((void **)&r.vx_cwx[tb], X.cwx.data(), X.cwx.size() * 4)) != cudaSuccess) break;
            if ((e = up((void **)&r.vp_vposmap[tb], X.vposmap.data(), X.vposmap.size() * 2)) != cudaSuccess) break;
            if ((e = up((void **)&r.vp_vinvmap[tb], X.vinvmap.data(), X.vinvmap.size() * 2)) != cudaSuccess) break;
            r.vx_cells[tb] = X.plane_cells;
            for (int k = 0; k < 8; ++k) { r.vx_pcnt[tb][k] = X.pcnt[k]; r.vx_pbase[tb][k] = X.pbase[k]; }
            if (tb == 0) {                                            // degree by position, for the on-chip erasure kernel
                std::vector<uint8_t> vd((size_t)np, 0xffu);
                for (int v = 0; v < t.n; ++v) vd[X.vposmap[v]] = (uint8_t)(var_ptr[v + 1] - var_ptr[v]);
                if ((e = up((void **)&r.vx_vdeg, vd.data(), vd.size())) != cudaSuccess) break;
            }
            if (tb == 0) { r.plan[0] = V.cn_ideal; r.plan[1] = V.cn_file; r.plan[3] = V.cn_plan; r.plan[4] = V.vn_ideal; r.plan[5] = V.vn_file; r.plan[6] = V.vn_plan; }
            else r.plan[2] = V.cn_plan;
        }
        if (e != cudaSuccess) return fail(nullptr, LDPC_ECUDA, std::string("resident table upload: ") + cudaGetErrorString(e));
        if (fits) {
            r.vx = true;
            r.ok = true;
            return LDPC_OK;
        }
        for (int tb = 0; tb < 2; ++tb) {                              // does not fit: the check-major kernel below
            if (r.vx_cwx[tb]) { cudaFree(r.vx_cwx[tb]); r.vx_cwx[tb] = nullptr; }
            if (r.vp_vposmap[tb]) { cudaFree(r.vp_vposmap[tb]); r.vp_vposmap[tb] = nullptr; }
            if (r.vp_vinvmap[tb]) { cudaFree(r.vp_vinvmap[tb]); r.vp_vinvmap[tb] = nullptr; }
        }
        if (r.vx_vdeg) { cudaFree(r.vx_vdeg); r.vx_vdeg = nullptr; }
    }

    if (!bp_fits) return LDPC_OK;
    ResPlanner planner(t.n, t.m, t.E, chk_ptr, edge_var, var_ptr, var_edges, G);
    const ResLayout L = planner.plan(12345u, h->plan_effort);
    const long pl[7] = {L.cn_ideal, L.cn_file, L.cn_plan_natural, L.cn_plan, L.vn_ideal, L.vn_file, L.vn_plan};
    std::copy(pl, pl + 7, r.plan);

    std::vector<uint8_t> cdeg((size_t)mp, 0), vdeg((size_t)np, 0);
    std::vector<uint16_t> vposmap((size_t)t.n), vinvmap((size_t)np, 0xffffu);
    for (int c = 0; c < t.m; ++c) cdeg[L.cpos[c]] = (uint8_t)(chk_ptr[c + 1] - chk_ptr[c]);
    for (int v = 0; v < t.n; ++v) {
        vdeg[L.vpos[v]] = (uint8_t)(var_ptr[v + 1] - var_ptr[v]);
        vposmap[v] = (uint16_t)L.vpos[v];
        vinvmap[L.vpos[v]] = (uint16_t)v;
    }
    for (int tb = 0; tb < 2 && e == cudaSuccess; ++tb) {
        const std::vector<uint8_t> &plane = tb == 0 ? L.eord : L.enat;
        std::vector<uint16_t> cvar((size_t)mp * 8, (uint16_t)np), vrow((size_t)np * 8, 0);
        for (int ed = 0; ed < t.E; ++ed)
            cvar[(size_t)L.cpos[edge_chk[ed]] * 8 + plane[ed]] = (uint16_t)L.vpos[edge_var[ed]];
        for (int v = 0; v < t.n; ++v)
            for (int p0 = var_ptr[v], k = 0; p0 < var_ptr[v + 1]; ++p0, ++k) {
                const int ed = var_edges[p0];
                vrow[(size_t)L.vpos[v] * 8 + k] = (uint16_t)(plane[ed] * mp + L.cpos[edge_chk[ed]]);
            }
        if ((e = up((void **)&r.cvar[tb], cvar.data(), cvar.size() * 2)) != cudaSuccess) break;
        e = up((void **)&r.vrow[tb], vrow.data(), vrow.size() * 2);
    }
    if (e != cudaSuccess || (e = up((void **)&r.cdeg, cdeg.data(), cdeg.size())) != cudaSuccess ||
        (e = up((void **)&r.vdeg, vdeg.data(), vdeg.size())) != cudaSuccess ||
        (e = up((void **)&r.vposmap, vposmap.data(), vposmap.size() * 2)) != cudaSuccess ||
        (e = up((void **)&r.vinvmap, vinvmap.data(), vinvmap.size() * 2)) != cudaSuccess)
        return fail(nullptr, LDPC_ECUDA, std::string("resident table upload: ") + cudaGetErrorString(e));
    r.ok = true;
    return LDPC_OK;
}

int decode_any(ldpc_t *h, int algo, int dtype, const InSpec &in, int B, int max_iter, int iter_cap,
               uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *marg_out,
               void *ws, size_t ws_bytes, unsigned flags, cudaStream_t s)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0) return fail(h, LDPC_EINVAL, "B must be positive");
    if (B > kMaxFramesPerCall) return fail(h, LDPC_EINVAL, "B too large: at most 2096640 frames per call (grid limits of the layout kernels); split the batch");
    if (!in.src || !x_hat || !iters || !ws) return fail(h, LDPC_EINVAL, "null buffer");
    const unsigned path = flags & LDPC_PATH_MASK;
    ENTER(h);
    if (algo == LDPC_BEC) {
        if (in.channel != LDPC_CH_BEC) return fail(h, LDPC_EINVAL, "BEC decoder needs symbol input");
        if (marg_out) return fail(h, LDPC_EINVAL, "marg_out is MSA/SPA only");
        const bool bec_res = bec_resident_eligible(h);
        if (path == LDPC_PATH_RESIDENT && !bec_res)
            return fail(h, LDPC_EUNSUPPORTED, "the on-chip erasure kernel needs a code on the variable-plane tables (two-CTA geometry: regular (3,6) or check degrees 2..6 / variable degrees <= 8, n up to ~1280)");
        if (path == LDPC_PATH_RESIDENT || (path == LDPC_PATH_AUTO && bec_res))
            return decode_bec_resident(h, (const uint8_t *)in.src, B, max_iter, iter_cap, x_hat, iters, reason, ws, ws_bytes, s);
        return decode_bec_stream(h, (const uint8_t *)in.src, B, max_iter, iter_cap, x_hat, iters, reason, ws, ws_bytes, s);
    }
    if (algo != LDPC_MSA && algo != LDPC_SPA) return fail(h, LDPC_EINVAL, "bad algo");
    if (in.channel == LDPC_CH_BEC) return fail(h, LDPC_EINVAL, "MSA/SPA cannot take BEC symbols");
    const bool res_ok = resident_eligible(h, algo, dtype, marg_out);
    if (path == LDPC_PATH_RESIDENT && !res_ok)
        return fail(h, LDPC_EUNSUPPORTED, "resident path needs float32 MSA/SPA (or float64 MSA on a code of n <= 1280 with check degrees <= 6), no marg_out, degrees <= 8 and a code that fits in shared memory");
    if (path == LDPC_PATH_RESIDENT || (path == LDPC_PATH_AUTO && res_ok))
        return decode_bp_resident(h, algo, dtype, in, B, max_iter, iter_cap, x_hat, iters, reason, ws, ws_bytes, flags, s);
    if (dtype == LDPC_F32)
        return decode_bp_stream<float>(h, algo, in, B, max_iter, iter_cap, x_hat, iters, reason, marg_out, ws, ws_bytes, flags, s);
    if (dtype == LDPC_F64)
        return decode_bp_stream<double>(h, algo, in, B, max_iter, iter_cap, x_hat, iters, reason, marg_out, ws, ws_bytes, flags, s);
    return fail(h, LDPC_EINVAL, "bad dtype");
}

size_t workspace_bytes_impl(const ldpc_t *h, int algo, int dtype, int B, bool want_marg)
{
    if (!h || B <= 0) return 0;
    if (algo == LDPC_BEC) return bec_layout(h->t, B).bytes;
    if (algo != LDPC_MSA && algo != LDPC_SPA) return 0;
    if (dtype != LDPC_F32 && dtype != LDPC_F64) return 0;
    return bp_layout(h->t, dtype, B, want_marg).bytes;
}

}  // namespace

// One isolated sweep in the reference layout [B][E] (teacher-forced parity).
template <typename T>
static int debug_step_t(ldpc_t *h, int algo, int which, int B, const void *prior_in, const void *msg_in,
                        void *msg_out, void *marg_out, void *ws, size_t ws_bytes, cudaStream_t s)
{
    const Tables &t = h->t;
    const int dtype = sizeof(T) == 8 ? LDPC_F64 : LDPC_F32;
    const BpLayout L = bp_layout(t, dtype, B, true);
    if (ws_bytes < L.bytes) return fail(h, LDPC_EWORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(h, LDPC_EWORKSPACE, "workspace must be 256-byte aligned");
    Carver cv(ws);
    BpParams<T> p;
    p.n = t.n; p.m = t.m; p.E = t.E;
    p.Bp = L.Bp; p.wpr = L.wpr; p.ngroups = L.ngroups;
    p.chk_ptr = t.chk_ptr; p.edge_var = t.edge_var; p.var_ptr = t.var_ptr; p.var_edges = t.var_edges;
    p.msg = cv.take<T>((size_t)t.E * L.Bp);
    T *prior = cv.take<T>((size_t)t.n * L.Bp);
    p.prior = prior;
    p.marg = cv.take<T>((size_t)t.n * L.Bp);
    p.xbits = cv.take<uint32_t>((size_t)t.n * L.wpr);
    p.act = cv.take<uint32_t>(L.wpr);
    p.unsat = cv.take<uint32_t>(L.wpr);
    p.iters = cv.take<int>(L.Bp);
    p.first = 0; p.skip_syn = 1;

    const dim3 block(32, 8);
    CUDA_TRY(h, cudaMemsetAsync(ws, 0, L.bytes, s));
    LAUNCH(h, (transpose_tile<T>), dim3((t.E + 31) / 32, (B + 31) / 32), block, s, (const T *)msg_in, p.msg, B, t.E, t.E, L.Bp);
    LAUNCH(h, init_flags, (L.Bp + 255) / 256, 256, s, p.act, p.unsat, p.iters, B, L.Bp, L.wpr, 1);
    int rc;
    if (which == 0) {
        rc = launch_cn_algo<T>(h, algo, p, true, s, L.Bp);
        if (rc) return rc;
    } else {
        if (!prior_in) return fail(h, LDPC_EINVAL, "variable-node step needs priors");
        LAUNCH(h, (transpose_tile<T>), dim3((t.n + 31) / 32, (B + 31) / 32), block, s, (const T *)prior_in, prior, B, t.n, t.n, L.Bp);
        const dim3 grid = sweep_grid(h, L.ngroups, t.n, &p.per_cta);
        rc = launch_vn<T>(h, p, grid, s);
        if (rc) return rc;
        if (marg_out)
            LAUNCH(h, (transpose_tile<T>), dim3((B + 31) / 32, (t.n + 31) / 32), block, s, p.marg, (T *)marg_out, t.n, B, L.Bp, t.n);
    }
    LAUNCH(h, (transpose_tile<T>), dim3((B + 31) / 32, (t.E + 31) / 32), block, s, p.msg, (T *)msg_out, t.E, B, L.Bp, t.E);
    return check_launch(h, "debug_step");
}

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int ldpc_abi_version(void) { return LDPC_ABI_VERSION; }

const char *ldpc_last_error(const ldpc_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

unsigned long long ldpc_launch_count(const ldpc_t *h) { return h ? h->launches : 0ull; }

int ldpc_resident_frames(const ldpc_t *h) { return (h && h->res.ok) ? 4 * h->res.Q : 0; }

const char *ldpc_resident_kernel(const ldpc_t *h)
{
    if (!h || !h->res.ok) return "";
    return (h->res.vp || h->res.vx) ? "resident_vp" : "resident_bp";
}

int ldpc_resident_plan(const ldpc_t *h, long *out)
{
    if (!h || !out) return LDPC_EINVAL;
    if (!h->res.ok) return LDPC_EUNSUPPORTED;
    std::copy(h->res.plan, h->res.plan + 7, out);
    return LDPC_OK;
}

int ldpc_create(ldpc_t **out, int device, int n, int m, int E,
                const int32_t *chk_ptr, const int32_t *edge_var,
                const int32_t *var_ptr, const int32_t *var_edges)
{
    if (!out) return fail(nullptr, LDPC_EINVAL, "out is NULL");
    *out = nullptr;
    if (n <= 0 || m <= 0 || E <= 0 || !chk_ptr || !edge_var || !var_ptr || !var_edges)
        return fail(nullptr, LDPC_EINVAL, "bad table arguments");
    // ---- validate: np.where(H) order = check-major, ascending variable; var_edges ascending per variable
    if (chk_ptr[0] != 0 || chk_ptr[m] != E || var_ptr[0] != 0 || var_ptr[n] != E)
        return fail(nullptr, LDPC_EINVAL, "pointer tables must start at 0 and end at E");
    int max_dc = 0, max_dv = 0, min_dc = E, min_dv = E;
    for (int c = 0; c < m; ++c) {
        const int d = chk_ptr[c + 1] - chk_ptr[c];
        if (d < 0) return fail(nullptr, LDPC_EINVAL, "chk_ptr not monotone");
        max_dc = std::max(max_dc, d); min_dc = std::min(min_dc, d);
        for (int e = chk_ptr[c]; e < chk_ptr[c + 1]; ++e) {
            if (edge_var[e] < 0 || edge_var[e] >= n) return fail(nullptr, LDPC_EINVAL, "edge_var out of range");
            if (e > chk_ptr[c] && edge_var[e] <= edge_var[e - 1]) return fail(nullptr, LDPC_EINVAL, "edge_var not ascending within a check");
        }
    }
    std::vector<char> seen((size_t)E, 0);
    for (int v = 0; v < n; ++v) {
        const int d = var_ptr[v + 1] - var_ptr[v];
        if (d < 0) return fail(nullptr, LDPC_EINVAL, "var_ptr not monotone");
        max_dv = std::max(max_dv, d); min_dv = std::min(min_dv, d);
        for (int k = var_ptr[v]; k < var_ptr[v + 1]; ++k) {
            const int e = var_edges[k];
            if (e < 0 || e >= E || edge_var[e] != v || seen[e]) return fail(nullptr, LDPC_EINVAL, "var_edges inconsistent with edge_var");
            if (k > var_ptr[v] && e <= var_edges[k - 1]) return fail(nullptr, LDPC_EINVAL, "var_edges not ascending within a variable");
            seen[e] = 1;
        }
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, LDPC_ECUDA, std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, LDPC_EINVAL, "bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, LDPC_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));

    ldpc_t *h = new (std::nothrow) ldpc_t();
    if (!h) return fail(nullptr, LDPC_ENOMEM, "handle");
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        h->sm_count = prop.multiProcessorCount;
        h->smem_optin = prop.sharedMemPerBlockOptin;
        h->smem_per_sm = prop.sharedMemPerMultiprocessor;
    }
    if (const char *pe = getenv("LDPC_PLAN_EFFORT")) h->plan_effort = std::max(0.0, atof(pe));
    Tables &t = h->t;
    t.n = n; t.m = m; t.E = E;
    t.max_dc = max_dc; t.max_dv = max_dv;
    t.uni_dc = (max_dc == min_dc) ? max_dc : 0;
    t.uni_dv = (max_dv == min_dv) ? max_dv : 0;
    auto up = [&](int **dst, const int32_t *src, size_t cnt) -> cudaError_t {
        cudaError_t r = cudaMalloc((void **)dst, cnt * sizeof(int));
        if (r != cudaSuccess) return r;
        return cudaMemcpy(*dst, src, cnt * sizeof(int), cudaMemcpyHostToDevice);
    };
    if ((e = up(&t.chk_ptr, chk_ptr, (size_t)m + 1)) != cudaSuccess || (e = up(&t.edge_var, edge_var, (size_t)E)) != cudaSuccess ||
        (e = up(&t.var_ptr, var_ptr, (size_t)n + 1)) != cudaSuccess || (e = up(&t.var_edges, var_edges, (size_t)E)) != cudaSuccess) {
        const std::string msg = std::string("table upload: ") + cudaGetErrorString(e);
        ldpc_destroy(h);
        return fail(nullptr, LDPC_ECUDA, msg);
    }
    if (build_resident(h, chk_ptr, edge_var, var_ptr, var_edges) != LDPC_OK) {
        const std::string msg = g_create_error;
        ldpc_destroy(h);
        return fail(nullptr, LDPC_ECUDA, msg);
    }
    *out = h;
    return LDPC_OK;
}

int ldpc_profile_enable(ldpc_t *h, int on)
{
    if (!h) return LDPC_EINVAL;
    h->prof = on != 0;
    return LDPC_OK;
}

int ldpc_profile_read(ldpc_t *h, double *cn_ms, unsigned long long *cn_launches, double *vn_ms,
                      unsigned long long *vn_launches)
{
    if (!h) return LDPC_EINVAL;
    double ms[2] = {0.0, 0.0};
    unsigned long long cnt[2] = {0ull, 0ull};
    for (size_t i = 0; i < h->prof_used; ++i) {
        ProfEvent &ev = h->prof_ev[i];
        CUDA_TRY(h, cudaEventSynchronize(ev.t1));
        float t = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&t, ev.t0, ev.t1));
        ms[ev.kind & 1] += (double)t;
        cnt[ev.kind & 1] += 1ull;
    }
    h->prof_used = 0;
    if (cn_ms) *cn_ms = ms[0];
    if (vn_ms) *vn_ms = ms[1];
    if (cn_launches) *cn_launches = cnt[0];
    if (vn_launches) *vn_launches = cnt[1];
    return LDPC_OK;
}

void ldpc_destroy(ldpc_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &ev : h->prof_ev) {
        if (ev.t0) cudaEventDestroy(ev.t0);
        if (ev.t1) cudaEventDestroy(ev.t1);
    }
    if (h->stage) {
        for (auto &sl : h->stage->slot) {
            if (sl.d_y) cudaFree(sl.d_y);
            if (sl.d_x) cudaFree(sl.d_x);
            if (sl.d_yp) cudaFree(sl.d_yp);
            if (sl.d_xp) cudaFree(sl.d_xp);
            if (sl.d_it) cudaFree(sl.d_it);
            if (sl.d_rs) cudaFree(sl.d_rs);
            if (sl.ws) cudaFree(sl.ws);
            if (sl.stream) cudaStreamDestroy(sl.stream);
        }
        if (h->stage->h_flag) cudaFreeHost(h->stage->h_flag);
        delete h->stage;
    }
    for (int tb = 0; tb < 2; ++tb) {
        if (h->res.cvar[tb]) cudaFree(h->res.cvar[tb]);
        if (h->res.vrow[tb]) cudaFree(h->res.vrow[tb]);
    }
    for (int tb = 0; tb < 2; ++tb) {
        if (h->res.vp_cw[tb]) cudaFree(h->res.vp_cw[tb]);
        if (h->res.vx_cwx[tb]) cudaFree(h->res.vx_cwx[tb]);
        if (h->res.vp_vposmap[tb]) cudaFree(h->res.vp_vposmap[tb]);
        if (h->res.vp_vinvmap[tb]) cudaFree(h->res.vp_vinvmap[tb]);
    }
    if (h->res.vx_vdeg) cudaFree(h->res.vx_vdeg);
    if (h->res.vposmap) cudaFree(h->res.vposmap);
    if (h->res.vinvmap) cudaFree(h->res.vinvmap);
    if (h->res.cdeg) cudaFree(h->res.cdeg);
    if (h->res.vdeg) cudaFree(h->res.vdeg);
    if (h->t.chk_ptr) cudaFree(h->t.chk_ptr);
    if (h->t.edge_var) cudaFree(h->t.edge_var);
    if (h->t.var_ptr) cudaFree(h->t.var_ptr);
    if (h->t.var_edges) cudaFree(h->t.var_edges);
    delete h;
}

size_t ldpc_workspace_bytes(const ldpc_t *h, int algo, int dtype, int B, unsigned flags)
{
    (void)flags;
    return workspace_bytes_impl(h, algo, dtype, B, true);
}

int ldpc_decode(ldpc_t *h, int algo, int dtype, const void *input, const uint8_t *y_hard, int B,
                int max_iter, int iter_cap, uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *marg_out,
                void *workspace, size_t workspace_bytes, unsigned flags, void *stream)
{
    InSpec in;
    in.channel = (algo == LDPC_BEC) ? LDPC_CH_BEC : LDPC_CH_PRIORS;
    in.in_dtype = dtype;
    in.src = input;
    in.y_hard = y_hard;
    in.param = 0.0;
    return decode_any(h, algo, dtype, in, B, max_iter, iter_cap, x_hat, iters, reason, marg_out,
                      workspace, workspace_bytes, flags, (cudaStream_t)stream);
}

int ldpc_decode_channel(ldpc_t *h, int channel, int algo, int dtype, double param,
                        const void *y, int y_dtype, int B, int max_iter, int iter_cap,
                        uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *marg_out,
                        void *workspace, size_t workspace_bytes, unsigned flags, void *stream)
{
    if (!h) return LDPC_EINVAL;
    if (algo == LDPC_BEC) channel = LDPC_CH_BEC;
    if (channel < LDPC_CH_PRIORS || channel > LDPC_CH_BEC) return fail(h, LDPC_EINVAL, "bad channel");
    if ((channel == LDPC_CH_PRIORS || channel == LDPC_CH_BIAWGN) && !in_dtype_ok(y_dtype))
        return fail(h, LDPC_EINVAL, "bad y_dtype");
    InSpec in;
    in.channel = channel;
    in.in_dtype = y_dtype;
    in.src = y;
    in.y_hard = nullptr;
    in.param = param;
    return decode_any(h, algo, dtype, in, B, max_iter, iter_cap, x_hat, iters, reason, marg_out,
                      workspace, workspace_bytes, flags, (cudaStream_t)stream);
}

int ldpc_llr_bsc(ldpc_t *h, int dtype, double llr, const uint8_t *y, void *priors, size_t count, void *stream)
{
    if (!h || !y || !priors) return fail(h, LDPC_EINVAL, "null buffer");
    if (count == 0) return LDPC_OK;
    ENTER(h);
    const int blocks = (int)std::min<size_t>((count + 255) / 256, (size_t)h->sm_count * 32);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LDPC_F32) LAUNCH(h, (llr_flat<uint8_t, float, IN_BSC>), blocks, 256, s, y, (float *)priors, count, llr);
    else if (dtype == LDPC_F64) LAUNCH(h, (llr_flat<uint8_t, double, IN_BSC>), blocks, 256, s, y, (double *)priors, count, llr);
    else return fail(h, LDPC_EINVAL, "bad dtype");
    return check_launch(h, "llr_bsc");
}

int ldpc_llr_biawgn(ldpc_t *h, int y_dtype, int dtype, double noise_var, const void *y, void *priors, size_t count, void *stream)
{
    if (!h || !y || !priors) return fail(h, LDPC_EINVAL, "null buffer");
    if (count == 0) return LDPC_OK;
    ENTER(h);
    const int blocks = (int)std::min<size_t>((count + 255) / 256, (size_t)h->sm_count * 32);
    cudaStream_t s = (cudaStream_t)stream;
    if (y_dtype == LDPC_F64 && dtype == LDPC_F64) LAUNCH(h, (llr_flat<double, double, IN_BIAWGN>), blocks, 256, s, (const double *)y, (double *)priors, count, noise_var);
    else if (y_dtype == LDPC_F64 && dtype == LDPC_F32) LAUNCH(h, (llr_flat<double, float, IN_BIAWGN>), blocks, 256, s, (const double *)y, (float *)priors, count, noise_var);
    else if (y_dtype == LDPC_F32 && dtype == LDPC_F64) LAUNCH(h, (llr_flat<float, double, IN_BIAWGN>), blocks, 256, s, (const float *)y, (double *)priors, count, noise_var);
    else if (y_dtype == LDPC_F32 && dtype == LDPC_F32) LAUNCH(h, (llr_flat<float, float, IN_BIAWGN>), blocks, 256, s, (const float *)y, (float *)priors, count, noise_var);
    else return fail(h, LDPC_EINVAL, "bad dtype");
    return check_launch(h, "llr_biawgn");
}

int ldpc_channel_generate(ldpc_t *h, int channel, double param, const uint8_t *x,
                          unsigned long long seed, unsigned long long frame0, int B, void *y, void *stream)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0 || !y) return fail(h, LDPC_EINVAL, "bad arguments");
    if (!(param >= 0.0)) return fail(h, LDPC_EINVAL, "channel parameter must be >= 0");
    ENTER(h);
    const Tables &t = h->t;
    const long long total = (long long)B * ((t.n + 3) / 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
    cudaStream_t s = (cudaStream_t)stream;
    switch (channel) {
    case LDPC_CH_BSC: LAUNCH(h, (channel_generate<GEN_BSC>), blocks, 256, s, y, x, B, t.n, param, seed, frame0); break;
    case LDPC_CH_BEC: LAUNCH(h, (channel_generate<GEN_BEC>), blocks, 256, s, y, x, B, t.n, param, seed, frame0); break;
    case LDPC_CH_BIAWGN: LAUNCH(h, (channel_generate<GEN_BIAWGN>), blocks, 256, s, y, x, B, t.n, sqrt(param), seed, frame0); break;
    default: return fail(h, LDPC_EINVAL, "bad channel");
    }
    return check_launch(h, "channel_generate");
}

int ldpc_count_errors(ldpc_t *h, const uint8_t *x_hat, const uint8_t *x, int B, int32_t *bit_errs, void *stream)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0 || !x_hat || !bit_errs) return fail(h, LDPC_EINVAL, "bad arguments");
    ENTER(h);
    LAUNCH(h, count_errors, (B + 7) / 8, 256, (cudaStream_t)stream, x_hat, x, B, h->t.n, bit_errs);
    return check_launch(h, "count_errors");
}

int ldpc_debug_step(ldpc_t *h, int algo, int dtype, int which, int B, const void *prior, const void *msg_in,
                    void *msg_out, void *marg, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0 || !msg_in || !msg_out || !workspace) return fail(h, LDPC_EINVAL, "bad arguments");
    if (algo != LDPC_MSA && algo != LDPC_SPA) return fail(h, LDPC_EINVAL, "debug step is MSA/SPA only");
    if (which != 0 && which != 1) return fail(h, LDPC_EINVAL, "which must be 0 (CN) or 1 (VN)");
    ENTER(h);
    if (dtype == LDPC_F32)
        return debug_step_t<float>(h, algo, which, B, prior, msg_in, msg_out, marg, workspace, workspace_bytes, (cudaStream_t)stream);
    if (dtype == LDPC_F64)
        return debug_step_t<double>(h, algo, which, B, prior, msg_in, msg_out, marg, workspace, workspace_bytes, (cudaStream_t)stream);
    return fail(h, LDPC_EINVAL, "bad dtype");
}

size_t ldpc_packed_row_bytes(int n) { return n > 0 ? align_up(((size_t)n + 7) / 8, 16) : 0; }

int ldpc_decode_host(ldpc_t *h, int channel, int algo, int dtype, double param,
                     const void *y, int y_dtype, int B, int max_iter, int iter_cap,
                     uint8_t *x_hat, int32_t *iters, uint8_t *reason, int chunk, unsigned flags)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0 || !y || !x_hat || !iters) return fail(h, LDPC_EINVAL, "bad arguments");
    if (algo == LDPC_BEC) channel = LDPC_CH_BEC;
    ENTER(h);
    int rc = ensure_stage(h);
    if (rc) return rc;
    const Tables &t = h->t;
    const bool in_packed = (flags & LDPC_IN_PACKED) != 0, out_packed = (flags & LDPC_OUT_PACKED) != 0;
    const int planes = (channel == LDPC_CH_BEC) ? 2 : 1;                  // erasure symbols: value plane + erasure plane
    const size_t pstride = ldpc_packed_row_bytes(t.n);
    size_t in_es;
    switch (channel) {
    case LDPC_CH_PRIORS: in_es = in_elem_size(y_dtype); if (!in_dtype_ok(y_dtype)) return fail(h, LDPC_EINVAL, "bad y_dtype"); break;
    case LDPC_CH_BIAWGN: in_es = in_elem_size(y_dtype); if (!in_dtype_ok(y_dtype)) return fail(h, LDPC_EINVAL, "bad y_dtype"); break;
    case LDPC_CH_BSC: case LDPC_CH_BEC: in_es = 1; break;
    default: return fail(h, LDPC_EINVAL, "bad channel");
    }
    if (in_packed && channel != LDPC_CH_BSC && channel != LDPC_CH_BEC)
        return fail(h, LDPC_EINVAL, "LDPC_IN_PACKED is for BSC / BEC symbol input");
    const unsigned dflags = flags & ~(LDPC_IN_PACKED | LDPC_OUT_PACKED | LDPC_HOST_ASYNC);
    const size_t in_row = in_packed ? planes * pstride : (size_t)t.n * in_es;      // bytes of one frame on the host side
    const size_t out_row = out_packed ? planes * pstride : (size_t)t.n;
    if (chunk <= 0) {
        // Enough chunks to overlap H2D / decode / D2H on the 3 slot streams and keep the pipeline's ramp and tail short,
        // each chunk big enough that its copies (>= ~4 MiB in) outweigh the ~40 us of launches and copy set-up it costs:
        // float32 rows of the n = 1200 code -> B / 16 (scripts/e2e_probe.py: 2048-frame chunks reach 0.90 of a plain
        // pinned H2D copy, 5461-frame chunks 0.87, one chunk 0.63); bit-packed symbols are 30 x smaller per frame and want
        // few, large chunks (round 2: BSC packed 26 -> 33 M frames/s, BEC packed 99 -> 150 M).  The chunk's message
        // workspace stays bounded (~1 GiB per slot).
        const size_t per_frame = (size_t)t.E * elem_size(dtype) + (size_t)t.n * elem_size(dtype) * 2;
        size_t cap = ((size_t)1 << 30) / std::max<size_t>(per_frame, 1);
        cap = std::max<size_t>(256, std::min<size_t>(cap, 65536));
        const size_t total_in = (size_t)B * in_row;
        size_t chunks = std::min<size_t>(16, std::max<size_t>(3, (total_in + ((size_t)4 << 20) - 1) / ((size_t)4 << 20)));
        size_t c = std::max<size_t>(((size_t)B + chunks - 1) / chunks, 2048);
        c = std::min(c, cap);
        chunk = (int)std::max<size_t>(128, (c + 127) / 128 * 128);
    }
    chunk = std::min(chunk, B);

    HostStage *st = h->stage;
    int idx = st->next;
    for (int b0 = 0; b0 < B; b0 += chunk, ++idx) {
        const int nb = std::min(chunk, B - b0);
        HostSlot &sl = st->slot[idx % HostStage::kSlots];
        const size_t wsb = workspace_bytes_impl(h, algo, dtype, nb, false);
        if (wsb == 0) return fail(h, LDPC_EINVAL, "bad algo/dtype");
        if ((rc = grow(h, &sl.d_y, &sl.y_bytes, (size_t)nb * t.n * in_es)) != 0) return rc;
        if ((rc = grow(h, &sl.d_x, &sl.x_bytes, (size_t)nb * t.n)) != 0) return rc;
        if (in_packed && (rc = grow(h, &sl.d_yp, &sl.yp_bytes, (size_t)nb * in_row)) != 0) return rc;
        if (out_packed && (rc = grow(h, &sl.d_xp, &sl.xp_bytes, (size_t)nb * out_row)) != 0) return rc;
        if (sl.f_cap < (size_t)nb) {
            if (sl.d_it) cudaFree(sl.d_it);
            if (sl.d_rs) cudaFree(sl.d_rs);
            sl.d_it = nullptr; sl.d_rs = nullptr; sl.f_cap = 0;
            CUDA_TRY(h, cudaMalloc((void **)&sl.d_it, (size_t)nb * sizeof(int32_t)));
            CUDA_TRY(h, cudaMalloc((void **)&sl.d_rs, (size_t)nb));
            sl.f_cap = (size_t)nb;
        }
        if ((rc = grow(h, &sl.ws, &sl.ws_bytes, wsb)) != 0) return rc;

        const char *src = (const char *)y + (size_t)b0 * in_row;
        const int pblocks = (int)std::min<long long>(((long long)nb * ((t.n + 7) / 8) + 255) / 256, (long long)h->sm_count * 16);
        if (in_packed) {
            CUDA_TRY(h, cudaMemcpyAsync(sl.d_yp, src, (size_t)nb * in_row, cudaMemcpyHostToDevice, sl.stream));
            LAUNCH(h, unpack_rows, pblocks, 256, sl.stream, sl.d_yp, (uint8_t *)sl.d_y, nb, t.n, (int)pstride, planes);
        } else {
            CUDA_TRY(h, cudaMemcpyAsync(sl.d_y, src, (size_t)nb * in_row, cudaMemcpyHostToDevice, sl.stream));
        }
        InSpec in;
        in.channel = channel; in.in_dtype = y_dtype; in.src = sl.d_y; in.y_hard = nullptr; in.param = param;
        rc = decode_any(h, algo, dtype, in, nb, max_iter, iter_cap, sl.d_x, sl.d_it, sl.d_rs, nullptr,
                        sl.ws, sl.ws_bytes, dflags, sl.stream);
        if (rc) return rc;
        if (out_packed) {
            LAUNCH(h, pack_rows, std::min((nb + 7) / 8, h->sm_count * 16), 256, sl.stream, sl.d_x, sl.d_xp, nb, t.n, (int)pstride, planes);
            CUDA_TRY(h, cudaMemcpyAsync(x_hat + (size_t)b0 * out_row, sl.d_xp, (size_t)nb * out_row, cudaMemcpyDeviceToHost, sl.stream));
        } else {
            CUDA_TRY(h, cudaMemcpyAsync(x_hat + (size_t)b0 * out_row, sl.d_x, (size_t)nb * out_row, cudaMemcpyDeviceToHost, sl.stream));
        }
        CUDA_TRY(h, cudaMemcpyAsync(iters + b0, sl.d_it, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
        if (reason) CUDA_TRY(h, cudaMemcpyAsync(reason + b0, sl.d_rs, (size_t)nb, cudaMemcpyDeviceToHost, sl.stream));
    }
    st->next = idx % HostStage::kSlots;
    if (flags & LDPC_HOST_ASYNC) return LDPC_OK;
    for (auto &sl : st->slot) CUDA_TRY(h, cudaStreamSynchronize(sl.stream));
    return check_launch(h, "decode_host");
}

// ------------------------------------------------------------------------------------------------
// Monte-Carlo round on the device (the loop body of src/main.py:37-45 for B frames at once)
// ------------------------------------------------------------------------------------------------
static size_t mc_y_bytes(const ldpc_t *h, int channel, int B) { return align_up((size_t)B * h->t.n * (channel == LDPC_CH_BIAWGN ? 4 : 1), 256); }

size_t ldpc_mc_scratch_bytes(const ldpc_t *h, int channel, int algo, int dtype, int B)
{
    if (!h || B <= 0) return 0;
    if (channel == LDPC_CH_BEC) algo = LDPC_BEC;
    const size_t ws = workspace_bytes_impl(h, algo, dtype, B, false);
    if (ws == 0) return 0;
    return mc_y_bytes(h, channel, B) + align_up((size_t)B * h->t.n, 256) + align_up((size_t)B * 4, 256) + align_up(ws, 256) + 256;
}

int ldpc_count_accumulate(ldpc_t *h, const uint8_t *x_hat, const uint8_t *x, const int32_t *iters, int B,
                          int32_t *bit_errs, long long *counters, int nhist, void *stream)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0 || !x_hat || !iters || !counters || nhist < 0) return fail(h, LDPC_EINVAL, "bad arguments");
    ENTER(h);
    LAUNCH(h, count_accumulate, (B + 7) / 8, 256, (cudaStream_t)stream, x_hat, x, iters, B, h->t.n, bit_errs,
           reinterpret_cast<unsigned long long *>(counters), nhist);
    return check_launch(h, "count_accumulate");
}

int ldpc_mc_round(ldpc_t *h, int channel, int algo, int dtype, double ch_param, double dec_param,
                  const uint8_t *x, unsigned long long seed, unsigned long long frame0, int B,
                  int max_iter, int iter_cap, long long *counters, int nhist,
                  void *scratch, size_t scratch_bytes, unsigned flags, void *stream)
{
    if (!h) return LDPC_EINVAL;
    if (B <= 0 || !counters || !scratch || nhist < 0) return fail(h, LDPC_EINVAL, "bad arguments");
    if (channel == LDPC_CH_BEC) algo = LDPC_BEC;
    if (scratch_bytes < ldpc_mc_scratch_bytes(h, channel, algo, dtype, B) || (reinterpret_cast<uintptr_t>(scratch) & 255u) != 0)
        return fail(h, LDPC_EWORKSPACE, "scratch too small (ldpc_mc_scratch_bytes) or not 256-byte aligned");
    const Tables &t = h->t;
    Carver cv(scratch);
    void *y = cv.take<char>(mc_y_bytes(h, channel, B));
    uint8_t *x_hat = cv.take<uint8_t>((size_t)B * t.n);
    int32_t *iters = cv.take<int32_t>((size_t)B);
    const size_t ws_bytes = workspace_bytes_impl(h, algo, dtype, B, false);
    void *ws = cv.take<char>(ws_bytes);
    int rc = ldpc_channel_generate(h, channel, ch_param, x, seed, frame0, B, y, stream);
    if (rc) return rc;
    rc = ldpc_decode_channel(h, channel, algo, dtype, dec_param, y, LDPC_F32, B, max_iter, iter_cap, x_hat, iters, nullptr, nullptr,
                             ws, ws_bytes, flags, stream);
    if (rc) return rc;
    return ldpc_count_accumulate(h, x_hat, x, iters, B, nullptr, counters, nhist, stream);
}

int ldpc_host_sync(ldpc_t *h)
{
    if (!h) return LDPC_EINVAL;
    if (!h->stage) return LDPC_OK;
    ENTER(h);
    for (auto &sl : h->stage->slot) CUDA_TRY(h, cudaStreamSynchronize(sl.stream));
    return LDPC_OK;
}

}  // extern "C"
