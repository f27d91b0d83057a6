// stream_bp.cuh — streaming flooding BP (MSA / SPA) over edge-major, frame-contiguous messages.
//
// HBM layout (T = float or double, Bp = frames padded to a multiple of 32*FPT):
//   msg  [E][Bp]     ONE message array, updated in place: after the CN sweep it holds c2v, after the
//                    VN sweep v2c.  Edge e = position in np.where(H) (src/bpa.py:12), so the dc edges
//                    of a check are dc consecutive rows; a variable's rows are scattered but every row
//                    access is a contiguous run of frames -> fully coalesced 128-bit loads / stores.
//   prior[n][Bp]     channel LLRs (read by every VN sweep, and by the first CN sweep: v2c = priors[yy])
//   xbits[n][Bp/32]  hard decisions, bit l of word w = frame 32w + l
//   act / unsat [Bp/32]  per-frame flags (frame still running / some check unsatisfied)
//   iters[Bp]        the reference's iter_count
//
// One iteration = cn_sweep -> book -> vn_sweep  (src/bpa.py:27-63):
//   cn_sweep  c2v = CN(v2c) for every (check, frame); also XORs the current hard decisions of the
//             check's variables and ORs "unsatisfied" into unsat  (the syndrome test of bpa.py:29)
//   book      run = act & unsat: frames whose syndrome was zero stop here with iter_count unchanged,
//             the others get iter_count += 1 (bpa.py:63); act = run; unsat = 0
//   vn_sweep  marg = prior + sum c2v (ordered), v2c = marg - c2v, x_hat = marg < 0 for running frames
// Thread = (check or variable, FPT consecutive frames); warp = 32*FPT frames ("group"); CTA = 8 groups
// walking over a chunk of checks / variables.  A warp whose frames have all stopped returns at once.
#pragma once
#include "common.cuh"

namespace ldpc {

template <typename T> struct BpParams {
    int n, m, E;
    int Bp, wpr, ngroups;                    // wpr = Bp / 32 flag words per row
    const int *chk_ptr, *edge_var, *var_ptr, *var_edges;
    T *msg;
    const T *prior;
    T *marg;                                 // optional [n][Bp]: last marginal of running frames
    uint32_t *xbits, *act, *unsat;
    int *iters;
    int per_cta;                             // checks (CN) or variables (VN) per CTA
    int first;                               // CN: read v2c = prior[edge_var[e]] (src/bpa.py:19)
    int skip_syn;                            // CN: do not evaluate the syndrome (iteration 0 without hard input)
};


template <typename T, int ALGO, int DCMAX> struct CnMath;
template <typename T, int DCMAX> struct CnMath<T, ALGO_MSA, DCMAX> {
    static __device__ __forceinline__ void run(const T (&v)[DCMAX], int dc, T (&o)[DCMAX]) { cn_msa<T, DCMAX>(v, dc, o); }
};
template <int DCMAX> struct CnMath<double, ALGO_SPA_REF, DCMAX> {
    static __device__ __forceinline__ void run(const double (&v)[DCMAX], int dc, double (&o)[DCMAX]) { cn_spa_ref<DCMAX>(v, dc, o); }
};
template <int DCMAX> struct CnMath<float, ALGO_SPA_PHI, DCMAX> {
    static __device__ __forceinline__ void run(const float (&v)[DCMAX], int dc, float (&o)[DCMAX]) { cn_spa_sc<DCMAX>(v, dc, o); }
};

// ------------------------------------------------------------------------------------------------
// Check-node sweep.  UNIFORM: every check has exactly DCMAX edges (e0 = c * DCMAX, no pointer loads).
// ------------------------------------------------------------------------------------------------
template <typename T, int FPT, int ALGO, int DCMAX, bool UNIFORM>
__global__ void __launch_bounds__(kCtaThreads) cn_sweep(const BpParams<T> p)
{
    using P = Pack<T, FPT>;
    using F = Field<FPT>;
    const int lane = threadIdx.x & 31;
    const int group = blockIdx.x * kCtaWarps + (threadIdx.x >> 5);
    if (group >= p.ngroups) return;
    const int word = F::word_of(group, lane);
    const uint32_t act = F::get(p.act[word], lane);
    if (!__any_sync(kFull, act != 0u)) return;                      // all 32*FPT frames of this warp have stopped

    const size_t fbase = (size_t)group * (32 * FPT) + (size_t)lane * FPT;
    const int c_begin = blockIdx.y * p.per_cta;
    const int c_end = min(p.m, c_begin + p.per_cta);
    const bool need_var = p.first || !p.skip_syn;
    uint32_t unsat = 0u;

    for (int c = c_begin; c < c_end; ++c) {
        int e0, dc;
        if (UNIFORM) { e0 = c * DCMAX; dc = DCMAX; }
        else { e0 = __ldg(p.chk_ptr + c); dc = __ldg(p.chk_ptr + c + 1) - e0; }

        P v[DCMAX];
        uint32_t syn = 0u;
#pragma unroll
        for (int k = 0; k < DCMAX; ++k) {
            if (UNIFORM || k < dc) {
                const int var = need_var ? __ldg(p.edge_var + e0 + k) : 0;
                if (p.first) v[k] = ld_ro<P>(p.prior + (size_t)var * p.Bp + fbase);
                else v[k] = ld_stream<P>(p.msg + (size_t)(e0 + k) * p.Bp + fbase);
                if (!p.skip_syn) syn ^= __ldg(p.xbits + (size_t)var * p.wpr + word);
            }
        }
        unsat |= F::get(syn, lane);

#pragma unroll
        for (int j = 0; j < FPT; ++j) {
            T a[DCMAX], o[DCMAX];
#pragma unroll
            for (int k = 0; k < DCMAX; ++k) a[k] = (UNIFORM || k < dc) ? v[k].x[j] : (T)0;
            CnMath<T, ALGO, DCMAX>::run(a, dc, o);
#pragma unroll
            for (int k = 0; k < DCMAX; ++k)
                if (UNIFORM || k < dc) v[k].x[j] = o[k];
        }
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (UNIFORM || k < dc) st_stream<P>(p.msg + (size_t)(e0 + k) * p.Bp + fbase, v[k]);
    }

    if (!p.skip_syn) {
        const uint32_t w = F::assemble(unsat, lane);
        if (F::leader(lane) && w != 0u) atomicOr(p.unsat + word, w);
    }
}

// ------------------------------------------------------------------------------------------------
// Book-keeping between the two sweeps: one warp per flag word, lane = frame.
// ------------------------------------------------------------------------------------------------
// any_active[0] = some frame still runs; any_active[1] += number of running frames (the host's compaction decision).
__global__ void bp_book(uint32_t *act, uint32_t *unsat, int *iters, int wpr, int *any_active)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= wpr) return;
    const uint32_t run = act[w] & unsat[w];
    if ((run >> lane) & 1u) iters[(size_t)w * 32 + lane] += 1;      // bpa.py:63
    __syncwarp();
    if (lane == 0) {
        act[w] = run;
        unsat[w] = 0u;
        if (run != 0u) {
            any_active[0] = 1;
            atomicAdd(any_active + 1, __popc(run));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Active-frame compaction (north star (4)): when at most half of the frame columns still run, the live columns are
// packed to the front of every row, so that the sweeps touch [rows][live] instead of [rows][all] — converged frames stop
// consuming bandwidth at FRAME granularity, not only when a whole 512-frame tile is done.  Frames are independent, so
// the results are unchanged; `orig` remembers which frame a column holds.
//   compact_plan   src_of[j] = column of the j-th live frame (ascending: src_of[j] >= j)
//   compact_rows   row[j] = row[src_of[j]] for j < live, in place: one CTA owns a row and walks it in ascending chunks
//                  (load a chunk's sources, barrier, store) - sources are never behind the write front
//   compact_bits   the same for the bit-packed hard decisions
//   compact_flags  act = the first `live` columns
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) compact_plan(const uint32_t *__restrict__ act, int wx, int *__restrict__ src_of)
{
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int per = (wx + 1023) / 1024, w0 = tid * per, w1 = min(wx, w0 + per);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(act[w]);
    part[tid] = cnt;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                             // inclusive scan
        const int v = (tid >= d) ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int j = part[tid] - cnt;
    for (int w = w0; w < w1; ++w) {
        uint32_t a = act[w];
        while (a) {
            const int b = __ffs(a) - 1;
            a &= a - 1u;
            src_of[j++] = w * 32 + b;
        }
    }
}

constexpr int kCompactChunk = 1024;
template <typename U>
__global__ void __launch_bounds__(256) compact_rows(U *__restrict__ base, size_t pitch, const int *__restrict__ src_of, int live)
{
    U *row = base + (size_t)blockIdx.x * pitch;
    for (int c = 0; c < live; c += kCompactChunk) {
        U v[kCompactChunk / 256];
#pragma unroll
        for (int i = 0; i < kCompactChunk / 256; ++i) {
            const int j = c + threadIdx.x + i * 256;
            if (j < live) v[i] = row[__ldg(src_of + j)];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kCompactChunk / 256; ++i) {
            const int j = c + threadIdx.x + i * 256;
            if (j < live) row[j] = v[i];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) compact_bits(uint32_t *__restrict__ base, int wpr, const int *__restrict__ src_of, int live)
{
    uint32_t *row = base + (size_t)blockIdx.x * wpr;
    const int lw = (live + 31) / 32;
    for (int c = 0; c < lw; c += 128) {
        const int w = c + threadIdx.x;
        uint32_t out = 0u;
        if (w < lw) {
            for (int b = 0; b < 32; ++b) {
                const int j = w * 32 + b;
                if (j < live) {
                    const int sj = __ldg(src_of + j);
                    out |= ((row[sj >> 5] >> (sj & 31)) & 1u) << b;
                }
            }
        }
        __syncthreads();
        if (w < lw) row[w] = out;
        __syncthreads();
    }
}

// ... and the columns behind them (padding up to the tile size) hold no frame any more: orig = "none".
__global__ void compact_flags(uint32_t *act, uint32_t *unsat, int *orig, int wx, int live)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= wx) return;
    const int lo = w * 32;
    act[w] = (live >= lo + 32) ? 0xffffffffu : (live <= lo ? 0u : ((1u << (live - lo)) - 1u));
    unsat[w] = 0u;
    for (int j = max(lo, live); j < lo + 32; ++j) orig[j] = 0x7fffffff;
}

// orig[j] = j for the identity start.
__global__ void iota_kernel(int *a, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) a[i] = i;
}

// ------------------------------------------------------------------------------------------------
// Variable-node sweep.  UNIFORM: every variable has exactly DVMAX edges.
// ------------------------------------------------------------------------------------------------
template <typename T, int FPT, int DVMAX, bool UNIFORM>
__global__ void __launch_bounds__(kCtaThreads) vn_sweep(const BpParams<T> p)
{
    using P = Pack<T, FPT>;
    using F = Field<FPT>;
    const int lane = threadIdx.x & 31;
    const int group = blockIdx.x * kCtaWarps + (threadIdx.x >> 5);
    if (group >= p.ngroups) return;
    const int word = F::word_of(group, lane);
    const uint32_t run = F::get(p.act[word], lane);                 // act == frames running this iteration (after book)
    if (!__any_sync(kFull, run != 0u)) return;
    const uint32_t run_word = F::assemble(run, lane);

    const size_t fbase = (size_t)group * (32 * FPT) + (size_t)lane * FPT;
    const int v_begin = blockIdx.y * p.per_cta;
    const int v_end = min(p.n, v_begin + p.per_cta);

    for (int v = v_begin; v < v_end; ++v) {
        int p0, dv;
        if (UNIFORM) { p0 = v * DVMAX; dv = DVMAX; }
        else { p0 = __ldg(p.var_ptr + v); dv = __ldg(p.var_ptr + v + 1) - p0; }

        int eid[DVMAX];
        P c[DVMAX];
#pragma unroll
        for (int k = 0; k < DVMAX; ++k) {
            if (UNIFORM || k < dv) {
                eid[k] = __ldg(p.var_edges + p0 + k);
                c[k] = ld_stream<P>(p.msg + (size_t)eid[k] * p.Bp + fbase);
            }
        }
        const P pr = ld_ro<P>(p.prior + (size_t)v * p.Bp + fbase);

        uint32_t bits = 0u;
        P mg;
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
            T a[DVMAX], o[DVMAX];
#pragma unroll
            for (int k = 0; k < DVMAX; ++k) a[k] = (UNIFORM || k < dv) ? c[k].x[j] : (T)0;
            const T marg = vn_update<T, DVMAX>(pr.x[j], a, dv, o);
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (UNIFORM || k < dv) c[k].x[j] = o[k];
            bits |= (marg < (T)0 ? 1u : 0u) << j;                   // bpa.py:38,62: NaN -> 0 -> bit 0
            mg.x[j] = marg;
        }
#pragma unroll
        for (int k = 0; k < DVMAX; ++k)
            if (UNIFORM || k < dv) st_stream<P>(p.msg + (size_t)eid[k] * p.Bp + fbase, c[k]);

        const uint32_t w = F::assemble(bits, lane);
        if (F::leader(lane)) {
            uint32_t *xw = p.xbits + (size_t)v * p.wpr + word;
            *xw = (*xw & ~run_word) | (w & run_word);               // stopped frames keep their word
        }
        if (p.marg != nullptr) {
#pragma unroll
            for (int j = 0; j < FPT; ++j)
                if ((run >> j) & 1u) p.marg[(size_t)v * p.Bp + fbase + j] = mg.x[j];
        }
    }
}

}  // namespace ldpc
