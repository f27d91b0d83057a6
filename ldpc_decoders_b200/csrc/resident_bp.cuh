// resident_bp.cuh — on-chip flooding BP for short codes, with continuous frame refill: the check-major
// message layout.  Since resident_vp.cuh (variable-plane layout) took over every shipped code, this kernel serves the
// degree profiles that one does not cover (check degree 7..8, degree-1 checks) and LDPC_RESIDENT_LAYOUT=check; it also
// holds what both kernels share: ResParams, the LLR front end of the refill, the packed-index helpers.
//
// For n = 1200 one frame's decoder state is 19 KB, so a CTA keeps F = 4 Q frames ("slots", Q = 1 as shipped) in shared
// memory and registers for ALL their iterations: HBM sees the received row once on the way in and the hard
// decisions once on the way out (~6 KB per frame instead of 63 KB per frame-iteration).
//
// State of one slot (frames are the innermost dimension: a float4 = the 4 frames of one "quad", Q quads per CTA):
//   marg [n][Q]        float4  shared   last marginal of every variable (bpa.py:35); before the first
//                                       iteration it holds the prior, so v2c = marg - 0 = priors[yy] (bpa.py:19)
//   prior[n][Q]        float4  shared
//   c2v  [dc][m][Q]    float4  shared   check-to-variable messages, plane k = k-th edge of every check
//   c2v of the thread's OWN checks      additionally stays in registers across iterations (REGC variant)
// One iteration (src/bpa.py:27-63), two phases separated by __syncthreads:
//   CN  thread = (check, quad): gathers marg of its dc variables (their sign bits are the hard decisions,
//       so the syndrome of bpa.py:29 costs three LOP3), forms v2c = marg - c2v_old ("total minus own",
//       bpa.py:37, the same subtraction the reference does), runs the check-node rule, stores c2v.
//   VN  thread = (variable, quad): marg = prior + ((0 + c0) + c1 + ...) in ascending edge order (bpa.py:35).
// Shared-memory traffic per frame-iteration is 3E + 2n floats instead of the 4E + n of a message-array
// design, the per-thread graph indices live in registers (loaded once per CTA), and no hard-decision array
// is kept at all.
//
// Slots are independent: a frame that leaves (syndrome zero after the CN phase, or iteration bound after
// the VN phase) is written out at once and its slot is refilled from a ring of received rows that the
// bulk-copy engine (cp.async.bulk + mbarrier, TMA 1-D) keeps landing in shared memory while the CTA
// iterates — converged frames never occupy a lane, and no thread ever waits for HBM in steady state.
// Frames are handed out by a global atomic counter; results do not depend on which slot decodes a frame.
//
// Arithmetic and exit rules are those of the streaming sweeps (same ldpc_math.cuh functions; the (3,6)
// min-sum uses cn_msa_lean, value-identical for the finite inputs it is given), so words, iteration
// counts and exit reasons are bit-identical to them and to the reference.
#pragma once
#include "common.cuh"
#include "io_kernels.cuh"
#include "stream_bp_tma.cuh"

namespace ldpc {

// Two geometries: Q = 2 quads (8 slots) in one CTA per SM of up to 608 threads, or Q = 1 quad (4 slots) in CTAs of up
// to 320 threads, two of which share an SM and hide each other's barriers and refills.
constexpr int kResMaxQ = 2;
constexpr int kResMaxF = 4 * kResMaxQ;       // slots per CTA
__host__ __device__ constexpr int res_max_threads(int Q) { return Q == 2 ? 608 : 320; }    // 96 registers per thread either way
constexpr int kResCnPasses = 2;              // check items per thread  (m * Q <= 2 * threads)
constexpr int kResVnPasses = 4;              // variable items per thread (n * Q <= 4 * threads)
constexpr int kResRingMax = 8;               // staged received rows

struct ResParams {
    int n, m;                          // POSITIONS of variables / checks (res_layout.h; >= the code's n, m, multiples of 8)
    int nref;                          // the code's n: row length of src / x_hat
    int planes;                        // c2v planes = max check degree
    const uint16_t *cvar;              // [m][8]  position of the variable read in step k of the check at a position (padding: n)
    const uint16_t *vrow;              // [n][8]  c2v row (plane * m + check position) of the variable's edges, ascending edge order
    const uint8_t *cdeg, *vdeg;        // [m], [n] degrees by position (0 = hole)
    const uint16_t *vposmap;           // [nref] position of variable v
    const uint16_t *vinvmap;           // [n]    variable at a position (0xffff = hole)
    const uint16_t *cw;                // resident_vp.cuh: [m][8] (variable position << 4) | (edge rank at the variable + 1)
    const uint32_t *cwx;               // resident_vp.cuh, irregular codes: [m][8] (message cell << 16) | (variable position << 4) [| degree, k = 0]
    int pcnt[8], pbase[8];             // resident_vp.cuh, irregular codes: cells of plane k (a prefix of the positions), its byte offset
    int plane_cells;                   //   and the cells of all planes together
    int cn_items, vn_items;            // m * Q, n * Q
    const void *src;                   // [B][n] received block (or priors)
    int in_mode;                       // IN_COPY / IN_BSC / IN_BIAWGN
    int in_es;                         // element size of src: 1 (BSC), 2 (binary16), 4, 8
    const uint8_t *y_hard;             // optional hard input for IN_COPY
    double param, inv_param;           // inv_param = 1 / param (BIAWGN fast path)
    int B, limit, bound_reason;
    float sat_llr;                     // SPA: reference saturation point (ldpc_math.cuh)
    uint8_t *x_hat;
    int *iters;
    uint8_t *reason;
    int *counter;                      // frame dispenser (zeroed before launch)
    int ring;                          // staged rows (0: rows are read straight from global memory at refill)
    int stage_stride;                  // bytes between staged rows
};

// Shared-memory carve-up, shared by the kernel and the host-side size computation.
struct ResSmem {
    size_t c2v, marg, prior, vtab, stage, hb, imap, bars, total;
};
// vtab_words: 32-bit words per variable item of the shared-memory edge table (0: the table lives in registers).
__host__ __device__ inline ResSmem resident_smem_layout(int Q, int n, int m, int planes, int vtab_words, int ring, int stage_stride)
{
    ResSmem L;
    size_t o = 0;
    L.c2v = o;   o += ((size_t)planes * m * Q + 1) * 16;          // + one all-zero cell (padding edges of a variable)
    L.marg = o;  o += ((size_t)n * Q + 1) * 16;                   // + one +inf cell (padding edges of a check)
    L.prior = o; o += (size_t)n * Q * 16;
    L.vtab = o;  o += ((size_t)n * Q * vtab_words * 4 + 15) / 16 * 16;
    L.stage = o; o += (size_t)ring * stage_stride;
    L.bars = o;  o += (size_t)kResRingMax * 8;
    L.hb = o;    o += ((size_t)n + 15) / 16 * 16;
    L.imap = o;  o += ((size_t)n * 2 + 15) / 16 * 16;                 // variable at a position (output)
    L.total = o + 16;
    return L;
}

// Channel LLR of one received value, rounded exactly like the reference's float64 expression cast to float32
// (io_kernels.cuh llr_map; BIAWGN without a float64 division in the common case, ldpc_math.cuh llr_biawgn_f32).
__device__ __noinline__ float res_llr_biawgn_exact(double t, double param) { return (float)(t / param); }

// Element v of a received row as float64: in_es = 8 (float64), 4 (float32) or 2 (binary16, LDPC_F16); all exact.
__device__ __forceinline__ double res_in_f64(const void *row, int v, int in_es)
{
    if (in_es == 8) return ((const double *)row)[v];
    if (in_es == 4) return (double)((const float *)row)[v];
    return (double)__half2float(((const __half *)row)[v]);
}

__device__ __forceinline__ float res_llr(const void *row, int v, int in_mode, int in_es, double param, double inv_param, uint32_t *hard)
{
    *hard = 0u;
    float val;
    if (in_mode == IN_BSC) {
        const uint8_t y = ((const uint8_t *)row)[v];
        *hard = (uint32_t)(y != 0);
        const float lf = (float)param;                                   // (float)(L * (+-1)) == +-(float)L
        val = y ? -lf : lf;
    } else if (in_mode == IN_BIAWGN) {
        const double y = res_in_f64(row, v, in_es);
        const double t = -2.0 * y;                                       // exact
        const double pr = t * inv_param;
        val = (float)pr;
        if (!llr_biawgn_fast_ok(pr)) val = res_llr_biawgn_exact(t, param);   // rare: keep the division out of line
    } else {
        val = (float)res_in_f64(row, v, in_es);
    }
    return __fadd_rn(val, 0.0f);          // -0.0 -> +0.0 (value-neutral, see cn_msa_lean), NaN -> canonical
}

// Half of a packed index word.  Opaque to the optimiser on purpose: otherwise the unpacked offsets are hoisted out of
// the iteration loop as 24 loop-invariant registers, which do not exist (the kernel is capped at 96) and get spilled.
__device__ __forceinline__ uint32_t lds_u16x2(const uint32_t w, int hi)
{
    uint32_t r;
    if (hi) asm volatile("shr.u32 %0, %1, 16;" : "=r"(r) : "r"(w));
    else asm volatile("and.b32 %0, %1, 0xffff;" : "=r"(r) : "r"(w));
    return r;
}

template <int Q, int ALGO, int DCP, bool UDC, int DVP, bool UDV, bool REGC, bool VSM>
__global__ void __launch_bounds__(res_max_threads(Q), 3 - Q) resident_bp(const ResParams p)
{
    constexpr int F = 4 * Q, QSH = (Q == 2) ? 1 : 0;
    constexpr int CH = (DCP + 1) / 2, VH = (DVP + 1) / 2;
    constexpr int VR = VSM ? 1 : kResVnPasses;
    constexpr uint32_t ALL = (1u << F) - 1u;
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = p.n, m = p.m;
    const ResSmem L = resident_smem_layout(Q, n, m, p.planes, VSM ? VH : 0, p.ring, p.stage_stride);
    uint32_t *vtab = reinterpret_cast<uint32_t *>(smem + L.vtab);      // VSM: [n * Q][VH] packed c2v indices of a variable item
    float4 *c2v = reinterpret_cast<float4 *>(smem + L.c2v);
    float4 *marg = reinterpret_cast<float4 *>(smem + L.marg);
    float4 *prior = reinterpret_cast<float4 *>(smem + L.prior);
    unsigned char *stage = smem + L.stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint16_t *imap = reinterpret_cast<uint16_t *>(smem + L.imap);
    uint8_t *hb = smem + L.hb;                                       // hard input bits of the slots being loaded
    const uint32_t zero_cell = (uint32_t)p.planes * m * Q;           // float4 index into c2v
    const uint32_t inf_off = (uint32_t)n * Q * 16;                   // byte offset into marg

    __shared__ int s_frame[kResMaxF], s_it[kResMaxF], s_assign[kResMaxF];
    __shared__ int r_frame[kResRingMax], r_uses[kResRingMax];
    __shared__ uint32_t s_unsat[2], s_maxed[2], s_unsat0, s_newmask, s_exhausted;

    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
    const int q = tid & (Q - 1);                                     // T is even: every item of a thread is in quad q
    const bool async = p.ring > 0;
    const bool have_hard = (p.in_mode == IN_BSC) || (p.in_mode == IN_COPY && p.y_hard != nullptr);
    const int nref = p.nref;
    const size_t row_bytes = (size_t)nref * p.in_es;

    // ---- per-thread graph indices -> registers (once per CTA)
    uint32_t cidx[kResCnPasses][CH];          // byte offsets into marg of the check's variables, two per word
    int cdeg[kResCnPasses];
    uint32_t vidx[VR][VH];                    // float4 indices into c2v of the variable's edges, two per word (VSM: staged in vtab)
#pragma unroll
    for (int ps = 0; ps < kResCnPasses; ++ps) {
        const int item = tid + ps * T;
        cdeg[ps] = 0;
#pragma unroll
        for (int h = 0; h < CH; ++h) cidx[ps][h] = inf_off | (inf_off << 16);
        if (item < p.cn_items) {
            const int c = item >> QSH;
            cdeg[ps] = UDC ? DCP : (int)p.cdeg[c];
#pragma unroll
            for (int k = 0; k < DCP; ++k) {
                const uint32_t var = p.cvar[(size_t)c * 8 + k];
                const uint32_t off = (k < cdeg[ps]) ? (var * Q + q) * 16u : inf_off;
                if (k & 1) cidx[ps][k >> 1] = (cidx[ps][k >> 1] & 0xffffu) | (off << 16);
                else cidx[ps][k >> 1] = (cidx[ps][k >> 1] & 0xffff0000u) | off;
            }
        }
    }
#pragma unroll
    for (int ps = 0; ps < kResVnPasses; ++ps) {
        const int item = tid + ps * T;
        uint32_t w[VH];
#pragma unroll
        for (int h = 0; h < VH; ++h) w[h] = zero_cell | (zero_cell << 16);
        if (item < p.vn_items) {
            const int v = item >> QSH;
            const int dv = UDV ? DVP : (int)p.vdeg[v];
#pragma unroll
            for (int k = 0; k < DVP; ++k) {
                const uint32_t row = p.vrow[(size_t)v * 8 + k];
                const uint32_t idx = (k < dv) ? row * Q + q : zero_cell;
                if (k & 1) w[k >> 1] = (w[k >> 1] & 0xffffu) | (idx << 16);
                else w[k >> 1] = (w[k >> 1] & 0xffff0000u) | idx;
            }
        }
#pragma unroll
        for (int h = 0; h < VH; ++h) {
            if (VSM) { if (item < p.vn_items) vtab[(size_t)item * VH + h] = w[h]; }
            else vidx[ps][h] = w[h];
        }
    }
    float4 old[REGC ? kResCnPasses : 1][REGC ? DCP : 1];            // c2v of the thread's own checks
    if (REGC) {
#pragma unroll
        for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
            for (int k = 0; k < DCP; ++k) old[ps][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int i = tid; i < n; i += T) imap[i] = p.vinvmap[i];
    // ---- constants cells, slot state, ring
    if (tid == 0) {
        c2v[zero_cell] = make_float4(0.f, 0.f, 0.f, 0.f);
        marg[n * Q] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        s_unsat[0] = s_unsat[1] = s_maxed[0] = s_maxed[1] = 0u;
        s_unsat0 = 0u;
        s_exhausted = 0u;
        for (int e = 0; e < kResRingMax; ++e) { r_frame[e] = -1; r_uses[e] = 0; }
        if (async) {
            for (int e = 0; e < p.ring; ++e) mbar_init(&bars[e], 1u);
            fence_mbar_init();
        }
    }
    __syncthreads();
    int head = 0;                                                    // thread 0: next ring entry to hand out
    auto issue = [&](int e) {                                        // thread 0: fetch the next frame into ring entry e
        const int g = atomicAdd(p.counter, 1);
        if (g < p.B) {
            r_frame[e] = g;
            mbar_expect_tx(&bars[e], (uint32_t)row_bytes);
            bulk_g2s(stage + (size_t)e * p.stage_stride, (const char *)p.src + (size_t)g * row_bytes, (uint32_t)row_bytes, &bars[e]);
        } else {
            r_frame[e] = -1;
        }
    };
    if (tid == 0 && async)
        for (int e = 0; e < p.ring; ++e) issue(e);

    // Hard decisions (bpa.py:62; NaN and 0 -> bit 0) of the thread's own variable items after the last VN phase:
    // bit 4 * pass + j = frame j of the quad.  A leaving frame is written out from these registers.
    uint32_t hbits = 0u;
    auto output_bits = [&](uint32_t mask, int why) {
        uint32_t mq = (mask >> (4 * q)) & 0xFu;
        while (mq != 0u) {
            const int j = __ffs(mq) - 1;
            mq &= mq - 1u;
            uint8_t *dst = p.x_hat + (size_t)s_frame[4 * q + j] * nref;
#pragma unroll
            for (int ps = 0; ps < kResVnPasses; ++ps) {
                const int item = tid + ps * T;
                if (item < p.vn_items) {
                    const uint32_t v = imap[item >> QSH];
                    if (v != 0xffffu) dst[v] = (uint8_t)((hbits >> (4 * ps + j)) & 1u);
                }
            }
        }
        if (tid < F && ((mask >> tid) & 1u)) {
            const int g = s_frame[tid];
            p.iters[g] = s_it[tid];
            if (p.reason != nullptr) p.reason[g] = (uint8_t)why;
        }
    };

    uint32_t active = 0u, fresh = 0u, freem = ALL;                   // CTA-uniform slot masks
    bool exhausted = false;
    int par = 0;

    for (;;) {
        // ======================================= refill free slots =======================================
        while (freem != 0u && !exhausted) {
            __syncthreads();                               // outputs of leaving frames have read marg / hb
            if (tid == 0) {
                uint32_t nm = 0u;
                int used = 0;
                uint32_t exh = 0u;
                for (int s = 0; s < F; ++s) {
                    s_assign[s] = -1;
                    if (!((freem >> s) & 1u) || exh) continue;
                    int g, e = 0;
                    if (async) {
                        if (used == p.ring) continue;      // the rest is refilled at the next refill point
                        e = head % p.ring;
                        g = r_frame[e];
                    } else {
                        g = atomicAdd(p.counter, 1);
                        if (g >= p.B) g = -1;
                    }
                    if (g < 0) { exh = 1u; continue; }
                    s_assign[s] = e; s_frame[s] = g; s_it[s] = 0;
                    nm |= 1u << s;
                    ++head; ++used;
                }
                s_newmask = nm;
                s_exhausted = exh;
                s_unsat0 = 0u;
            }
            __syncthreads();
            const uint32_t nm = __reduce_or_sync(kFull, s_newmask);          // warp-uniform by construction: keep the
            exhausted = __reduce_or_sync(kFull, s_exhausted) != 0u;          // slot masks in uniform registers
            if (nm == 0u) break;

            // ---- received rows -> prior / marg columns, one new slot at a time (cost proportional to the frames loaded)
            for (int s = 0; s < F; ++s) {
                if (!((nm >> s) & 1u)) continue;
                const int g = s_frame[s];
                const void *row;
                if (async) {
                    const int e = s_assign[s];
                    mbar_wait(&bars[e], (uint32_t)(r_uses[e] & 1));
                    row = stage + (size_t)e * p.stage_stride;
                } else {
                    row = (const char *)p.src + (size_t)g * row_bytes;
                }
                const uint8_t *hrow = (p.in_mode == IN_COPY && p.y_hard != nullptr) ? p.y_hard + (size_t)g * nref : nullptr;
                float *mcol = reinterpret_cast<float *>(marg) + (s >> 2) * 4 + (s & 3);
                float *pcol = reinterpret_cast<float *>(prior) + (s >> 2) * 4 + (s & 3);
                for (int v = tid; v < nref; v += T) {
                    uint32_t hbit;
                    const float val = res_llr(row, v, p.in_mode, p.in_es, p.param, p.inv_param, &hbit);
                    const uint32_t pos = __ldg(p.vposmap + v);
                    mcol[(size_t)pos * (Q * 4)] = val;
                    pcol[(size_t)pos * (Q * 4)] = val;
                    if (have_hard) {
                        if (hrow != nullptr) hbit = (uint32_t)(hrow[v] != 0);
                        hb[pos] = (uint8_t)((hb[pos] & ~(1u << s)) | (hbit << s));   // the same thread owns hb[pos] for every slot
                    }
                }
            }
            // ---- the new frames start from c2v = 0
            const uint32_t nq = (nm >> (4 * q)) & 0xFu;
            if (nq != 0u) {
                if (REGC) {
#pragma unroll
                    for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
                        for (int k = 0; k < DCP; ++k) {
                            if (nq & 1u) old[ps][k].x = 0.f;
                            if (nq & 2u) old[ps][k].y = 0.f;
                            if (nq & 4u) old[ps][k].z = 0.f;
                            if (nq & 8u) old[ps][k].w = 0.f;
                        }
                } else {
#pragma unroll
                    for (int ps = 0; ps < kResCnPasses; ++ps) {
                        const int item = tid + ps * T;
                        if (item < p.cn_items)
                            for (int k = 0; k < cdeg[ps]; ++k) {
                                float *cell = reinterpret_cast<float *>(c2v + (size_t)k * m * Q + item);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if ((nq >> j) & 1u) cell[j] = 0.f;
                            }
                    }
                }
            }
            __syncthreads();                               // columns / hb visible; staged rows consumed
            if (tid == 0 && async) {
                for (int s = 0; s < F; ++s)
                    if ((nm >> s) & 1u) { const int e = s_assign[s]; r_uses[e] += 1; issue(e); }
            }
            uint32_t z = 0u;
            if (have_hard) {
                // ---- iteration-0 exit (bpa.py:29 on x_hat = y): syndrome of the hard input of the new frames
                uint32_t u0 = 0u;
#pragma unroll
                for (int ps = 0; ps < kResCnPasses; ++ps) {
                    const int item = tid + ps * T;
                    if (item < p.cn_items) {
                        uint32_t syn = 0u;
#pragma unroll
                        for (int k = 0; k < DCP; ++k)
                            if (UDC || k < cdeg[ps]) syn ^= hb[lds_u16x2(cidx[ps][k >> 1], k & 1) >> (4 + QSH)];
                        u0 |= ((syn >> (4 * q)) & 0xFu) << (4 * q);
                    }
                }
                u0 = __reduce_or_sync(kFull, u0);
                if (lane == 0 && u0 != 0u) atomicOr(&s_unsat0, u0);
                __syncthreads();
                z = nm & ~__reduce_or_sync(kFull, s_unsat0);
                for (int s = 0; s < F; ++s) {
                    if (!((z >> s) & 1u)) continue;
                    const int g = s_frame[s];
                    uint8_t *dst = p.x_hat + (size_t)g * nref;
                    for (int v = tid; v < nref; v += T) dst[v] = (uint8_t)((hb[__ldg(p.vposmap + v)] >> s) & 1u);
                    if (tid == 0) {
                        p.iters[g] = 0;
                        if (p.reason != nullptr) p.reason[g] = (uint8_t)LDPC_REASON_DECODED;
                    }
                }
            }
            const uint32_t started = nm & ~z;
            active |= started; fresh |= started; freem &= ~started;
            if (z == 0u) break;
        }
        if (active == 0u) break;

        // ======================================= check-node phase =======================================
        uint32_t unsat = 0u;
#pragma unroll
        for (int ps = 0; ps < kResCnPasses; ++ps) {
            const int item = tid + ps * T;
            if (item < p.cn_items) {
                const int dc = cdeg[ps];
                float4 mg[DCP];
#pragma unroll
                for (int k = 0; k < DCP; ++k)
                    mg[k] = *reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(marg) + lds_u16x2(cidx[ps][k >> 1], k & 1));
                float4 *own = c2v + item;
                // v2c = marg - c2v_old (bpa.py:37); the sign bits of marg are the current hard decisions (bpa.py:62)
                uint32_t sx[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int k = 0; k < DCP; ++k) {
                    float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (REGC) ov = old[ps][k];
                    else if (UDC || k < dc) ov = own[(size_t)k * m * Q];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float mv = (&mg[k].x)[j];
                        sx[j] ^= f32_bits(mv);           // sign bit == (marg < 0): marg is never -0.0, and a NaN is the FADD's +NaN
                        (&mg[k].x)[j] = (UDC || k < dc) ? __fsub_rn(mv, (&ov.x)[j]) : INFINITY;
                    }
                }
                const uint32_t syn = (sx[0] >> 31) | ((sx[1] >> 31) << 1) | ((sx[2] >> 31) << 2) | ((sx[3] >> 31) << 3);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a[DCP], o[DCP];
#pragma unroll
                    for (int k = 0; k < DCP; ++k) a[k] = (&mg[k].x)[j];
                    if (ALGO == ALGO_MSA) {
                        if (UDC) cn_msa_lean<DCP>(a, o);
                        else cn_msa_bits<DCP>(a, dc, o);
                    } else {
                        cn_spa_sc<DCP>(a, UDC ? DCP : dc, o, p.sat_llr);
                    }
#pragma unroll
                    for (int k = 0; k < DCP; ++k) (&mg[k].x)[j] = o[k];
                }
#pragma unroll
                for (int k = 0; k < DCP; ++k)
                    if (UDC || k < dc) {
                        own[(size_t)k * m * Q] = mg[k];
                        if (REGC) old[ps][k] = mg[k];
                    }
                unsat |= syn << (4 * q);
            }
        }
        unsat = __reduce_or_sync(kFull, unsat);
        if (lane == 0 && unsat != 0u) atomicOr(&s_unsat[par], unsat);
        __syncthreads();

        // ---- book-keeping: frames whose syndrome was zero leave here (bpa.py:29), iteration count unchanged
        const uint32_t us = __reduce_or_sync(kFull, s_unsat[par]) | fresh;            // a new frame's marg is its prior: no syndrome yet
        const uint32_t decoded = active & ~us;
        const uint32_t run = active & us;
        fresh = 0u;
        if (tid < F && ((run >> tid) & 1u)) {
            const int it = ++s_it[tid];                      // bpa.py:63
            if (it >= p.limit) atomicOr(&s_maxed[par], 1u << tid);       // bpa.py:28 at the top of the next round
        }
        if (tid == 0) { s_unsat[par ^ 1] = 0u; s_maxed[par ^ 1] = 0u; }
        if (decoded != 0u) output_bits(decoded, LDPC_REASON_DECODED);     // hbits still are those of the last VN phase

        // ======================================= variable-node phase =======================================
#pragma unroll
        for (int ps = 0; ps < kResVnPasses; ++ps) {
            const int item = tid + ps * T;
            if (item < p.vn_items) {
                uint32_t w[VH];
#pragma unroll
                for (int h = 0; h < VH; ++h) w[h] = VSM ? vtab[(size_t)item * VH + h] : vidx[VSM ? 0 : ps][h];
                float4 c[DVP];
#pragma unroll
                for (int k = 0; k < DVP; ++k) c[k] = c2v[VSM ? ((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu)) : lds_u16x2(w[k >> 1], k & 1)];
                const float4 pr = prior[item];
                float4 mgv;
                uint32_t hb4 = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float s = 0.0f;
#pragma unroll
                    for (int k = 0; k < DVP; ++k) s = __fadd_rn(s, (&c[k].x)[j]);      // padding edges add +0.0: exact
                    const float mj = __fadd_rn((&pr.x)[j], s);                         // bpa.py:35
                    (&mgv.x)[j] = mj;
                    hb4 |= (f32_bits(mj) >> 31) << j;                                  // == (mj < 0), see the CN phase
                }
                hbits = (hbits & ~(0xFu << (4 * ps))) | (hb4 << (4 * ps));
                marg[item] = mgv;
            }
        }
        __syncthreads();
        const uint32_t maxed = __reduce_or_sync(kFull, s_maxed[par]);
        if (maxed != 0u) output_bits(maxed, p.bound_reason);
        active = run & ~maxed;
        freem |= decoded | maxed;
        par ^= 1;
    }
}

}  // namespace ldpc
