// resident_bp.cuh — on-chip flooding BP for short codes.
//
// For n = 1200 one frame's whole decoder state is 19 KB, so a CTA keeps F frames (8 for float32) in shared
// memory for ALL iterations: HBM sees the received block once on the way in and the hard decisions once on
// the way out (~6 KB per frame instead of 63 KB per frame-iteration).  The arithmetic, its order and the
// exit rules are exactly those of the streaming sweeps (same ldpc_math.cuh functions), so results are
// bit-identical to them and to the reference.
//
// Shared-memory layout (Q = F/4 quads of 4 frames; one float4 = one quad of one row):
//   msg  [m*S][Q] float4   check-major rows, S = max_dc | 1 (odd): row of edge k of check c is c*S + k.
//                          A quarter-warp (8 lanes = 8/Q checks x Q quads) then touches 8 different
//                          16-byte bank groups in the check-node phase -> conflict-free LDS.128 / STS.128.
//   prior[n][Q]    float4
//   xb   [Q][n]    uint8   hard decisions of a quad (low 4 bits)
//   cvar [m][DCP]  uint16  variable index of every edge of a check (padded)
//   vpos [n][DVP]  uint16  msg row of every edge of a variable, ascending edge order (padded)
//   cdeg [m], vdeg [n] uint8
// Thread = (check or variable, quad); a CTA walks its items in passes, phases separated by __syncthreads:
//   load (LLR map fused) -> { CN + syndrome -> sync -> run = act & unsat -> VN -> sync } x iterations -> store.
// The F frames of a CTA run in lock-step; a frame that converged keeps its word and counter (masked), the CTA
// leaves the loop as soon as none of its frames runs.  CTAs fetch batches of F frames from an atomic counter.
#pragma once
#include "common.cuh"
#include "io_kernels.cuh"

namespace ldpc {

constexpr int kResMaxThreads = 640;

struct ResParams {
    int n, m, S;
    const uint16_t *cvar, *vpos;
    const uint8_t *cdeg, *vdeg;
    const void *src;                   // [B][n] received block (or priors)
    int in_mode;                       // IN_COPY / IN_BSC / IN_BIAWGN
    int in_f64;                        // element type of src for COPY / BIAWGN
    const uint8_t *y_hard;             // optional hard input for IN_COPY
    double param;
    int B, limit, bound_reason;
    float sat_llr;                     // SPA: reference saturation point (ldpc_math.cuh)
    uint8_t *x_hat;
    int *iters;
    uint8_t *reason;
    int *counter;                      // batch dispenser (zeroed before launch)
};

template <int DCP> struct IdxVec;
template <> struct IdxVec<8> { uint4 raw; __device__ __forceinline__ int get(int k) const { const uint32_t w = (&raw.x)[k >> 1]; return (k & 1) ? (int)(w >> 16) : (int)(w & 0xffffu); } };
template <> struct IdxVec<4> { uint2 raw; __device__ __forceinline__ int get(int k) const { const uint32_t w = (&raw.x)[k >> 1]; return (k & 1) ? (int)(w >> 16) : (int)(w & 0xffffu); } };

__device__ __forceinline__ float res_llr(const ResParams &p, size_t idx, uint8_t *hard)
{
    *hard = 0;
    if (p.in_mode == IN_BSC) {
        const uint8_t y = ((const uint8_t *)p.src)[idx];
        *hard = (uint8_t)(y != 0);
        return (float)(p.param * (double)(1 - 2 * (int)y));
    }
    const double y = p.in_f64 ? ((const double *)p.src)[idx] : (double)((const float *)p.src)[idx];
    if (p.in_mode == IN_BIAWGN) return (float)((-2.0 * y) / p.param);
    if (p.y_hard != nullptr) *hard = (uint8_t)(p.y_hard[idx] != 0);
    return (float)y;
}

template <int ALGO, int F, int DCP, int DVP>
__global__ void __launch_bounds__(kResMaxThreads) resident_bp(const ResParams p)
{
    constexpr int Q = F / 4;
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = p.n, m = p.m, S = p.S;
    float4 *msg = reinterpret_cast<float4 *>(smem);                                   // [m*S][Q]
    float4 *prior = msg + (size_t)m * S * Q;                                          // [n][Q]
    uint16_t *cvar = reinterpret_cast<uint16_t *>(prior + (size_t)n * Q);             // [m][DCP]
    uint16_t *vpos = cvar + (size_t)m * DCP;                                          // [n][DVP]
    uint8_t *xb = reinterpret_cast<uint8_t *>(vpos + (size_t)n * DVP);                // [Q][n]
    uint8_t *cdeg = xb + (size_t)Q * n;                                               // [m]
    uint8_t *vdeg = cdeg + m;                                                         // [n]
    __shared__ uint32_t s_unsat[2];
    __shared__ int s_batch;

    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;

    // ---- tables -> shared memory (once per CTA)
    for (int i = tid; i < m * DCP / 2; i += T) reinterpret_cast<uint32_t *>(cvar)[i] = reinterpret_cast<const uint32_t *>(p.cvar)[i];
    for (int i = tid; i < n * DVP / 2; i += T) reinterpret_cast<uint32_t *>(vpos)[i] = reinterpret_cast<const uint32_t *>(p.vpos)[i];
    for (int i = tid; i < m; i += T) cdeg[i] = p.cdeg[i];
    for (int i = tid; i < n; i += T) vdeg[i] = p.vdeg[i];

    const bool have_hard = (p.in_mode == IN_BSC) || (p.in_mode == IN_COPY && p.y_hard != nullptr);
    float *prior_f = reinterpret_cast<float *>(prior);                                // [n][F] scalar view

    for (;;) {
        __syncthreads();                                   // previous batch fully stored / tables visible
        if (tid == 0) s_batch = atomicAdd(p.counter, 1);
        if (tid < 2) s_unsat[tid] = 0u;
        __syncthreads();
        const int g0 = s_batch * F;
        if (g0 >= p.B) break;
        const uint32_t valid = (p.B - g0 >= F) ? ((1u << F) - 1u) : ((1u << (p.B - g0)) - 1u);

        // ---- load F frames: lane = frame + F * (variable mod 32/F); shared stores are conflict-free
        for (int i = tid; i < ((n * F + 31) & ~31); i += T) {      // whole warps stay in the loop (ballot below)
            const int f = i % F, v = i / F;
            const bool in = i < n * F;
            uint8_t hb = 0;
            float pr = 0.0f;
            if (in && ((valid >> f) & 1u)) pr = res_llr(p, (size_t)(g0 + f) * n + v, &hb);
            if (in) prior_f[i] = pr;
            // hard bits of the F frames of one variable sit in F adjacent lanes
            const uint32_t bal = __ballot_sync(kFull, hb != 0);
            if (in && f == 0) {
                const uint32_t byte = (bal >> (lane & ~(F - 1))) & ((1u << F) - 1u);
#pragma unroll
                for (int q = 0; q < Q; ++q) xb[q * n + v] = (uint8_t)((byte >> (4 * q)) & 0xFu);
            }
        }
        __syncthreads();

        uint32_t act = valid;
        int my_iters = 0;                                  // threads 0..F-1 count their frame's iterations
        int it = 0;
        for (; it < p.limit; ++it) {
            const bool first = (it == 0);
            const bool skip_syn = first && !have_hard;
            // ================= check-node phase (+ syndrome of the current hard decisions) =================
            uint32_t unsat = 0u;
            for (int item = tid; item < m * Q; item += T) {
                const int c = item / Q, q = item % Q;
                const int dc = cdeg[c];
                IdxVec<DCP> vars;
                vars.raw = *reinterpret_cast<const decltype(vars.raw) *>(cvar + (size_t)c * DCP);
                const int row0 = c * S;
                float4 v[DCP];
                uint32_t syn = 0u;
#pragma unroll
                for (int k = 0; k < DCP; ++k) {
                    if (k < dc) {
                        const int var = vars.get(k);
                        v[k] = first ? prior[var * Q + q] : msg[(row0 + k) * Q + q];
                        if (!skip_syn) syn ^= xb[q * n + var];
                    }
                }
                unsat |= (syn & 0xFu) << (4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a[DCP], o[DCP];
#pragma unroll
                    for (int k = 0; k < DCP; ++k) a[k] = (k < dc) ? (&v[k].x)[j] : 0.0f;
                    if (ALGO == ALGO_MSA) cn_msa_bits<DCP>(a, dc, o);
                    else cn_spa_phi<DCP>(a, dc, o, p.sat_llr);
#pragma unroll
                    for (int k = 0; k < DCP; ++k)
                        if (k < dc) (&v[k].x)[j] = o[k];
                }
#pragma unroll
                for (int k = 0; k < DCP; ++k)
                    if (k < dc) msg[(row0 + k) * Q + q] = v[k];
            }
            if (skip_syn) unsat = act;
            unsat = __reduce_or_sync(kFull, unsat);
            if (lane == 0 && unsat != 0u) atomicOr(&s_unsat[it & 1], unsat);
            __syncthreads();
            // ================= book-keeping (every thread computes the same masks) =================
            const uint32_t run = act & s_unsat[it & 1];      // frames whose syndrome was zero stop here (bpa.py:29)
            act = run;
            if (tid < F && ((run >> tid) & 1u)) my_iters += 1;                        // bpa.py:63
            if (tid == 0) s_unsat[(it + 1) & 1] = 0u;        // nobody touches the other buffer until the next CN phase
            if (run == 0u) break;
            // ================= variable-node phase =================
            for (int item = tid; item < n * Q; item += T) {
                const int vv = item / Q, q = item % Q;
                const int dv = vdeg[vv];
                IdxVec<DVP> pos;
                pos.raw = *reinterpret_cast<const decltype(pos.raw) *>(vpos + (size_t)vv * DVP);
                float4 cmsg[DVP];
#pragma unroll
                for (int k = 0; k < DVP; ++k)
                    if (k < dv) cmsg[k] = msg[pos.get(k) * Q + q];
                const float4 pr = prior[vv * Q + q];
                uint32_t bits = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a[DVP], o[DVP];
#pragma unroll
                    for (int k = 0; k < DVP; ++k) a[k] = (k < dv) ? (&cmsg[k].x)[j] : 0.0f;
                    const float marg = vn_update<float, DVP>((&pr.x)[j], a, dv, o);
#pragma unroll
                    for (int k = 0; k < DVP; ++k)
                        if (k < dv) (&cmsg[k].x)[j] = o[k];
                    bits |= (marg < 0.0f ? 1u : 0u) << j;
                }
#pragma unroll
                for (int k = 0; k < DVP; ++k)
                    if (k < dv) msg[pos.get(k) * Q + q] = cmsg[k];
                const uint32_t runq = (run >> (4 * q)) & 0xFu;
                uint8_t *xw = xb + q * n + vv;
                *xw = (uint8_t)((*xw & ~runq) | (bits & runq));                       // stopped frames keep their bits
            }
            __syncthreads();
        }

        // ---- store: words, iteration counts, exit reasons (frames still running hit the loop bound)
        for (int i = tid; i < n * F; i += T) {
            const int f = i / n, v = i % n;
            if ((valid >> f) & 1u) p.x_hat[(size_t)(g0 + f) * n + v] = (uint8_t)((xb[(f >> 2) * n + v] >> (f & 3)) & 1u);
        }
        if (tid < F && ((valid >> tid) & 1u)) {
            p.iters[g0 + tid] = my_iters;
            if (p.reason != nullptr)
                p.reason[g0 + tid] = ((act >> tid) & 1u) ? (uint8_t)p.bound_reason : (uint8_t)LDPC_REASON_DECODED;
        }
    }
}

inline size_t resident_smem_bytes(int n, int m, int S, int F, int DCP, int DVP)
{
    const int Q = F / 4;
    size_t b = 0;
    b += (size_t)m * S * Q * 16;        // msg
    b += (size_t)n * Q * 16;            // prior
    b += (size_t)m * DCP * 2;           // cvar
    b += (size_t)n * DVP * 2;           // vpos
    b += (size_t)Q * n;                 // xb
    b += (size_t)m + n;                 // degrees
    return align_up(b, 16) + 16;
}

}  // namespace ldpc
