// resident_vp.cuh — on-chip flooding BP for regular short codes, "variable-plane" message layout.
//
// Same decoder as resident_bp.cuh (frames stay in shared memory and registers for all their iterations, slots are
// refilled from a bulk-copy ring, src/bpa.py:17-63), but the check-to-variable messages are stored where the
// VARIABLE will read them instead of where the check produced them:
//
//   marg  [np]          float4  last marginal of every variable position (4 frames = the lanes of a float4)
//   plane [DV][np]      float4  plane s, position v = the message on the s-th edge of variable v
//                               (s = rank of the edge in ascending-check order, the summation order of bpa.py:35)
//   prior [np]          float4
//
// CN  thread = (check, 4 frames): gathers marg of its DC variables, v2c = marg - c2v_old (c2v_old lives in the
//     thread's registers), check rule, then SCATTERS the DC results to plane[s_k][v_k].  Gather and scatter of step k
//     touch the same 16-byte bank group (np is a multiple of 8), so one placement makes both conflict-free.
// VN  thread = (variable, 4 frames): three CONTIGUOUS float4 loads + the prior, marg = prior + ((c0 + c1) + c2).
//     No edge table, no index arithmetic, no bank conflicts: the variable phase only streams.
// Compared with resident_bp.cuh the random shared-memory accesses per iteration stay at 2E, but both sit in the check
// phase where the placement (res_layout.h, vn_contiguous) has every degree of freedom, and the variable phase loses
// half of its instructions.
//
// Two value-neutral simplifications the oracle tests cover (tests/test_host_emu.py):
//   * the reference's sum starts from 0.0 ((0 + c0) + c1 ...); 0 + c0 differs from c0 only for c0 = -0.0, and the
//     final marg = prior + s cannot tell -0.0 from +0.0 in s because priors are folded to +0.0 at refill;
//   * hard decisions are the sign bits of marg (never -0.0, NaN comes out of FADD with a clear sign bit), read back
//     from shared memory when a frame leaves instead of being tracked every iteration.
//
// IRR = true: the same kernel for IRREGULAR codes (check degrees 2..DC, variable degrees 0..8, holes allowed):
//   * res_layout.h orders the positions of a colour by descending degree, so the variables with more than k edges are
//     a prefix of the positions and plane k only has pcnt[k] cells (a multiple of 8) at byte offset pbase[k]: the
//     planes of `1200_rho_x5_*` take E + ~3 % cells instead of 8 n.  Cells of a plane whose variable has fewer edges
//     are never written and stay +0.0 (s + 0.0 == s up to the sign of zero, which marg = prior + s cannot see);
//   * one 32-bit index word per edge: (message cell << 16) | (variable position << 4), the check's degree in the low
//     bits of word 0; a check with fewer than DC edges is padded with edges that read a +inf marginal cell (neutral
//     for the minimum and for the parity) and write to a scratch cell, each in a bank group its step leaves free;
//   * sum-product (natural edge order) sees the padding as saturated inputs, u = 0, which is exactly how cn_spa_sc pads a
//     short check for every caller: messages are bit-identical to the streaming sweeps whatever the degree.
#pragma once
#include "resident_bp.cuh"

namespace ldpc {

struct VpSmem {
    size_t marg, planes, prior, stage, bars, hb, imap, total;
};
// imap_in_smem = false: the position -> variable map of the output stays in global memory (the one-CTA-per-SM
// geometry of codes up to n ≈ 2850, where the 2 n bytes are what makes room for the row ring).
__host__ __device__ inline VpSmem vp_smem_layout(int np, int dv, int ring, int stage_stride, bool imap_in_smem = true)
{
    VpSmem L;
    size_t o = 0;
    L.marg = o;   o += (size_t)np * 16;                       // offset 0: a packed variable offset IS the marg address
    L.planes = o; o += (size_t)dv * np * 16;
    L.prior = o;  o += (size_t)np * 16;
    L.stage = o;  o += (size_t)ring * stage_stride;
    L.bars = o;   o += (size_t)kResRingMax * 8;
    L.hb = o;     o += ((size_t)np + 15) / 16 * 16;
    L.imap = o;   o += imap_in_smem ? ((size_t)np * 2 + 15) / 16 * 16 : 0;
    L.total = o + 16;
    return L;
}

// Irregular codes: 8 padding cells (+inf) behind marg, 8 scratch cells behind the planes, one hard bit more per padding cell.
__host__ __device__ inline VpSmem vx_smem_layout(int np, int plane_cells, int ring, int stage_stride)
{
    VpSmem L;
    size_t o = 0;
    L.marg = o;   o += ((size_t)np + 8) * 16;                   // == vx_planes_offset(np), res_layout.h
    L.planes = o; o += ((size_t)plane_cells + 8) * 16;
    L.prior = o;  o += (size_t)np * 16;
    L.stage = o;  o += (size_t)ring * stage_stride;
    L.bars = o;   o += (size_t)kResRingMax * 8;
    L.hb = o;     o += ((size_t)np + 8 + 15) / 16 * 16;
    L.imap = o;   o += ((size_t)np * 2 + 15) / 16 * 16;
    L.total = o + 16;
    return L;
}

// Fields of a packed index word (two edges of a check):  [pos1 << 4 : 16][pos0 << 4 | sl1 << 2 | sl0 : 16], sl = slot + 1.
// volatile on purpose: the decoded offsets are loop-invariant, and hoisting them out of the iteration loop would need
// a dozen registers the kernel does not have.
__device__ __forceinline__ uint32_t vp_off0(uint32_t w) { uint32_t r; asm volatile("and.b32 %0, %1, 0xfff0;" : "=r"(r) : "r"(w)); return r; }
__device__ __forceinline__ uint32_t vp_off1(uint32_t w) { uint32_t r; asm volatile("shr.u32 %0, %1, 16;" : "=r"(r) : "r"(w)); return r; }
__device__ __forceinline__ uint32_t vp_sl0(uint32_t w) { uint32_t r; asm volatile("and.b32 %0, %1, 3;" : "=r"(r) : "r"(w)); return r; }
__device__ __forceinline__ uint32_t vp_sl1x4(uint32_t w) { uint32_t r; asm volatile("and.b32 %0, %1, 12;" : "=r"(r) : "r"(w)); return r; }
// Irregular codes, one word per edge: byte offset of the variable's marg cell / of the edge's message cell.
__device__ __forceinline__ uint32_t vx_goff(uint32_t w) { uint32_t r; asm volatile("and.b32 %0, %1, 0xfff0;" : "=r"(r) : "r"(w)); return r; }
__device__ __forceinline__ uint32_t vx_soff(uint32_t w)
{
    uint32_t r;
    asm volatile("{\n.reg .b32 t;\nshr.u32 t, %1, 12;\nand.b32 %0, t, 0xffff0;\n}" : "=r"(r) : "r"(w));
    return r;
}

// Four consecutive received values of a row -> channel LLRs (exactly the reference's float64 expression rounded to
// float32, see res_llr) and hard input bits.  `vec`: the row is 16-byte aligned (always true for a staged row).
__device__ __forceinline__ void vp_load4(const unsigned char *row, int i4, int in_mode, int in_es, bool vec,
                                         double param, double nscale, float (&val)[4], uint32_t &hard4)
{
    hard4 = 0u;
    if (in_mode == IN_BSC) {
        uint32_t w;
        if (vec) w = *reinterpret_cast<const uint32_t *>(row + (size_t)i4 * 4);
        else w = (uint32_t)row[i4 * 4] | ((uint32_t)row[i4 * 4 + 1] << 8) | ((uint32_t)row[i4 * 4 + 2] << 16) | ((uint32_t)row[i4 * 4 + 3] << 24);
        const float lf = (float)param;                                   // (float)(L * (+-1)) == +-(float)L
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool one = ((w >> (8 * j)) & 0xffu) != 0u;
            hard4 |= (one ? 1u : 0u) << j;
            val[j] = one ? -lf : lf;
        }
    } else {
        double y[4];
        if (in_es == 8) {
            if (vec) {
                const double2 a = *reinterpret_cast<const double2 *>(row + (size_t)i4 * 32);
                const double2 b = *reinterpret_cast<const double2 *>(row + (size_t)i4 * 32 + 16);
                y[0] = a.x; y[1] = a.y; y[2] = b.x; y[3] = b.y;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) y[j] = reinterpret_cast<const double *>(row)[i4 * 4 + j];
            }
        } else if (in_es == 2) {                                         // binary16 rows (LDPC_F16)
            __half hv[4];
            if (vec) {
                const uint2 a = *reinterpret_cast<const uint2 *>(row + (size_t)i4 * 8);
                *reinterpret_cast<uint2 *>(hv) = a;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) hv[j] = reinterpret_cast<const __half *>(row)[i4 * 4 + j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = (double)__half2float(hv[j]);
        } else {
            float f[4];
            if (vec) {
                const float4 a = *reinterpret_cast<const float4 *>(row + (size_t)i4 * 16);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = reinterpret_cast<const float *>(row)[i4 * 4 + j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = (double)f[j];
        }
        if (in_mode == IN_BIAWGN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double pr = y[j] * nscale;                         // == (-2 y) * (1 / noise_var): scaling by -2 is exact
                val[j] = (float)pr;
                if (!llr_biawgn_fast_ok(pr)) val[j] = res_llr_biawgn_exact(-2.0 * y[j], param);   // rare, out of line
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) val[j] = (float)y[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) val[j] = __fadd_rn(val[j], 0.0f);        // -0.0 -> +0.0, NaN -> canonical
}

// ALGO: ALGO_MSA / ALGO_SPA_PHI.  DC: degree of every check.  DV: degree of every variable (<= 3: two slot bits).
// TT: threads per CTA when known at compile time (0 = blockDim.x).  np and mp are multiples of 8 without holes,
// mp <= 2 * T, np <= 4 * T, np % 4 == 0.
// NPC: number of variable positions when known at compile time (0 = p.n); the shipped ensemble is n = 1200, and with
// np, mp and T constant every shared-memory address of the variable phase is base + immediate and the pass bounds
// fold away.
// MAXT: 320 = two CTAs per SM (4 + 4 frames); 640 = ONE CTA per SM for codes whose 4 frames need the whole shared
// memory (n up to ≈ 2850: the Margulis code, n = 2640), 96 registers per thread either way.  20 warps is what 96
// registers allow (warps are allocated in fours: 21 would get 80 registers and spill), so that geometry makes five
// variable passes and, for the checks beyond 2 x 640, a third check pass whose c2v_old is read back from the planes
// it was scattered to instead of being kept in registers.
constexpr int kVpBigThreads = 640;
constexpr int kVpBigVnPasses = 5;
template <int ALGO, int DC, int DV, int TT, int NPC, bool IRR = false, int MAXT = 320>
__global__ void __launch_bounds__(MAXT, MAXT > 320 ? 1 : 2) resident_vp(const ResParams p)
{
    static_assert(MAXT == 320 || (MAXT == kVpBigThreads && !IRR && TT == 0 && NPC == 0), "two geometries");
    constexpr bool BIG = MAXT > 320;
    constexpr int VNP = BIG ? kVpBigVnPasses : kResVnPasses;
    static_assert(DC >= 2 && DC <= 8 && DV >= 1 && (IRR ? (DV <= 8 && DC <= 6) : DV <= 3),
                  "regular: slot field is two bits, index words hold two edges; irregular: c2v of two checks in registers");
    constexpr int F = 4, CH = IRR ? DC : (DC + 1) / 2;
    constexpr uint32_t ALL = 0xFu;
    extern __shared__ __align__(128) unsigned char smem[];
    // check positions that go with NPC: n dv / dc for a regular code; the irregular instance is only dispatched for the
    // rate-1/2 ensemble (np = 1200, mp = 600, decode_bp_resident)
    constexpr int MPC = IRR ? NPC / 2 : NPC * DV / DC;
    const int np = NPC ? NPC : p.n, mp = NPC ? MPC : p.m;
    const uint32_t S = (uint32_t)np * 16u;                            // bytes per plane
    const VpSmem L = IRR ? vx_smem_layout(np, p.plane_cells, p.ring, p.stage_stride) : vp_smem_layout(np, DV, p.ring, p.stage_stride, !BIG);
    float4 *marg = reinterpret_cast<float4 *>(smem + L.marg);
    float4 *planes = reinterpret_cast<float4 *>(smem + L.planes);
    float4 *prior = reinterpret_cast<float4 *>(smem + L.prior);
    unsigned char *stage = smem + L.stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint8_t *hb = smem + L.hb;                                       // hard input bits of the slots being loaded, by position
    const uint16_t *imap = BIG ? p.vinvmap : reinterpret_cast<const uint16_t *>(smem + L.imap);    // variable at a position (output)

    __shared__ int s_frame[F], s_it[F], s_assign[F];
    __shared__ int r_frame[kResRingMax], r_uses[kResRingMax];
    __shared__ uint32_t s_unsat[2], s_maxed[2], s_unsat0, s_newmask, s_exhausted;

    const int tid = threadIdx.x, T = TT ? TT : (int)blockDim.x, lane = tid & 31;
    const bool async = p.ring > 0;
    const bool have_hard = (p.in_mode == IN_BSC) || (p.in_mode == IN_COPY && p.y_hard != nullptr);
    const int nref = NPC ? NPC : p.nref;                             // == np (regular codes)
    const size_t row_bytes = (size_t)nref * p.in_es;
    const bool src_vec = ((reinterpret_cast<uintptr_t>(p.src) | row_bytes) & 15u) == 0;
    const double nscale = -2.0 * p.inv_param;

    // ---- per-thread graph indices -> registers (once per CTA)
    uint32_t cw[kResCnPasses][CH];
#pragma unroll
    for (int ps = 0; ps < kResCnPasses; ++ps) {
        const int c = tid + ps * T;
#pragma unroll
        for (int h = 0; h < CH; ++h) cw[ps][h] = 0u;
        if (c < mp) {
            if (IRR) {
#pragma unroll
                for (int k = 0; k < DC; ++k) cw[ps][k] = p.cwx[(size_t)c * 8 + k];
            } else {
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    const uint32_t e = p.cw[(size_t)c * 8 + k];      // (position << 4) | (slot + 1)
                    if (k & 1) cw[ps][k >> 1] |= (e & 0xfff0u) << 16 | (e & 3u) << 2;
                    else cw[ps][k >> 1] |= e & 0xfff3u;
                }
            }
        }
    }
    auto goff = [&](int ps, int k) -> uint32_t {                     // byte offset of the marg cell read in step k
        if (IRR) return vx_goff(cw[ps][k]);
        const uint32_t w = cw[ps][k >> 1];
        return (k & 1) ? vp_off1(w) : vp_off0(w);
    };
    // BIG: the check of the third pass (position tid + 2 T), its gather / scatter byte offsets straight from the table
    const int ctail = tid + kResCnPasses * T;
    auto tail_offsets = [&](uint32_t (&g)[DC], uint32_t (&sc)[DC]) {
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            const uint32_t e = p.cw[(size_t)ctail * 8 + k];          // (position << 4) | (slot + 1)
            g[k] = e & 0xfff0u;
            sc[k] = (e & 3u) * S + g[k];
        }
    };
    float4 old[kResCnPasses][DC];                                    // c2v of the thread's own checks
#pragma unroll
    for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
        for (int k = 0; k < DC; ++k) old[ps][k] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (!BIG)
        for (int i = tid; i < np; i += T) reinterpret_cast<uint16_t *>(smem + L.imap)[i] = p.vinvmap[i];
    if (IRR) {
        // cells nobody writes must read as +0.0 (short planes, holes), the padding cells behind marg as +inf
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), i4 = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        for (int i = tid; i < (int)(L.stage / 16); i += T) reinterpret_cast<float4 *>(smem)[i] = z4;
        for (int i = tid; i < np + 8; i += T) hb[i] = 0;
        __syncthreads();
        if (tid < 8) marg[np + tid] = i4;
    }
    if (tid == 0) {
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

    // A leaving frame's word = sign bits of its lane of marg (bpa.py:62; NaN and 0 -> bit 0).  Every thread reads the
    // cells it wrote itself in the variable phase, so no barrier is needed before the next phase overwrites them.
    auto output_bits = [&](uint32_t mask, int why) {
        uint32_t mq = mask & ALL;
        while (mq != 0u) {
            const int j = __ffs(mq) - 1;
            mq &= mq - 1u;
            uint8_t *dst = p.x_hat + (size_t)s_frame[j] * nref;
            const uint32_t *mj = reinterpret_cast<const uint32_t *>(marg) + j;
#pragma unroll
            for (int ps = 0; ps < VNP; ++ps) {
                const int item = tid + ps * T;
                if (item < np) {
                    const uint32_t v = imap[item];
                    if (!IRR || v != 0xffffu) dst[v] = (uint8_t)(mj[(size_t)item * 4] >> 31);
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
            const uint32_t nm = __reduce_or_sync(kFull, s_newmask);          // CTA-uniform: keep the masks in uniform registers
            exhausted = __reduce_or_sync(kFull, s_exhausted) != 0u;
            if (nm == 0u) break;

            // ---- received rows -> prior / marg columns, one new slot at a time, four values per thread and step
            for (int s = 0; s < F; ++s) {
                if (!((nm >> s) & 1u)) continue;
                const int g = s_frame[s];
                const unsigned char *row;
                bool vec;
                if (async) {
                    const int e = s_assign[s];
                    mbar_wait(&bars[e], (uint32_t)(r_uses[e] & 1));
                    row = stage + (size_t)e * p.stage_stride;
                    vec = true;
                } else {
                    row = (const unsigned char *)p.src + (size_t)g * row_bytes;
                    vec = src_vec;
                }
                const uint8_t *hrow = (p.in_mode == IN_COPY && p.y_hard != nullptr) ? p.y_hard + (size_t)g * nref : nullptr;
                float *mcol = reinterpret_cast<float *>(marg) + s;
                float *pcol = reinterpret_cast<float *>(prior) + s;
                for (int i4 = tid; i4 < nref / 4; i4 += T) {
                    float val[4];
                    uint32_t hard4;
                    vp_load4(row, i4, p.in_mode, p.in_es, vec, p.param, nscale, val, hard4);
                    const uint2 pw = __ldg(reinterpret_cast<const uint2 *>(p.vposmap) + i4);
                    const uint32_t pos[4] = {pw.x & 0xffffu, pw.x >> 16, pw.y & 0xffffu, pw.y >> 16};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        mcol[(size_t)pos[j] * 4] = val[j];
                        pcol[(size_t)pos[j] * 4] = val[j];
                    }
                    if (have_hard) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t hbit = (hard4 >> j) & 1u;
                            if (hrow != nullptr) hbit = (uint32_t)(hrow[i4 * 4 + j] != 0);
                            hb[pos[j]] = (uint8_t)((hb[pos[j]] & ~(1u << s)) | (hbit << s));   // the same thread owns hb[pos] for every slot
                        }
                    }
                }
                if (IRR) {                                 // the last n % 4 values of the row
                    for (int v = (nref & ~3) + tid; v < nref; v += T) {
                        uint32_t hbit;
                        const float val = res_llr(row, v, p.in_mode, p.in_es, p.param, p.inv_param, &hbit);
                        const uint32_t pos = p.vposmap[v];
                        mcol[(size_t)pos * 4] = val;
                        pcol[(size_t)pos * 4] = val;
                        if (have_hard) {
                            if (hrow != nullptr) hbit = (uint32_t)(hrow[v] != 0);
                            hb[pos] = (uint8_t)((hb[pos] & ~(1u << s)) | (hbit << s));
                        }
                    }
                }
            }
            // ---- the new frames start from c2v = 0 (one new frame is the common case)
            if ((nm & (nm - 1u)) == 0u) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (nm == (1u << j)) {
#pragma unroll
                        for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
                            for (int k = 0; k < DC; ++k) (&old[ps][k].x)[j] = 0.f;
                    }
            } else {
#pragma unroll
                for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        if (nm & 1u) old[ps][k].x = 0.f;
                        if (nm & 2u) old[ps][k].y = 0.f;
                        if (nm & 4u) old[ps][k].z = 0.f;
                        if (nm & 8u) old[ps][k].w = 0.f;
                    }
            }
            if (BIG && ctail < mp) {                       // third-pass checks keep their c2v in the planes only
                uint32_t g[DC], sc[DC];
                tail_offsets(g, sc);
#pragma unroll
                for (int k = 0; k < DC; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((nm >> j) & 1u) reinterpret_cast<float *>(smem + sc[k])[j] = 0.f;
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
                    if (tid + ps * T < mp) {
                        uint32_t syn = 0u;
#pragma unroll
                        for (int k = 0; k < DC; ++k) syn ^= hb[goff(ps, k) >> 4];
                        u0 |= syn & ALL;
                    }
                }
                if (BIG && ctail < mp) {
                    uint32_t g[DC], sc[DC], syn = 0u;
                    tail_offsets(g, sc);
#pragma unroll
                    for (int k = 0; k < DC; ++k) syn ^= hb[g[k] >> 4];
                    u0 |= syn & ALL;
                }
                u0 = __reduce_or_sync(kFull, u0);
                if (lane == 0 && u0 != 0u) atomicOr(&s_unsat0, u0);
                __syncthreads();
                z = nm & ~__reduce_or_sync(kFull, s_unsat0);
                for (int s = 0; s < F; ++s) {
                    if (!((z >> s) & 1u)) continue;
                    const int g = s_frame[s];
                    uint8_t *dst = p.x_hat + (size_t)g * nref;
                    for (int i = tid; i < np; i += T) {
                        const uint32_t v = imap[i];
                        if (!IRR || v != 0xffffu) dst[v] = (uint8_t)((hb[i] >> s) & 1u);
                    }
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
            if (tid + ps * T < mp) {
                float4 mg[DC];
#pragma unroll
                for (int k = 0; k < DC; ++k) mg[k] = *reinterpret_cast<const float4 *>(smem + goff(ps, k));
                // v2c = marg - c2v_old (bpa.py:37); the sign bits of marg are the current hard decisions (bpa.py:62)
                uint32_t sx[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int k = 0; k < DC; ++k) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float mv = (&mg[k].x)[j];
                        sx[j] ^= f32_bits(mv);           // sign bit == (marg < 0): marg is never -0.0, and a NaN is the FADD's +NaN
                        (&mg[k].x)[j] = __fsub_rn(mv, (&old[ps][k].x)[j]);
                    }
                }
                const uint32_t syn = (sx[0] >> 31) | ((sx[1] >> 31) << 1) | ((sx[2] >> 31) << 2) | ((sx[3] >> 31) << 3);
                // Irregular codes, sum-product: a padding edge reads +inf, which the rule saturates to the neutral u = 0
                // (cn_spa_sc pads short checks the same way, so the streaming sweeps agree bit for bit); its own output
                // is forced to 0 so that the next v2c = inf - old stays +inf.
                const int dcr = (IRR && ALGO != ALGO_MSA) ? (int)(cw[ps][0] & 15u) : DC;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a[DC], o[DC];
#pragma unroll
                    for (int k = 0; k < DC; ++k) a[k] = (&mg[k].x)[j];
                    if (ALGO == ALGO_MSA) cn_msa_lean<DC>(a, o);             // irregular: padding edges read +inf, neutral
                    else cn_spa_sc<DC>(a, DC, o, p.sat_llr);
#pragma unroll
                    for (int k = 0; k < DC; ++k) (&old[ps][k].x)[j] = (IRR && ALGO != ALGO_MSA && k >= 2 && k >= dcr) ? 0.f : o[k];
                }
                // scatter: plane (slot) of the variable's edge, same bank group as the gather of the same step
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    uint32_t coff;
                    if (IRR) {
                        coff = vx_soff(cw[ps][k]);
                    } else {
                        const uint32_t w = cw[ps][k >> 1];
                        coff = (k & 1) ? vp_sl1x4(w) * (S >> 2) + vp_off1(w) : vp_sl0(w) * S + vp_off0(w);   // decoded again: registers
                    }
                    *reinterpret_cast<float4 *>(smem + coff) = old[ps][k];
                }
                unsat |= syn;
            }
        }
        if (BIG && ctail < mp) {                             // third pass: c2v_old comes back from the planes
            uint32_t g[DC], sc[DC];
            tail_offsets(g, sc);
            float4 mg[DC], ol[DC];
#pragma unroll
            for (int k = 0; k < DC; ++k) {
                mg[k] = *reinterpret_cast<const float4 *>(smem + g[k]);
                ol[k] = *reinterpret_cast<const float4 *>(smem + sc[k]);
            }
            uint32_t sx[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < DC; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float mv = (&mg[k].x)[j];
                    sx[j] ^= f32_bits(mv);
                    (&mg[k].x)[j] = __fsub_rn(mv, (&ol[k].x)[j]);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a[DC], o[DC];
#pragma unroll
                for (int k = 0; k < DC; ++k) a[k] = (&mg[k].x)[j];
                if (ALGO == ALGO_MSA) cn_msa_lean<DC>(a, o);
                else cn_spa_sc<DC>(a, DC, o, p.sat_llr);
#pragma unroll
                for (int k = 0; k < DC; ++k) (&ol[k].x)[j] = o[k];
            }
#pragma unroll
            for (int k = 0; k < DC; ++k) *reinterpret_cast<float4 *>(smem + sc[k]) = ol[k];
            unsat |= (sx[0] >> 31) | ((sx[1] >> 31) << 1) | ((sx[2] >> 31) << 2) | ((sx[3] >> 31) << 3);
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
        if (decoded != 0u) output_bits(decoded, LDPC_REASON_DECODED);     // marg still holds the last variable phase

        // ======================================= variable-node phase =======================================
#pragma unroll
        for (int ps = 0; ps < VNP; ++ps) {
            const int item = tid + ps * T;
            if (IRR) {
                if (item < np) {
                    // plane k = a prefix of the positions (descending degree): ascending edge order, bpa.py:35
                    float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (item < p.pcnt[0]) sm = *reinterpret_cast<const float4 *>(smem + p.pbase[0] + (size_t)item * 16);
#pragma unroll
                    for (int k = 1; k < DV; ++k) {
                        if (item >= p.pcnt[k]) break;
                        const float4 c = *reinterpret_cast<const float4 *>(smem + p.pbase[k] + (size_t)item * 16);
                        sm.x = __fadd_rn(sm.x, c.x); sm.y = __fadd_rn(sm.y, c.y);
                        sm.z = __fadd_rn(sm.z, c.z); sm.w = __fadd_rn(sm.w, c.w);
                    }
                    const float4 pr = prior[item];
                    marg[item] = make_float4(__fadd_rn(pr.x, sm.x), __fadd_rn(pr.y, sm.y), __fadd_rn(pr.z, sm.z), __fadd_rn(pr.w, sm.w));
                }
            } else if (item < np) {
                float4 c[DV];
#pragma unroll
                for (int k = 0; k < DV; ++k) c[k] = planes[(size_t)k * np + item];
                const float4 pr = prior[item];
                float4 mgv;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float s = (&c[0].x)[j];                                            // 0 + c0: see the header
#pragma unroll
                    for (int k = 1; k < DV; ++k) s = __fadd_rn(s, (&c[k].x)[j]);
                    (&mgv.x)[j] = __fadd_rn((&pr.x)[j], s);                            // bpa.py:35
                }
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
