// resident_vq.cuh — resident_vp with the frame hand-over FUSED INTO THE VARIABLE PHASE (regular codes, float32).
//
// In resident_vp a frame that leaves costs three CTA-wide barriers and a separate refill pass: every thread converts
// four received values and scatters them into one lane of marg / prior as 4-byte stores that are 4-way bank-conflicted
// by construction (19.9 M wavefronts for 5.4 M ideal), and the leaving frame's word is read back from shared memory.
// ncu (profiles/, r1j): refill 16 % + output 11 % of the kernel's time; frames that never leave iterate 25 % faster.
//
// Here nothing of that exists as a phase:
//   * every slot decision is CTA-UNIFORM and taken by every thread from the same shared flags right after the check
//     phase's barrier: which frames decoded (syndrome zero), which complete their last iteration in THIS variable phase
//     (the per-slot iteration counters live in registers of every thread), which ring entry each free slot takes (ring
//     entries are consumed in order, so `head` is a register too).  No extra barrier, no thread-0 serial section;
//   * the variable phase writes marg as whole 16-byte cells anyway.  A thread that owns position `item` writes the hard
//     decision of a leaving frame straight from the cell it holds in registers (decoded frames: the cell as it was
//     before this phase, one extra LDS.128), then overwrites that LANE of the cell with the new frame's prior, which it
//     converts itself from the staged row (the value of variable imap[item]); the prior cell is rewritten the same way.
//     All stores are the STS.128 of the phase: zero bank conflicts, zero extra wavefronts for marg;
//   * the check phase zeroes the register-resident c2v of fresh lanes under a uniform branch.
// Received rows land through cp.async.bulk + mbarrier (3 entries); thread 0 re-arms consumed entries after the phase's
// barrier.  Arithmetic, exit rules and outputs are resident_vp's, bit for bit (tests/test_gpu_parity.py runs both).
//
// Handles: regular codes on resident_vp's tables in the two-CTA geometry, priors / BSC / BIAWGN input of any row type
// the bulk copy can stage (16-byte aligned rows), no separate hard input (ldpc_decode with y_hard keeps resident_vp).
#pragma once
#include "resident_vd.cuh"
#include "resident_vp.cuh"

namespace ldpc {

constexpr int kVqRing = 3;

// The scalar type of the messages: float (4 frames per 16-byte cell) or double (2 frames per cell, the reference's own
// arithmetic — min-sum only, resident_vd's node rule).  A cell is a cell: layout, tables and placement do not change.
template <typename T> struct VqType;
template <> struct VqType<float> {
    using Cell = float4;
    static constexpr int F = 4;
    static __device__ __forceinline__ uint32_t signword(float v) { return f32_bits(v); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct VqType<double> {
    using Cell = double2;
    static constexpr int F = 2;
    static __device__ __forceinline__ uint32_t signword(double v) { return f64_hi(v); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};
template <typename C, typename T> __device__ __forceinline__ T &lane_of(C &c, int j) { return (&c.x)[j]; }

// One received value -> prior.  INMODE / INES >= 0: the input mode and element size are compile-time (the headline
// BIAWGN float32 rows and BSC bytes get their own instances); -1: read them from the parameters (res_llr's branches).
template <typename T, int INMODE, int INES>
__device__ __forceinline__ T vq_llr(const unsigned char *row, int v, int in_mode, int in_es, double param, double inv_param)
{
    uint32_t hbit;
    if (sizeof(T) == 8) {                                                // float64: the reference's expression with its division
        if (INMODE < 0) return (T)vd_llr(row, v, in_mode, in_es, param, &hbit);
        return (T)vd_llr(row, v, INMODE, INES, param, &hbit);
    }
    if (INMODE < 0) return (T)res_llr(row, v, in_mode, in_es, param, inv_param, &hbit);
    return (T)res_llr(row, v, INMODE, INES, param, inv_param, &hbit);
}

// The check rule at the message type: float32 min-sum / sum-product (ldpc_math.cuh), float64 min-sum (resident_vd.cuh).
template <int ALGO, int DC> __device__ __forceinline__ void vq_check(const float (&a)[DC], float (&o)[DC], float sat)
{
    if (ALGO == ALGO_MSA) cn_msa_lean<DC>(a, o);
    else cn_spa_sc<DC>(a, DC, o, sat);
}
template <int ALGO, int DC> __device__ __forceinline__ void vq_check(const double (&a)[DC], double (&o)[DC], float)
{
    cn_msa_lean_f64<DC>(a, o);
}

// INMODE / INES: see vq_llr.  IRR: the irregular instance of resident_vp (check degrees 2..DC <= 6, variable degrees
// 0..8, holes; one index word per edge, planes are prefixes of the positions, short checks padded with +inf cells).
// MAXT: 320 = two CTAs per SM; kVpBigThreads = ONE CTA per SM for regular codes whose 4 frames need the whole shared memory
// (n up to ~2850, the Margulis code: resident_vp's geometry — five variable passes, a third check pass whose c2v_old lives
// in the planes, the position -> variable map in global memory, one ring entry).
template <int ALGO, int DC, int DV, int TT, int NPC, int INMODE = -1, int INES = -1, bool IRR = false, typename TS = float, int MAXT = 320>
__global__ void __launch_bounds__(MAXT, MAXT > 320 ? 1 : 2) resident_vq(const ResParams p)
{
    static_assert(MAXT == 320 || (MAXT == kVpBigThreads && !IRR && TT == 0 && NPC == 0 && sizeof(TS) == 4), "two geometries");
    constexpr bool BIG = MAXT > 320;
    static_assert(DC >= 2 && DC <= 8 && DV >= 1 && (IRR ? (DV <= 8 && DC <= 6) : DV <= 3), "see resident_vp");
    static_assert(sizeof(TS) == 4 || ALGO == ALGO_MSA, "float64 on chip is min-sum only");
    using VT = VqType<TS>;
    using Cell = typename VT::Cell;
    constexpr int F = VT::F, CH = IRR ? DC : (DC + 1) / 2, VNP = BIG ? kVpBigVnPasses : kResVnPasses;
    constexpr uint32_t ALL = (1u << F) - 1u;
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int MPC = IRR ? NPC / 2 : NPC * DV / DC;
    const int np = NPC ? NPC : p.n, mp = NPC ? MPC : p.m;
    const uint32_t S = (uint32_t)np * 16u;
    const VpSmem L = IRR ? vx_smem_layout(np, p.plane_cells, p.ring, p.stage_stride) : vp_smem_layout(np, DV, p.ring, p.stage_stride, !BIG);
    Cell *marg = reinterpret_cast<Cell *>(smem + L.marg);
    Cell *planes = reinterpret_cast<Cell *>(smem + L.planes);
    Cell *prior = reinterpret_cast<Cell *>(smem + L.prior);
    unsigned char *stage = smem + L.stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    const uint16_t *imap = BIG ? p.vinvmap : reinterpret_cast<const uint16_t *>(smem + L.imap);       // variable at a position

    __shared__ int r_frame[kVqRing];
    __shared__ uint32_t s_unsat[2];

    const int tid = threadIdx.x, T = TT ? TT : (int)blockDim.x, lane = tid & 31;
    const bool have_hard = (p.in_mode == IN_BSC);                        // iteration-0 exit on the received bits (bpa.py:29)
    const int nref = NPC ? NPC : p.nref;
    const size_t row_bytes = (size_t)nref * p.in_es;
    const int R = p.ring;

    // ---- per-thread graph indices -> registers (once per CTA), as resident_vp
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
                    const uint32_t e = p.cw[(size_t)c * 8 + k];          // (position << 4) | (slot + 1)
                    if (k & 1) cw[ps][k >> 1] |= (e & 0xfff0u) << 16 | (e & 3u) << 2;
                    else cw[ps][k >> 1] |= e & 0xfff3u;
                }
            }
        }
    }
    auto goff = [&](int ps, int k) -> uint32_t {
        if (IRR) return vx_goff(cw[ps][k]);
        const uint32_t w = cw[ps][k >> 1];
        return (k & 1) ? vp_off1(w) : vp_off0(w);
    };
    Cell old[kResCnPasses][DC];                                          // c2v of the thread's own checks
#pragma unroll
    for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
        for (int k = 0; k < DC; ++k)
#pragma unroll
            for (int j = 0; j < F; ++j) (&old[ps][k].x)[j] = (TS)0;

    if (!BIG)
        for (int i = tid; i < np; i += T) reinterpret_cast<uint16_t *>(smem + L.imap)[i] = p.vinvmap[i];
    // BIG: the check of the third pass (position tid + 2 T), its gather / scatter byte offsets straight from the table
    const int ctail = tid + kResCnPasses * T;
    auto tail_offsets = [&](uint32_t (&g)[DC], uint32_t (&sc)[DC]) {
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            const uint32_t e = p.cw[(size_t)ctail * 8 + k];              // (position << 4) | (slot + 1)
            g[k] = e & 0xfff0u;
            sc[k] = (e & 3u) * S + g[k];
        }
    };
    if (IRR) {
        // cells nobody writes must read as +0.0 (short planes, holes), the padding cells behind marg as +inf
        Cell i4;
#pragma unroll
        for (int j = 0; j < F; ++j) (&i4.x)[j] = (TS)INFINITY;
        for (int i = tid; i < (int)(L.stage / 16); i += T) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (tid < 8) marg[np + tid] = i4;
    }
    auto issue = [&](int e) {                                            // thread 0: fetch the next frame into ring entry e
        const int g = atomicAdd(p.counter, 1);
        if (g < p.B) {
            r_frame[e] = g;
            mbar_expect_tx(&bars[e], (uint32_t)row_bytes);
            bulk_g2s(stage + (size_t)e * p.stage_stride, (const char *)p.src + (size_t)g * row_bytes, (uint32_t)row_bytes, &bars[e]);
        } else {
            r_frame[e] = -1;
        }
    };
    if (tid == 0) {
        s_unsat[0] = s_unsat[1] = 0u;
        for (int e = 0; e < R; ++e) mbar_init(&bars[e], 1u);
        fence_mbar_init();
        for (int e = 0; e < R; ++e) issue(e);
    }
    __syncthreads();

    // CTA-uniform slot state, identical in every thread
    uint32_t active = 0u, fresh = 0u;
    // A slot is in `run` in every round between moving in and leaving, so its iteration count is the number of rounds
    // since then: start_s = the round at which the frame moved in, gi = the current round.  next_max = the first round at
    // which some active slot reaches the iteration bound: until then, and while every syndrome stays non-zero, a round
    // needs no per-slot book-keeping at all.
    int start_s[F], fr_s[F];
#pragma unroll
    for (int j = 0; j < F; ++j) { start_s[j] = 0; fr_s[j] = 0; }
    int gi = 0, next_max = 0x7fffffff;
    int head_e = 0;                                                      // ring entry the next frame comes from ...
    uint32_t head_par = 0u;                                              // ... and the phase parity of its mbarrier
    bool more = true;                                                    // the ring may still deliver frames
    int par = 0;

    for (;;) {
        // ======================================= check-node phase =======================================
        uint32_t unsat = 0u;
        if (active != 0u) {
            if (fresh != 0u) {                                           // new frames start from c2v = 0
#pragma unroll
                for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
                    for (int k = 0; k < DC; ++k)
#pragma unroll
                        for (int j = 0; j < F; ++j)
                            if ((fresh >> j) & 1u) (&old[ps][k].x)[j] = (TS)0;
                if (BIG && ctail < mp) {                                 // third-pass checks keep their c2v in the planes only
                    uint32_t g[DC], sc[DC];
                    tail_offsets(g, sc);
#pragma unroll
                    for (int k = 0; k < DC; ++k)
#pragma unroll
                        for (int j = 0; j < F; ++j)
                            if ((fresh >> j) & 1u) reinterpret_cast<TS *>(smem + sc[k])[j] = (TS)0;
                }
            }
#pragma unroll
            for (int ps = 0; ps < kResCnPasses; ++ps) {
                if (tid + ps * T < mp) {
                    Cell mg[DC];
#pragma unroll
                    for (int k = 0; k < DC; ++k) mg[k] = *reinterpret_cast<const Cell *>(smem + goff(ps, k));
                    uint32_t sx[F];
#pragma unroll
                    for (int j = 0; j < F; ++j) sx[j] = 0u;
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
#pragma unroll
                        for (int j = 0; j < F; ++j) {
                            const TS mv = (&mg[k].x)[j];
                            sx[j] ^= VT::signword(mv);                   // sign bit == (marg < 0), see resident_vp
                            (&mg[k].x)[j] = VT::sub(mv, (&old[ps][k].x)[j]);         // v2c = marg - c2v_old (bpa.py:37)
                        }
                    }
                    uint32_t syn = 0u;
#pragma unroll
                    for (int j = 0; j < F; ++j) syn |= (sx[j] >> 31) << j;
                    // irregular, sum-product: a padding edge reads +inf (neutral), its own output is forced to 0 (resident_vp)
                    const int dcr = (IRR && ALGO != ALGO_MSA) ? (int)(cw[ps][0] & 15u) : DC;
#pragma unroll
                    for (int j = 0; j < F; ++j) {
                        TS a[DC], o[DC];
#pragma unroll
                        for (int k = 0; k < DC; ++k) a[k] = (&mg[k].x)[j];
                        vq_check<ALGO, DC>(a, o, p.sat_llr);
#pragma unroll
                        for (int k = 0; k < DC; ++k) (&old[ps][k].x)[j] = (IRR && ALGO != ALGO_MSA && k >= 2 && k >= dcr) ? (TS)0 : o[k];
                    }
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        uint32_t coff;
                        if (IRR) {
                            coff = vx_soff(cw[ps][k]);
                        } else {
                            const uint32_t w = cw[ps][k >> 1];
                            coff = (k & 1) ? vp_sl1x4(w) * (S >> 2) + vp_off1(w) : vp_sl0(w) * S + vp_off0(w);
                        }
                        *reinterpret_cast<Cell *>(smem + coff) = old[ps][k];
                    }
                    unsat |= syn;
                }
            }
            if (BIG && ctail < mp) {                                     // third pass: c2v_old comes back from the planes
                uint32_t g[DC], sc[DC];
                tail_offsets(g, sc);
                Cell mg[DC], ol[DC];
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    mg[k] = *reinterpret_cast<const Cell *>(smem + g[k]);
                    ol[k] = *reinterpret_cast<const Cell *>(smem + sc[k]);
                }
                uint32_t sx[F];
#pragma unroll
                for (int j = 0; j < F; ++j) sx[j] = 0u;
#pragma unroll
                for (int k = 0; k < DC; ++k)
#pragma unroll
                    for (int j = 0; j < F; ++j) {
                        const TS mv = (&mg[k].x)[j];
                        sx[j] ^= VT::signword(mv);
                        (&mg[k].x)[j] = VT::sub(mv, (&ol[k].x)[j]);
                    }
#pragma unroll
                for (int j = 0; j < F; ++j) {
                    TS a[DC], o[DC];
#pragma unroll
                    for (int k = 0; k < DC; ++k) a[k] = (&mg[k].x)[j];
                    vq_check<ALGO, DC>(a, o, p.sat_llr);
#pragma unroll
                    for (int k = 0; k < DC; ++k) (&ol[k].x)[j] = o[k];
                }
#pragma unroll
                for (int k = 0; k < DC; ++k) *reinterpret_cast<Cell *>(smem + sc[k]) = ol[k];
#pragma unroll
                for (int j = 0; j < F; ++j) unsat |= (sx[j] >> 31) << j;
            }
            unsat = __reduce_or_sync(kFull, unsat);
            if (lane == 0 && unsat != 0u) atomicOr(&s_unsat[par], unsat);
        }
        __syncthreads();

        auto vn_item = [&](int item, Cell &pr, Cell &mgv) {
            Cell sm;
            if (IRR) {
                // plane k = a prefix of the positions (descending degree): ascending edge order, bpa.py:35
#pragma unroll
                for (int j = 0; j < F; ++j) (&sm.x)[j] = (TS)0;
                if (item < p.pcnt[0]) sm = *reinterpret_cast<const Cell *>(smem + p.pbase[0] + (size_t)item * 16);
#pragma unroll
                for (int k = 1; k < DV; ++k) {
                    if (item >= p.pcnt[k]) break;
                    const Cell c = *reinterpret_cast<const Cell *>(smem + p.pbase[k] + (size_t)item * 16);
#pragma unroll
                    for (int j = 0; j < F; ++j) (&sm.x)[j] = VT::add((&sm.x)[j], (&c.x)[j]);
                }
            } else {
                Cell c[DV];
#pragma unroll
                for (int k = 0; k < DV; ++k) c[k] = planes[(size_t)k * np + item];
#pragma unroll
                for (int j = 0; j < F; ++j) {
                    TS s = (&c[0].x)[j];
#pragma unroll
                    for (int k = 1; k < DV; ++k) s = VT::add(s, (&c[k].x)[j]);
                    (&sm.x)[j] = s;
                }
            }
            pr = prior[item];
#pragma unroll
            for (int j = 0; j < F; ++j) (&mgv.x)[j] = VT::add((&pr.x)[j], (&sm.x)[j]);      // bpa.py:35
        };
        // ======================================= slot decisions (every thread, uniform) =======================================
        ++gi;
        uint32_t us = __reduce_or_sync(kFull, s_unsat[par]);
        if (!have_hard) us |= fresh;                                     // a new frame's marg is its prior: no syndrome test yet
        if (tid == 0) s_unsat[par ^ 1] = 0u;
        if (active != 0u && (active & ~us) == 0u && gi < next_max && (active == ALL || !more)) {
            // ---- the quiet round (two out of three): nobody decoded, nobody at the bound, no free slot to fill
#pragma unroll
            for (int ps = 0; ps < VNP; ++ps) {
                const int item = tid + ps * T;
                if (item < np) {
                    Cell pr, mgv;
                    vn_item(item, pr, mgv);
                    marg[item] = mgv;
                }
            }
            __syncthreads();
            fresh = 0u;
            par ^= 1;
            continue;
        }
        const uint32_t decoded = active & ~us;                           // leaves with its iteration count unchanged (bpa.py:29)
        const uint32_t run = active & us;
        uint32_t maxed = 0u;
#pragma unroll
        for (int s = 0; s < F; ++s)
            if (((run >> s) & 1u) && gi - start_s[s] >= p.limit) maxed |= 1u << s;   // bpa.py:63, then bpa.py:28 at the top of the next round
        const uint32_t leaving = decoded | maxed;
        if (tid < F && ((leaving >> tid) & 1u)) {
            int g = fr_s[0], st = start_s[0];
#pragma unroll
            for (int s = 1; s < F; ++s)
                if (tid == s) { g = fr_s[s]; st = start_s[s]; }
            const bool dec = ((decoded >> tid) & 1u) != 0u;
            p.iters[g] = gi - st - (dec ? 1 : 0);                        // a decoded frame did not run this round
            if (p.reason != nullptr) p.reason[g] = (uint8_t)(dec ? LDPC_REASON_DECODED : p.bound_reason);
        }
        // free slots take the next landed rows, in slot order
        uint32_t inst = 0u;
        int ent[F], nfr[F];                                              // only read under the matching bit of `inst`
        const uint32_t freem = (~active | leaving) & ALL;
        if (freem != 0u && more) {
            int taken = 0;
#pragma unroll
            for (int s = 0; s < F; ++s) {
                if (((freem >> s) & 1u) && more && taken < R) {
                    const int g = r_frame[head_e];
                    if (g < 0) {
                        more = false;
                    } else {
                        mbar_wait(&bars[head_e], head_par);
                        ent[s] = head_e; nfr[s] = g;
                        inst |= 1u << s;
                        ++taken;
                        if (++head_e == R) { head_e = 0; head_par ^= 1u; }
                    }
                }
            }
        }
        if (active == 0u && inst == 0u) break;                           // nothing running, nothing left to start

        // ======================================= variable-node phase =======================================
        if ((leaving | inst) == 0u) {                                    // the common iteration: nothing but the sums
#pragma unroll
            for (int ps = 0; ps < VNP; ++ps) {
                const int item = tid + ps * T;
                if (item < np) {
                    Cell pr, mgv;
                    vn_item(item, pr, mgv);
                    marg[item] = mgv;
                }
            }
        } else {                                                         // a frame leaves and / or a new one moves in
            uint8_t *dst[F];
            const unsigned char *row[F];
#pragma unroll
            for (int j = 0; j < F; ++j) {
                dst[j] = p.x_hat + (size_t)fr_s[j] * nref;
                row[j] = stage + (size_t)ent[j] * p.stage_stride;
            }
#pragma unroll 1
            for (int ps = 0; ps < VNP; ++ps) {
                const int item = tid + ps * T;
                if (item >= np) break;
                Cell pr, mgv;
                vn_item(item, pr, mgv);
                const uint32_t v = imap[item];
                if (IRR && v == 0xffffu) {                               // a position without a variable: nothing leaves, nothing moves in
                    marg[item] = mgv;
                    continue;
                }
                if (decoded != 0u) {                                     // word = the marginal this frame's last check phase saw
                    const Cell om = marg[item];
#pragma unroll
                    for (int j = 0; j < F; ++j)
                        if ((decoded >> j) & 1u) dst[j][v] = (uint8_t)(VT::signword((&om.x)[j]) >> 31);
                }
#pragma unroll
                for (int j = 0; j < F; ++j)
                    if ((maxed >> j) & 1u) dst[j][v] = (uint8_t)(VT::signword((&mgv.x)[j]) >> 31);
                if (inst != 0u) {
#pragma unroll
                    for (int j = 0; j < F; ++j)
                        if ((inst >> j) & 1u) {
                            const TS val = vq_llr<TS, INMODE, INES>(row[j], (int)v, p.in_mode, p.in_es, p.param, p.inv_param);
                            (&mgv.x)[j] = val;
                            (&pr.x)[j] = val;
                        }
                    prior[item] = pr;
                }
                marg[item] = mgv;
            }
        }
        __syncthreads();
        if (tid == 0 && inst != 0u) {                                    // the staged rows are consumed: fetch the next frames
#pragma unroll
            for (int s = 0; s < F; ++s)
                if ((inst >> s) & 1u) issue(ent[s]);
        }
#pragma unroll
        for (int s = 0; s < F; ++s)
            if ((inst >> s) & 1u) { fr_s[s] = nfr[s]; start_s[s] = gi; }
        active = (run & ~maxed) | inst;
        fresh = inst;
        next_max = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < F; ++s)
            if ((active >> s) & 1u) next_max = min(next_max, start_s[s] + p.limit);
        par ^= 1;
    }
}

}  // namespace ldpc
