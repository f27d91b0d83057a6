// resident_vd.cuh — the variable-plane on-chip decoder (resident_vp.cuh) in FLOAT64, min-sum, regular codes.
//
// float64 is the reference's own arithmetic (bpa.py runs on whatever dtype the priors have, and its channel front ends
// produce float64): decoding in it reproduces the reference's words and iteration counts bit for bit, not just the
// float32 oracle's.  The streaming sweeps do that at the HBM roofline of 8-byte messages (4.8 M frames/s on
// LDPC(1200,3,6)); here the frames stay on chip exactly as in resident_vp:
//
//   * a 16-byte cell is a double2 = TWO frames, a CTA owns 2 slots, two CTAs share an SM; marg / plane / prior arrays,
//     byte offsets, the packed index words and the conflict-free placement are those of resident_vp (same ResParams,
//     same tables — a cell is a cell);
//   * check node: v2c = marg - c2v_old in float64, lean min-sum on doubles (eight 3-input minima for degree 6, signs
//     handled on the high words: exact operations only, so the result is the reference's in any evaluation order);
//   * variable node: marg = prior + ((c0 + c1) + c2) in ascending edge order (bpa.py:35);
//   * refill: priors exactly as the reference computes them, (-2 y) / noise_var with the float64 division
//     (biawgn.py:28), L (1 - 2 y) for BSC (bsc.py:25), or the caller's float64 priors; -0.0 folded to +0.0.
// Exit rules, ring, frame dispenser and outputs are resident_vp's with 2-bit slot masks.  IRR = true is the irregular
// instance (planes as prefixes of the degree-sorted positions, one index word per edge, checks padded with +inf reads and
// scratch writes), exactly as in resident_vp.
#pragma once
#include "resident_vp.cuh"

namespace ldpc {

__device__ __forceinline__ uint32_t f64_hi(double d) { return (uint32_t)__double2hiint(d); }
__device__ __forceinline__ double f64_with_hi(double d, uint32_t hi) { return __hiloint2double((int)hi, __double2loint(d)); }

// Lean float64 min-sum for a check of exactly DC edges; see cn_msa_lean (ldpc_math.cuh) for the construction and the
// precondition (no NaN, no -0.0 input: guaranteed by the kernel the same way).  mag >= +0, so OR-ing the sign bit into
// its high word is the reference's sign * mag including (-1) * 0 = -0.0.
template <int DC>
__device__ __forceinline__ void cn_msa_lean_f64(const double (&v)[DC], double (&out)[DC])
{
    double a[DC], mag[DC];
#pragma unroll
    for (int k = 0; k < DC; ++k) a[k] = fabs(v[k]);
    if (DC == 6) {
        const double mL = fmin(fmin(a[0], a[1]), a[2]), mR = fmin(fmin(a[3], a[4]), a[5]);
        mag[0] = fmin(fmin(a[1], a[2]), mR); mag[1] = fmin(fmin(a[0], a[2]), mR); mag[2] = fmin(fmin(a[0], a[1]), mR);
        mag[3] = fmin(fmin(a[4], a[5]), mL); mag[4] = fmin(fmin(a[3], a[5]), mL); mag[5] = fmin(fmin(a[3], a[4]), mL);
    } else {
        double pre[DC], suf[DC];
        pre[0] = (double)INFINITY;
#pragma unroll
        for (int k = 1; k < DC; ++k) pre[k] = fmin(pre[k - 1], a[k - 1]);
        suf[DC - 1] = (double)INFINITY;
#pragma unroll
        for (int k = DC - 2; k >= 0; --k) suf[k] = fmin(suf[k + 1], a[k + 1]);
#pragma unroll
        for (int k = 0; k < DC; ++k) mag[k] = fmin(pre[k], suf[k]);
    }
    uint32_t x = 0u;
#pragma unroll
    for (int k = 0; k < DC; ++k) x ^= f64_hi(v[k]);
    const uint32_t xs = x & 0x80000000u;
#pragma unroll
    for (int k = 0; k < DC; ++k) out[k] = f64_with_hi(mag[k], f64_hi(mag[k]) | (xs ^ (f64_hi(v[k]) & 0x80000000u)));
}

// One received value -> float64 prior, the reference's expression (io_kernels.cuh llr_map<., double, .>), -0.0 folded.
__device__ __forceinline__ double vd_llr(const void *row, int v, int in_mode, int in_es, double param, uint32_t *hard)
{
    *hard = 0u;
    double val;
    if (in_mode == IN_BSC) {
        const uint8_t y = ((const uint8_t *)row)[v];
        *hard = (uint32_t)(y != 0);
        val = param * (double)(1 - 2 * (int)y);
    } else {
        const double y = res_in_f64(row, v, in_es);
        val = (in_mode == IN_BIAWGN) ? __ddiv_rn(-2.0 * y, param) : y;
    }
    return __dadd_rn(val, 0.0);
}

// DC, DV, TT, NPC, IRR as in resident_vp.
template <int DC, int DV, int TT, int NPC, bool IRR = false>
__global__ void __launch_bounds__(320, 2) resident_vd(const ResParams p)
{
    static_assert(DC >= 2 && DC <= 8 && DV >= 1 && (IRR ? (DV <= 8 && DC <= 6) : DV <= 3), "see resident_vp");
    constexpr int F = 2, CH = IRR ? DC : (DC + 1) / 2;
    constexpr uint32_t ALL = 0x3u;
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int MPC = IRR ? NPC / 2 : NPC * DV / DC;
    const int np = NPC ? NPC : p.n, mp = NPC ? MPC : p.m;
    const uint32_t S = (uint32_t)np * 16u;                            // bytes per plane
    const VpSmem L = IRR ? vx_smem_layout(np, p.plane_cells, p.ring, p.stage_stride) : vp_smem_layout(np, DV, p.ring, p.stage_stride);
    double2 *marg = reinterpret_cast<double2 *>(smem + L.marg);
    double2 *planes = reinterpret_cast<double2 *>(smem + L.planes);
    double2 *prior = reinterpret_cast<double2 *>(smem + L.prior);
    unsigned char *stage = smem + L.stage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint8_t *hb = smem + L.hb;
    uint16_t *imap = reinterpret_cast<uint16_t *>(smem + L.imap);

    __shared__ int s_frame[F], s_it[F], s_assign[F];
    __shared__ int r_frame[kResRingMax], r_uses[kResRingMax];
    __shared__ uint32_t s_unsat[2], s_maxed[2], s_unsat0, s_newmask, s_exhausted;

    const int tid = threadIdx.x, T = TT ? TT : (int)blockDim.x, lane = tid & 31;
    const bool async = p.ring > 0;
    const bool have_hard = (p.in_mode == IN_BSC) || (p.in_mode == IN_COPY && p.y_hard != nullptr);
    const int nref = NPC ? NPC : p.nref;                             // == np
    const size_t row_bytes = (size_t)nref * p.in_es;

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
    auto goff = [&](int ps, int k) -> uint32_t {
        if (IRR) return vx_goff(cw[ps][k]);
        const uint32_t w = cw[ps][k >> 1];
        return (k & 1) ? vp_off1(w) : vp_off0(w);
    };
    double2 old[kResCnPasses][DC];                                   // c2v of the thread's own checks
#pragma unroll
    for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
        for (int k = 0; k < DC; ++k) old[ps][k] = make_double2(0.0, 0.0);

    for (int i = tid; i < np; i += T) imap[i] = p.vinvmap[i];
    if (IRR) {
        // cells nobody writes must read as +0.0 (short planes, holes), the padding cells behind marg as +inf
        const double2 z2 = make_double2(0.0, 0.0);
        for (int i = tid; i < (int)(L.stage / 16); i += T) reinterpret_cast<double2 *>(smem)[i] = z2;
        for (int i = tid; i < np + 8; i += T) hb[i] = 0;
        __syncthreads();
        if (tid < 8) marg[np + tid] = make_double2((double)INFINITY, (double)INFINITY);
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
    int head = 0;
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

    // A leaving frame's word = sign bits of its lane of marg (bpa.py:62; NaN and 0 -> bit 0).
    auto output_bits = [&](uint32_t mask, int why) {
        uint32_t mq = mask & ALL;
        while (mq != 0u) {
            const int j = __ffs(mq) - 1;
            mq &= mq - 1u;
            uint8_t *dst = p.x_hat + (size_t)s_frame[j] * nref;
            const uint32_t *mj = reinterpret_cast<const uint32_t *>(marg) + 2 * j + 1;      // high word of lane j
#pragma unroll
            for (int ps = 0; ps < kResVnPasses; ++ps) {
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
            const uint32_t nm = __reduce_or_sync(kFull, s_newmask);
            exhausted = __reduce_or_sync(kFull, s_exhausted) != 0u;
            if (nm == 0u) break;

            // ---- received rows -> prior / marg columns, one new slot at a time
            for (int s = 0; s < F; ++s) {
                if (!((nm >> s) & 1u)) continue;
                const int g = s_frame[s];
                const unsigned char *row;
                if (async) {
                    const int e = s_assign[s];
                    mbar_wait(&bars[e], (uint32_t)(r_uses[e] & 1));
                    row = stage + (size_t)e * p.stage_stride;
                } else {
                    row = (const unsigned char *)p.src + (size_t)g * row_bytes;
                }
                const uint8_t *hrow = (p.in_mode == IN_COPY && p.y_hard != nullptr) ? p.y_hard + (size_t)g * nref : nullptr;
                double *mcol = reinterpret_cast<double *>(marg) + s;
                double *pcol = reinterpret_cast<double *>(prior) + s;
                for (int v = tid; v < nref; v += T) {
                    uint32_t hbit;
                    const double val = vd_llr(row, v, p.in_mode, p.in_es, p.param, &hbit);
                    const uint32_t pos = p.vposmap[v];
                    mcol[(size_t)pos * 2] = val;
                    pcol[(size_t)pos * 2] = val;
                    if (have_hard) {
                        if (hrow != nullptr) hbit = (uint32_t)(hrow[v] != 0);
                        hb[pos] = (uint8_t)((hb[pos] & ~(1u << s)) | (hbit << s));       // the same thread owns hb[pos] for every slot
                    }
                }
            }
            // ---- the new frames start from c2v = 0
#pragma unroll
            for (int ps = 0; ps < kResCnPasses; ++ps)
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    if (nm & 1u) old[ps][k].x = 0.0;
                    if (nm & 2u) old[ps][k].y = 0.0;
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
                double2 mg[DC];
#pragma unroll
                for (int k = 0; k < DC; ++k) mg[k] = *reinterpret_cast<const double2 *>(smem + goff(ps, k));
                // v2c = marg - c2v_old (bpa.py:37); the sign bits of marg are the current hard decisions (bpa.py:62)
                uint32_t sx[2] = {0u, 0u};
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    sx[0] ^= f64_hi(mg[k].x);
                    sx[1] ^= f64_hi(mg[k].y);
                    mg[k].x = __dsub_rn(mg[k].x, old[ps][k].x);
                    mg[k].y = __dsub_rn(mg[k].y, old[ps][k].y);
                }
                const uint32_t syn = (sx[0] >> 31) | ((sx[1] >> 31) << 1);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double a[DC], o[DC];
#pragma unroll
                    for (int k = 0; k < DC; ++k) a[k] = j ? mg[k].y : mg[k].x;
                    cn_msa_lean_f64<DC>(a, o);
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        if (j) old[ps][k].y = o[k];
                        else old[ps][k].x = o[k];
                    }
                }
                // scatter: plane (slot) of the variable's edge, same bank group as the gather of the same step
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    uint32_t coff;
                    if (IRR) {
                        coff = vx_soff(cw[ps][k]);
                    } else {
                        const uint32_t w = cw[ps][k >> 1];
                        coff = (k & 1) ? vp_sl1x4(w) * (S >> 2) + vp_off1(w) : vp_sl0(w) * S + vp_off0(w);
                    }
                    *reinterpret_cast<double2 *>(smem + coff) = old[ps][k];
                }
                unsat |= syn;
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
        if (decoded != 0u) output_bits(decoded, LDPC_REASON_DECODED);     // marg still holds the last variable phase

        // ======================================= variable-node phase =======================================
#pragma unroll
        for (int ps = 0; ps < kResVnPasses; ++ps) {
            const int item = tid + ps * T;
            if (IRR) {
                if (item < np) {
                    double2 sm = make_double2(0.0, 0.0);
                    if (item < p.pcnt[0]) sm = *reinterpret_cast<const double2 *>(smem + p.pbase[0] + (size_t)item * 16);
#pragma unroll
                    for (int k = 1; k < DV; ++k) {
                        if (item >= p.pcnt[k]) break;
                        const double2 c = *reinterpret_cast<const double2 *>(smem + p.pbase[k] + (size_t)item * 16);
                        sm.x = __dadd_rn(sm.x, c.x); sm.y = __dadd_rn(sm.y, c.y);
                    }
                    const double2 pr = prior[item];
                    marg[item] = make_double2(__dadd_rn(pr.x, sm.x), __dadd_rn(pr.y, sm.y));
                }
            } else if (item < np) {
                double2 c[DV];
#pragma unroll
                for (int k = 0; k < DV; ++k) c[k] = planes[(size_t)k * np + item];
                const double2 pr = prior[item];
                double sx0 = c[0].x, sx1 = c[0].y;                                     // 0 + c0: see resident_vp.cuh
#pragma unroll
                for (int k = 1; k < DV; ++k) { sx0 = __dadd_rn(sx0, c[k].x); sx1 = __dadd_rn(sx1, c[k].y); }
                marg[item] = make_double2(__dadd_rn(pr.x, sx0), __dadd_rn(pr.y, sx1)); // bpa.py:35
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
