// resident_bec.cuh — on-chip BEC erasure decoding (bec.SPA.decode, /root/reference/src/bec.py:83-122) for short codes.
//
// The streaming erasure sweeps (stream_bec.cuh) move every message plane through HBM twice per iteration and are
// latency-limited at ~0.6 of the DRAM rate; but a frame's whole decoder state is 2 bits x (E + 2 n) — 1.5 KB for
// LDPC(1200,3,6) — so 64 frames fit in half an SM's shared memory.  This kernel keeps a TILE of 64 frames on chip for
// all their iterations: HBM sees the symbols once ([B][n] bytes in) and the words once (out).
//
// Same literal integer message passing as the streaming kernels (NOT peeling: arbitrary, even inconsistent, symbol
// inputs reproduce the reference bit for bit), bit-sliced: a message in {-1, 0, +1} is (nz, pos), one bit per frame.
// A 16-byte shared-memory CELL holds (nz, pos) of TWO 32-frame words = 64 frames: one LDS.128 / STS.128 per edge.
//
// Layout = the variable-plane layout of resident_vp.cuh, cell for cell, so its placement tables are reused as they are
// (res_layout.h: checks in file order, variables 8-coloured so that every gather step of a quarter-warp hits 8 bank groups):
//   x     [np]      (xe0, xv0, xe1, xv1)   current word: erased plane, value plane          (where resident_vp keeps marg)
//   msg   [dv][np]  plane s, position v = the message on the s-th edge of variable v, updated IN PLACE:
//                   v2c after the variable phase, c2v after the check phase
//   prior [np]      messages[y] (bec.py:76,85)
// CN  thread = (check, 64 frames): gathers its dc message cells through the packed index words, erasure count
//     saturating at 2 + parity of the +1 votes (bec.py:100-112), scatters the c2v back to the same cells.
// VN  thread = (variable position, 64 frames): its dv cells are CONTIGUOUS per plane; marginal and the dv leave-one-out
//     signs as boolean functions (ldpc_math.cuh bec_vn3; irregular codes: bit-sliced integers, any degree <= 8);
//     x_new = symbols[sign(marginal)] merged under the run mask; `changed` / `has erasures` flags per frame.
// Book-keeping of a round (stream_bec.cuh bec_book) runs in warps 0 / 1 (one per word, lane = frame) while the other
// warps already execute the next check phase: two barriers per iteration.
//
// No per-frame hand-over: a tile runs until its last frame stops (or the iteration bound) — at the reference's
// operating points (max_iter 10) nearly every frame runs every iteration.  Symbols are transposed into bit planes
// inside the kernel: 32 rows land in the (not yet used) message region with a flat coalesced copy, then warp ballots
// build the planes; the words leave the same way in reverse.
#pragma once
#include "resident_vp.cuh"

namespace ldpc {

struct BecResParams {
    int np, mp, nref;
    const uint16_t *cw;          // regular codes   [mp][8]: (variable position << 4) | (edge rank at the variable + 1)
    const uint32_t *cwx;         // irregular codes [mp][8]: (message cell << 16) | (variable position << 4) [| degree, k = 0]
    const uint16_t *vposmap;     // [nref] position of variable v
    const uint8_t *vdeg;         // irregular codes [np]: degree of the variable at a position, 0xff = hole
    int pcnt[8], pbase[8];       // irregular codes: cells of plane k (a prefix of the positions), its byte offset
    int plane_cells;
    const uint8_t *y;            // [B][nref] symbols {0, 1, 2}
    int B, limit, bound_reason;
    uint8_t *x_hat;
    int *iters;
    uint8_t *reason;
    int *counter;                // tile dispenser (zeroed before launch)
};

struct BecSmem {
    size_t x, planes, planes_bytes, prior, vpos, total;
};
__host__ __device__ inline BecSmem bec_smem_layout(int np, int plane_cells, bool irr, int nref)
{
    BecSmem L;
    size_t o = 0;
    L.x = o;      o += ((size_t)np + (irr ? 8 : 0)) * 16;                  // irregular: == vx_planes_offset(np)
    L.planes = o; L.planes_bytes = irr ? ((size_t)plane_cells + 8) * 16 : (size_t)3 * np * 16;
    o += L.planes_bytes;
    L.prior = o;  o += (size_t)np * 16;
    L.vpos = o;   o += ((size_t)nref * 2 + 15) / 16 * 16;
    L.total = o + 16;
    return L;
}

// nbytes from src to dst by the whole CTA; 128-bit accesses when everything is 16-byte aligned.
__device__ __forceinline__ void bec_copy(void *dst, const void *src, size_t nbytes, int tid, int T)
{
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | nbytes) & 15u) == 0) {
        const uint4 *s = reinterpret_cast<const uint4 *>(src);
        uint4 *d = reinterpret_cast<uint4 *>(dst);
        for (size_t i = tid; i < nbytes / 16; i += T) d[i] = s[i];
    } else if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | nbytes) & 3u) == 0) {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d = reinterpret_cast<uint32_t *>(dst);
        for (size_t i = tid; i < nbytes / 4; i += T) d[i] = s[i];
    } else {
        const uint8_t *s = reinterpret_cast<const uint8_t *>(src);
        uint8_t *d = reinterpret_cast<uint8_t *>(dst);
        for (size_t i = tid; i < nbytes; i += T) d[i] = s[i];
    }
}

// IRR = false: regular code, every check 6 edges, every variable 3 (resident_vp's `cw` tables).
// IRR = true : check degrees 2..6, variable degrees 0..8, holes (resident_vp's irregular `cwx` tables).
// nref % 4 == 0 and 32 * nref <= planes_bytes (the staging area of the transposes) are checked by the host.
template <bool IRR>
__global__ void __launch_bounds__(320, 2) resident_bec(const BecResParams p)
{
    constexpr int DC = 6, DV = IRR ? 8 : 3, CH = IRR ? DC : 3;
    extern __shared__ __align__(128) unsigned char smem[];
    const int np = p.np, mp = p.mp, n = p.nref, n4 = n >> 2;
    const uint32_t S = (uint32_t)np * 16u;
    const BecSmem L = bec_smem_layout(np, p.plane_cells, IRR, n);
    uint4 *xc = reinterpret_cast<uint4 *>(smem + L.x);
    uint4 *prior = reinterpret_cast<uint4 *>(smem + L.prior);
    unsigned char *stage = smem + L.planes;                              // 32 rows of symbols, before / after the decode
    uint16_t *s_vpos = reinterpret_cast<uint16_t *>(smem + L.vpos);

    __shared__ uint32_t s_act[2], s_chg[2], s_has[2], s_stop[2];
    __shared__ int s_iters[64];
    __shared__ int s_tile;

    const int tid = threadIdx.x, T = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    // ---- per-thread graph indices -> registers (once per CTA), packed exactly as in resident_vp
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
                    const uint32_t e = p.cw[(size_t)c * 8 + k];
                    if (k & 1) cw[ps][k >> 1] |= (e & 0xfff0u) << 16 | (e & 3u) << 2;
                    else cw[ps][k >> 1] |= e & 0xfff3u;
                }
            }
        }
    }
    auto cell_off = [&](int ps, int k) -> uint32_t {                     // byte offset of the message cell of edge k
        if (IRR) return vx_soff(cw[ps][k]);
        const uint32_t w = cw[ps][k >> 1];
        return (k & 1) ? vp_sl1x4(w) * (S >> 2) + vp_off1(w) : vp_sl0(w) * S + vp_off0(w);
    };
    for (int i = tid; i < n; i += T) s_vpos[i] = p.vposmap[i];

    for (;;) {
        __syncthreads();                                                 // the previous tile is out; s_vpos is visible
        if (tid == 0) {
            s_tile = atomicAdd(p.counter, 1);
            s_has[0] = s_has[1] = 0u;
        }
        __syncthreads();
        const long long f0 = (long long)s_tile * 64;
        if (f0 >= p.B) break;
        int nvalid[2];
        uint32_t okw[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const long long left = (long long)p.B - (f0 + 32 * w);
            nvalid[w] = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
            okw[w] = nvalid[w] == 32 ? 0xffffffffu : ((1u << nvalid[w]) - 1u);
        }

        // ================================ symbols in: rows -> bit planes (x cells), one word at a time ================
#pragma unroll 1
        for (int w = 0; w < 2; ++w) {
            if (w) __syncthreads();                                      // the ballots of word 0 have read the staging area
            bec_copy(stage, p.y + (size_t)(f0 + 32 * w) * n, (size_t)nvalid[w] * n, tid, T);
            __syncthreads();
            const bool valid = lane < nvalid[w];
            uint32_t er_any = 0u;
            for (int q = warp; q < n4; q += nwarps) {
                const uint32_t wv = valid ? reinterpret_cast<const uint32_t *>(stage)[(size_t)lane * n4 + q] : 0u;
                uint32_t e[4], o[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t sy = (wv >> (8 * b)) & 0xffu;
                    e[b] = __ballot_sync(kFull, sy >= 2u);
                    o[b] = __ballot_sync(kFull, sy == 1u);
                }
                er_any |= e[0] | e[1] | e[2] | e[3];
                if (lane < 4) {
                    const uint32_t ee = lane == 0 ? e[0] : lane == 1 ? e[1] : lane == 2 ? e[2] : e[3];
                    const uint32_t oo = lane == 0 ? o[0] : lane == 1 ? o[1] : lane == 2 ? o[2] : o[3];
                    const uint32_t pos = s_vpos[4 * q + lane];
                    reinterpret_cast<uint2 *>(xc + pos)[w] = make_uint2(ee, oo);      // (xe, xv) of this word
                }
            }
            if (lane == 0 && er_any != 0u) atomicOr(&s_has[w], er_any);
        }
        __syncthreads();
        // priors = messages[y] (bec.py:85); v2c = priors[yy] (bec.py:86): the prior goes into every edge cell of the variable
#pragma unroll
        for (int ps = 0; ps < kResVnPasses; ++ps) {
            const int item = tid + ps * T;
            if (item < np) {
                int d = DV;
                if (IRR) d = p.vdeg[item];
                if (!IRR || d != 0xff) {
                    const uint4 x = xc[item];
                    const uint4 pr = make_uint4(okw[0] & ~x.x, x.y, okw[1] & ~x.z, x.w);
                    prior[item] = pr;
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        if (IRR && k >= d) break;
                        if (IRR) *reinterpret_cast<uint4 *>(smem + p.pbase[k] + (size_t)item * 16) = pr;
                        else *reinterpret_cast<uint4 *>(smem + (size_t)(k + 1) * S + (size_t)item * 16) = pr;
                    }
                }
            }
        }
        if (tid < 64) s_iters[tid] = 0;
        if (tid < 2) {
            s_act[tid] = okw[tid] & s_has[tid];                          // no erasures: 'decoded' at iteration 0 (bec.py:97)
            s_has[tid] = 0u;
            s_chg[tid] = 0u;
            s_stop[tid] = 0u;
        }
        __syncthreads();

        // ================================ iterations ================================
        int it = 0;
        for (;;) {
            // ---- book-keeping of the round that just ended (bec_book), warp w <-> word w, lane = frame;
            //      the other warps are already in the check phase, which does not depend on it
            if (it > 0 && warp < 2) {
                uint32_t a = s_act[warp];
                const uint32_t ch = s_chg[warp], h = s_has[warp];
                const uint32_t st = a & ~ch;                             // x_new == x_hat: 'stopping', iter_count stays (bec.py:120)
                a &= ch;
                if ((a >> lane) & 1u) s_iters[warp * 32 + lane] += 1;    // bec.py:122
                if (it != p.limit) a &= h;                               // no erasures left: 'decoded' (bec.py:97); the bound is tested first (bec.py:96)
                __syncwarp();
                if (lane == 0) {
                    s_act[warp] = a;
                    s_chg[warp] = 0u;
                    s_has[warp] = 0u;
                    if (st != 0u) s_stop[warp] |= st;
                }
            }
            if (it == p.limit) { __syncthreads(); break; }

            // ---- check-node phase (bec.py:100-112)
#pragma unroll
            for (int ps = 0; ps < kResCnPasses; ++ps) {
                if (tid + ps * T < mp) {
                    const int dcr = IRR ? (int)(cw[ps][0] & 15u) : DC;
                    uint4 m[DC];
                    BecCnAccT<uint32_t> a0, a1;
                    a0.init(); a1.init();
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        if (!IRR || k < dcr) {
                            m[k] = *reinterpret_cast<const uint4 *>(smem + cell_off(ps, k));
                            a0.push(m[k].x, m[k].y);
                            a1.push(m[k].z, m[k].w);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        if (!IRR || k < dcr) {
                            uint4 o;
                            a0.out(m[k].x, m[k].y, o.x, o.y);
                            a1.out(m[k].z, m[k].w, o.z, o.w);
                            *reinterpret_cast<uint4 *>(smem + cell_off(ps, k)) = o;
                        }
                    }
                }
            }
            __syncthreads();
            const uint32_t run0 = s_act[0], run1 = s_act[1];
            if ((run0 | run1) == 0u) break;                              // every frame of the tile has stopped

            // ---- variable-node phase (bec.py:115-119)
            uint32_t chg0 = 0u, chg1 = 0u, has0 = 0u, has1 = 0u;
#pragma unroll
            for (int ps = 0; ps < kResVnPasses; ++ps) {
                const int item = tid + ps * T;
                if (item >= np) continue;
                uint32_t mnz0, mpos0, mnz1, mpos1;
                if (!IRR) {
                    uint4 c[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k] = *reinterpret_cast<const uint4 *>(smem + (size_t)(k + 1) * S + (size_t)item * 16);
                    const uint4 pr = prior[item];
                    {
                        const uint32_t nz[4] = {pr.x, c[0].x, c[1].x, c[2].x}, ps4[4] = {pr.y, c[0].y, c[1].y, c[2].y};
                        uint32_t onz[3], opos[3];
                        bec_vn3<uint32_t>(nz, ps4, onz, opos, mnz0, mpos0);
#pragma unroll
                        for (int k = 0; k < 3; ++k) { c[k].x = onz[k]; c[k].y = opos[k]; }
                    }
                    {
                        const uint32_t nz[4] = {pr.z, c[0].z, c[1].z, c[2].z}, ps4[4] = {pr.w, c[0].w, c[1].w, c[2].w};
                        uint32_t onz[3], opos[3];
                        bec_vn3<uint32_t>(nz, ps4, onz, opos, mnz1, mpos1);
#pragma unroll
                        for (int k = 0; k < 3; ++k) { c[k].z = onz[k]; c[k].w = opos[k]; }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) *reinterpret_cast<uint4 *>(smem + (size_t)(k + 1) * S + (size_t)item * 16) = c[k];
                } else {
                    const int d = p.vdeg[item];
                    if (d == 0xff) continue;                             // a position without a variable
                    const uint4 pr = prior[item];
                    BsInt<5, uint32_t> s0, s1;
                    s0.set_ternary(pr.x, pr.y);
                    s1.set_ternary(pr.z, pr.w);
                    uint4 c[DV];
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        if (k < d) {
                            c[k] = *reinterpret_cast<const uint4 *>(smem + p.pbase[k] + (size_t)item * 16);
                            s0.add_ternary(c[k].x, c[k].y);
                            s1.add_ternary(c[k].z, c[k].w);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        if (k < d) {
                            BsInt<5, uint32_t> t0 = s0, t1 = s1;
                            t0.sub_ternary(c[k].x, c[k].y);
                            t1.sub_ternary(c[k].z, c[k].w);
                            uint4 o;
                            t0.sign(o.x, o.y);
                            t1.sign(o.z, o.w);
                            *reinterpret_cast<uint4 *>(smem + p.pbase[k] + (size_t)item * 16) = o;
                        }
                    }
                    s0.sign(mnz0, mpos0);
                    s1.sign(mnz1, mpos1);
                }
                // x_new = symbols[sign(marginal)] (bec.py:119), merged under the run mask; changed / has-erasures flags
                const uint4 xo = xc[item];
                const uint32_t xe0 = ~mnz0, xe1 = ~mnz1;
                chg0 |= ((xe0 ^ xo.x) | (~xe0 & (mpos0 ^ xo.y))) & run0;
                chg1 |= ((xe1 ^ xo.z) | (~xe1 & (mpos1 ^ xo.w))) & run1;
                has0 |= xe0 & run0;
                has1 |= xe1 & run1;
                xc[item] = make_uint4((xo.x & ~run0) | (xe0 & run0), (xo.y & ~run0) | (mpos0 & run0 & ~xe0),
                                      (xo.z & ~run1) | (xe1 & run1), (xo.w & ~run1) | (mpos1 & run1 & ~xe1));
            }
            chg0 = __reduce_or_sync(kFull, chg0); chg1 = __reduce_or_sync(kFull, chg1);
            has0 = __reduce_or_sync(kFull, has0); has1 = __reduce_or_sync(kFull, has1);
            if (lane == 0) {
                if (chg0) atomicOr(&s_chg[0], chg0);
                if (chg1) atomicOr(&s_chg[1], chg1);
                if (has0) atomicOr(&s_has[0], has0);
                if (has1) atomicOr(&s_has[1], has1);
            }
            __syncthreads();
            ++it;
        }

        // ================================ words out: bit planes -> rows ================================
#pragma unroll 1
        for (int w = 0; w < 2; ++w) {
            if (nvalid[w] == 0) break;
            if (w) __syncthreads();                                      // the copy of word 0 has left the staging area
            for (int q = warp; q < n4; q += nwarps) {
                uint32_t packed = 0u;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t pos = s_vpos[4 * q + b];
                    const uint2 xw = reinterpret_cast<const uint2 *>(xc + pos)[w];
                    const uint32_t sy = ((xw.x >> lane) & 1u) ? 2u : ((xw.y >> lane) & 1u);
                    packed |= sy << (8 * b);
                }
                reinterpret_cast<uint32_t *>(stage)[(size_t)lane * n4 + q] = packed;
            }
            __syncthreads();
            bec_copy(p.x_hat + (size_t)(f0 + 32 * w) * n, stage, (size_t)nvalid[w] * n, tid, T);
            if (warp == w && lane < nvalid[w]) {
                const long long f = f0 + 32 * w + lane;
                p.iters[f] = s_iters[w * 32 + lane];
                if (p.reason != nullptr) {
                    uint8_t r = LDPC_REASON_DECODED;
                    if ((s_stop[w] >> lane) & 1u) r = LDPC_REASON_STOPPING;
                    else if ((s_act[w] >> lane) & 1u) r = (uint8_t)p.bound_reason;
                    p.reason[f] = r;
                }
            }
        }
    }
}

}  // namespace ldpc
