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
// inside the kernel: a warp takes 32 variables x 32 frames, every lane reads the 32 bytes of its own row straight from
// global memory (whole 32-byte sectors: no staging, no barrier), packs them into a value and an erasure word, and five
// shuffle stages transpose the 32 x 32 bit tile (BecTr); the words leave the same way in reverse.
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

// 32 x 32 bit transpose across a warp: lane i holds row i in, column i out (bit j of the result = bit i of lane j's
// input).  Five butterfly stages; in each the lane ROTATES its word so that the half its partner wants sits where the
// partner keeps it (rotate amount and keep-mask depend on the lane only: loop invariants), one shuffle, one bit-select
// — against 64 ballots plus 64 compares for the same tile, which made the symbol transposes 46 % of the kernel's
// instructions (profiles/, r2a).
struct BecTr {
    uint32_t amt[5], keep[5];
    __device__ __forceinline__ explicit BecTr(int lane)
    {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int w = 16 >> s;
            const uint32_t m = (s == 0) ? 0x0000ffffu : (s == 1) ? 0x00ff00ffu : (s == 2) ? 0x0f0f0f0fu : (s == 3) ? 0x33333333u : 0x55555555u;
            const bool up = (lane & w) != 0;
            amt[s] = up ? (uint32_t)w : (uint32_t)(32 - w);          // upper lane: its low half moves up; lower lane: high half down
            keep[s] = up ? ~m : m;
            // opaque to the optimiser: otherwise it re-derives both per use from (constant ^ lane bit), three LOP3 per stage
            asm("" : "+r"(amt[s]), "+r"(keep[s]));
        }
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const
    {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const uint32_t y = __shfl_xor_sync(kFull, __funnelshift_l(x, x, amt[s]), 16 >> s);
            x = (x & keep[s]) | (y & ~keep[s]);
        }
        return x;
    }
};

// Symbol bytes {0, 1, >= 2 = erased} <-> bit words.  Eight 32-bit loads hold symbols 4 i + b (word i, byte b) of a
// 32-symbol block; the words below keep symbol 4 i + b at bit 8 b + i, a fixed permutation of the block that the lane
// <-> variable map of the transposed tile absorbs (bec_tile_var), so packing is one mask + one multiply-add per word
// and plane (the multiplies run on the FMA pipe, which the kernel leaves idle; its limiter is the ALU pipe).
__device__ __forceinline__ int bec_tile_var(int lane) { return 4 * (lane & 7) + (lane >> 3); }
__device__ __forceinline__ void sym32_to_words(const uint32_t (&sy)[8], uint32_t &val, uint32_t &er)
{
    uint32_t v = 0u, e = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t w = sy[i];
        v = (w & 0x01010101u) * (1u << i) + v;                                  // bit 0 of byte b -> bit 8 b + i (disjoint: + is |)
        const uint32_t f = (((w & 0x7e7e7e7eu) + 0x7e7e7e7eu) | w) & 0x80808080u;   // bit 7 of byte b: byte >= 2
        e = (i == 7) ? (e + f) : (__umulhi(f, 1u << (25 + i)) + e);             // >> (7 - i): bit 8 b + 7 -> bit 8 b + i
    }
    val = v & ~e;
    er = e;
}
// ... and back (value and erased are disjoint): word i of the block
__device__ __forceinline__ uint32_t words_to_sym4(uint32_t val, uint32_t er, int i)
{
    const uint32_t v = (val >> i) & 0x01010101u;
    const uint32_t e = (i == 0) ? ((er & 0x01010101u) << 1) : ((er >> (i - 1)) & 0x02020202u);
    return v | e;
}

// 32 consecutive symbols (variables 32 g ..) of one frame's row, straight from / to global memory.  A lane touches whole
// 32-byte sectors, so the warp's 32 rows cost exactly the bytes they hold; no staging, no barrier.  Rows are 16-byte
// aligned when n % 16 == 0 and the block itself is (vec16), 4-byte aligned when n % 4 == 0 (vec4), else bytes.
__device__ __forceinline__ void bec_load32(const uint8_t *row, int g, int n, bool vec16, bool vec4, uint32_t (&sy)[8])
{
    const uint8_t *q = row + (size_t)g * 32;
    if (vec16 && g * 32 + 32 <= n) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(q)), b = __ldcs(reinterpret_cast<const uint4 *>(q) + 1);
        sy[0] = a.x; sy[1] = a.y; sy[2] = a.z; sy[3] = a.w; sy[4] = b.x; sy[5] = b.y; sy[6] = b.z; sy[7] = b.w;
    } else if (vec4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sy[i] = (g * 32 + 4 * i < n) ? __ldcs(reinterpret_cast<const uint32_t *>(q) + i) : 0u;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t w = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (g * 32 + 4 * i + b < n) w |= (uint32_t)q[4 * i + b] << (8 * b);
            sy[i] = w;
        }
    }
}
__device__ __forceinline__ void bec_store32(uint8_t *row, int g, int n, bool vec16, bool vec4, const uint32_t (&sy)[8])
{
    uint8_t *q = row + (size_t)g * 32;
    if (vec16 && g * 32 + 32 <= n) {
        __stcs(reinterpret_cast<uint4 *>(q), make_uint4(sy[0], sy[1], sy[2], sy[3]));
        __stcs(reinterpret_cast<uint4 *>(q) + 1, make_uint4(sy[4], sy[5], sy[6], sy[7]));
    } else if (vec4) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (g * 32 + 4 * i < n) __stcs(reinterpret_cast<uint32_t *>(q) + i, sy[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (g * 32 + 4 * i + b < n) q[4 * i + b] = (uint8_t)(sy[i] >> (8 * b));
    }
}

// IRR = false: regular code, every check 6 edges, every variable 3 (resident_vp's `cw` tables).
// IRR = true : check degrees 2..6, variable degrees 0..8, holes (resident_vp's irregular `cwx` tables).
// CNP / VNP: check / variable items per thread (mp <= CNP * T, np <= VNP * T); MAXT: threads per CTA, two CTAs per SM.
// (2, 4, 320) is resident_vp's geometry; (1, 2, 608) gives every thread ONE check and TWO variables of an n = 1200 code:
// 19 + 19 warps per SM instead of 10 + 10 behind the same two barriers per iteration, at 48 registers per thread — measured
// SLOWER (318 against 356 M frames/s on config 2: the transposes spill, instructions grow 16 %), so it is only launched with
// LDPC_BEC_WIDE=1 (A/B runs); the default is resident_vp's geometry.
template <bool IRR, int CNP = kResCnPasses, int VNP = kResVnPasses, int MAXT = 320>
__global__ void __launch_bounds__(MAXT, 2) resident_bec(const BecResParams p)
{
    constexpr int DC = 6, DV = IRR ? 8 : 3, CH = IRR ? DC : 3;
    extern __shared__ __align__(128) unsigned char smem[];
    const int np = p.np, mp = p.mp, n = p.nref;
    // alignment of the symbol rows in global memory (both blocks): 128-bit, 32-bit or byte accesses
    const uintptr_t addr_or = reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.x_hat);
    const bool vec16 = ((n | addr_or) & 15) == 0, vec4 = ((n | addr_or) & 3) == 0;
    const uint32_t S = (uint32_t)np * 16u;
    const BecSmem L = bec_smem_layout(np, p.plane_cells, IRR, n);
    uint4 *xc = reinterpret_cast<uint4 *>(smem + L.x);
    uint4 *prior = reinterpret_cast<uint4 *>(smem + L.prior);
    uint16_t *s_vpos = reinterpret_cast<uint16_t *>(smem + L.vpos);

    __shared__ uint32_t s_act[2], s_chg[2], s_has[2], s_stop[2];
    __shared__ int s_iters[64];
    __shared__ int s_tile;

    const int tid = threadIdx.x, T = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int tvar = bec_tile_var(lane);

    // ---- per-thread graph indices -> registers (once per CTA), packed exactly as in resident_vp
    uint32_t cw[CNP][CH];
#pragma unroll
    for (int ps = 0; ps < CNP; ++ps) {
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

        // ================================ symbols in: rows -> bit planes (x cells) ================================
        // warp = 32 consecutive variables x 32 frames: lane = frame packs its 32 symbols into a value word and an
        // erasure word, two bit transposes turn them into the planes, lane = variable stores its half cell
        {
            const BecTr tr(lane);
            uint32_t er0 = 0u, er1 = 0u;
            const uint8_t *row0 = p.y + (size_t)(f0 + lane) * n, *row1 = row0 + (size_t)32 * n;
            for (int g = warp; g * 32 < n; g += nwarps) {
                uint32_t sa[8], sb[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { sa[i] = 0u; sb[i] = 0u; }
                if (lane < nvalid[0]) bec_load32(row0, g, n, vec16, vec4, sa);
                if (lane < nvalid[1]) bec_load32(row1, g, n, vec16, vec4, sb);
                uint32_t V0, E0, V1, E1;
                sym32_to_words(sa, V0, E0);
                sym32_to_words(sb, V1, E1);
                V0 = tr(V0); E0 = tr(E0);
                V1 = tr(V1); E1 = tr(E1);
                er0 |= E0; er1 |= E1;
                const int v = g * 32 + tvar;                                      // the variable this lane holds after the transposes
                if (v < n) xc[s_vpos[v]] = make_uint4(E0, V0, E1, V1);           // (xe, xv) of both words
            }
            er0 = __reduce_or_sync(kFull, er0);
            er1 = __reduce_or_sync(kFull, er1);
            if (lane == 0) {
                if (er0 != 0u) atomicOr(&s_has[0], er0);
                if (er1 != 0u) atomicOr(&s_has[1], er1);
            }
        }
        __syncthreads();
        // priors = messages[y] (bec.py:85); v2c = priors[yy] (bec.py:86): the prior goes into every edge cell of the variable
#pragma unroll
        for (int ps = 0; ps < VNP; ++ps) {
            const int item = tid + ps * T;
            if (item < np) {
                int d = DV;
                if (IRR) d = p.vdeg[item];
                if (!IRR || d != 0xff) {
                    const uint4 x = xc[item];
                    const uint4 pr = make_uint4(okw[0] & ~x.x, x.y, okw[1] & ~x.z, x.w);       // (nz, pos) of both words
                    // regular instance: the prior is kept as (is +1, is -1), the form bec_vn3_pn consumes
                    prior[item] = IRR ? pr : make_uint4(pr.y, pr.x & ~pr.y, pr.w, pr.z & ~pr.w);
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
            for (int ps = 0; ps < CNP; ++ps) {
                if (tid + ps * T < mp) {
                    uint4 m[DC];
                    if (!IRR) {
#pragma unroll
                        for (int k = 0; k < DC; ++k) m[k] = *reinterpret_cast<const uint4 *>(smem + cell_off(ps, k));
                        BecCn6<uint32_t> t0, t1;
                        {
                            const uint32_t nz[6] = {m[0].x, m[1].x, m[2].x, m[3].x, m[4].x, m[5].x};
                            const uint32_t ps6[6] = {m[0].y, m[1].y, m[2].y, m[3].y, m[4].y, m[5].y};
                            t0.reduce(nz, ps6);
                        }
                        {
                            const uint32_t nz[6] = {m[0].z, m[1].z, m[2].z, m[3].z, m[4].z, m[5].z};
                            const uint32_t ps6[6] = {m[0].w, m[1].w, m[2].w, m[3].w, m[4].w, m[5].w};
                            t1.reduce(nz, ps6);
                        }
#pragma unroll
                        for (int k = 0; k < DC; ++k) {
                            uint4 o;
                            t0.out(m[k].x, m[k].y, o.x, o.y);
                            t1.out(m[k].z, m[k].w, o.z, o.w);
                            *reinterpret_cast<uint4 *>(smem + cell_off(ps, k)) = o;
                        }
                    } else {
                        const int dcr = (int)(cw[ps][0] & 15u);
                        BecCnAccT<uint32_t> a0, a1;
                        a0.init(); a1.init();
#pragma unroll
                        for (int k = 0; k < DC; ++k) {
                            if (k < dcr) {
                                m[k] = *reinterpret_cast<const uint4 *>(smem + cell_off(ps, k));
                                a0.push(m[k].x, m[k].y);
                                a1.push(m[k].z, m[k].w);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < DC; ++k) {
                            if (k < dcr) {
                                uint4 o;
                                a0.out(m[k].x, m[k].y, o.x, o.y);
                                a1.out(m[k].z, m[k].w, o.z, o.w);
                                *reinterpret_cast<uint4 *>(smem + cell_off(ps, k)) = o;
                            }
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
            for (int ps = 0; ps < VNP; ++ps) {
                const int item = tid + ps * T;
                if (item >= np) continue;
                uint32_t mnz0, mpos0, mnz1, mpos1;
                if (!IRR) {
                    uint4 c[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k] = *reinterpret_cast<const uint4 *>(smem + (size_t)(k + 1) * S + (size_t)item * 16);
                    const uint4 pr = prior[item];                        // (is +1, is -1) of both words
                    const uint32_t pa[4] = {pr.x, c[0].y, c[1].y, c[2].y}, na[4] = {pr.y, c[0].x & ~c[0].y, c[1].x & ~c[1].y, c[2].x & ~c[2].y};
                    const uint32_t pb[4] = {pr.z, c[0].w, c[1].w, c[2].w}, nb[4] = {pr.w, c[0].z & ~c[0].w, c[1].z & ~c[1].w, c[2].z & ~c[2].w};
                    uint32_t onz0[3], opos0[3], onz1[3], opos1[3];
                    if ((bec_conflict<uint32_t>(pa, na) | bec_conflict<uint32_t>(pb, nb)) == 0u) {
                        // no frame has votes of both signs at this variable (always so for symbols of an erasure channel
                        // on a codeword): the sums' signs are ORs of the other inputs
                        bec_vn3_or<uint32_t>(pa, na, onz0, opos0, mnz0, mpos0);
                        bec_vn3_or<uint32_t>(pb, nb, onz1, opos1, mnz1, mpos1);
                    } else {                                             // inconsistent input: the literal rule
                        bec_vn3_pn<uint32_t>(pa, na, onz0, opos0, mnz0, mpos0);
                        bec_vn3_pn<uint32_t>(pb, nb, onz1, opos1, mnz1, mpos1);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k] = make_uint4(onz0[k], opos0[k], onz1[k], opos1[k]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) *reinterpret_cast<uint4 *>(smem + (size_t)(k + 1) * S + (size_t)item * 16) = c[k];
                } else {
                    const int d = p.vdeg[item];
                    if (d == 0xff) continue;                             // a position without a variable
                    const uint4 pr = prior[item];
                    uint4 c[DV];
                    BecVnOr<uint32_t> g0, g1;
                    g0.init(pr.x, pr.y);
                    g1.init(pr.z, pr.w);
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        if (k < d) {
                            c[k] = *reinterpret_cast<const uint4 *>(smem + p.pbase[k] + (size_t)item * 16);
                            g0.push(c[k].x, c[k].y);
                            g1.push(c[k].z, c[k].w);
                        }
                    }
                    if ((g0.conflict() | g1.conflict()) == 0u) {         // no conflicting votes: ORs of the other inputs
#pragma unroll
                        for (int k = 0; k < DV; ++k) {
                            if (k < d) {
                                uint4 o;
                                g0.out(c[k].x, c[k].y, o.x, o.y);
                                g1.out(c[k].z, c[k].w, o.z, o.w);
                                *reinterpret_cast<uint4 *>(smem + p.pbase[k] + (size_t)item * 16) = o;
                            }
                        }
                        g0.marg(mnz0, mpos0);
                        g1.marg(mnz1, mpos1);
                    } else {                                             // inconsistent input: bit-sliced integers, any degree <= 8
                        BsInt<5, uint32_t> s0, s1;
                        s0.set_ternary(pr.x, pr.y);
                        s1.set_ternary(pr.z, pr.w);
#pragma unroll
                        for (int k = 0; k < DV; ++k) {
                            if (k < d) {
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
                }
                // x_new = symbols[sign(marginal)] (bec.py:119), merged under the run mask; changed / has-erasures flags
                // (xv is a subset of ~xe on both sides, so "changed" is just a difference in either plane)
                const uint4 xo = xc[item];
                chg0 |= ((~mnz0 ^ xo.x) | (mpos0 ^ xo.y)) & run0;
                chg1 |= ((~mnz1 ^ xo.z) | (mpos1 ^ xo.w)) & run1;
                has0 |= ~mnz0 & run0;
                has1 |= ~mnz1 & run1;
                xc[item] = make_uint4((xo.x & ~run0) | (~mnz0 & run0), (xo.y & ~run0) | (mpos0 & run0),
                                      (xo.z & ~run1) | (~mnz1 & run1), (xo.w & ~run1) | (mpos1 & run1));
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
        {
            const BecTr tr(lane);
            uint8_t *row0 = p.x_hat + (size_t)(f0 + lane) * n, *row1 = row0 + (size_t)32 * n;
            for (int g = warp; g * 32 < n; g += nwarps) {
                const int v = g * 32 + tvar;
                uint4 xw = make_uint4(0u, 0u, 0u, 0u);
                if (v < n) xw = xc[s_vpos[v]];
                const uint32_t E0 = tr(xw.x), V0 = tr(xw.y);                      // lane = frame again
                const uint32_t E1 = tr(xw.z), V1 = tr(xw.w);
                uint32_t sy[8];
                if (lane < nvalid[0]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) sy[i] = words_to_sym4(V0, E0, i);
                    bec_store32(row0, g, n, vec16, vec4, sy);
                }
                if (lane < nvalid[1]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) sy[i] = words_to_sym4(V1, E1, i);
                    bec_store32(row1, g, n, vec16, vec4, sy);
                }
            }
            if (warp < 2 && lane < nvalid[warp]) {
                const long long f = f0 + 32 * warp + lane;
                p.iters[f] = s_iters[warp * 32 + lane];
                if (p.reason != nullptr) {
                    uint8_t r = LDPC_REASON_DECODED;
                    if ((s_stop[warp] >> lane) & 1u) r = LDPC_REASON_STOPPING;
                    else if ((s_act[warp] >> lane) & 1u) r = (uint8_t)p.bound_reason;
                    p.reason[f] = r;
                }
            }
        }
    }
}

}  // namespace ldpc
