// stream_bec.cuh — BEC erasure message passing (bec.SPA.decode, /root/reference/src/bec.py:83-122)
// on bit planes: a message in {-1,0,+1} is (nz, pos), 32 frames per 32-bit word.
//
// HBM layout (wpr = Bp/32 words per row):
//   mnz, mpos [E][wpr]   the message planes, updated in place (c2v after CN, v2c after VN)
//   pnz, ppos [n][wpr]   priors = messages[y]   (bec.py:76,85): y=0 -> -1, y=1 -> +1, y=2 -> 0
//   xe,  xv   [n][wpr]   x_hat: erased plane, value plane (symbols 0/1/2)
//   act, changed, haser, stopped [wpr]   per-frame flags;  iters [Bp]
//
// One round = bec_book -> bec_cn -> bec_vn:
//   bec_book  (after the previous VN) frames whose word did not change stop ("stopping", bec.py:120,
//             iter_count not incremented); the others count the round (bec.py:122); then frames without
//             erasures stop ("decoded", bec.py:97).
//   bec_cn    per check: erasure count saturating at 2 and parity of +1 votes -> c2v (bec.py:100-112)
//   bec_vn    marg = prior + sum c2v as a bit-sliced integer; v2c = sign(marg - c2v); x_new = symbols[sign(marg)]
// Thread = (check or variable, one word = 32 frames); no per-degree templates: both sweeps make two
// passes over the rows (the second pass hits L1), so any degree works.
#pragma once
#include "common.cuh"

namespace ldpc {

struct BecParams {
    int n, m, E, wpr;
    const int *chk_ptr, *edge_var, *var_ptr, *var_edges;
    uint32_t *mnz, *mpos;
    const uint32_t *pnz, *ppos;
    uint32_t *xe, *xv;
    uint32_t *act, *changed, *haser, *stopped;
    int *iters;
    int per_cta;
    int first;      // CN: v2c = priors[yy] (bec.py:86)
};

__global__ void __launch_bounds__(128) bec_cn(const BecParams p)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= p.wpr) return;
    if (p.act[w] == 0u) return;
    const int c_begin = blockIdx.y * p.per_cta;
    const int c_end = min(p.m, c_begin + p.per_cta);
    for (int c = c_begin; c < c_end; ++c) {
        const int e0 = __ldg(p.chk_ptr + c), e1 = __ldg(p.chk_ptr + c + 1);
        BecCnAcc acc;
        acc.init();
        for (int e = e0; e < e1; ++e) {
            uint32_t nz, pos;
            if (p.first) {
                const size_t r = (size_t)__ldg(p.edge_var + e) * p.wpr + w;
                nz = __ldg(p.pnz + r); pos = __ldg(p.ppos + r);
            } else {
                const size_t r = (size_t)e * p.wpr + w;
                nz = p.mnz[r]; pos = p.mpos[r];
            }
            acc.push(nz, pos);
        }
        for (int e = e0; e < e1; ++e) {
            uint32_t nz, pos, onz, opos;
            const size_t r = (size_t)e * p.wpr + w;
            if (p.first) {
                const size_t rp = (size_t)__ldg(p.edge_var + e) * p.wpr + w;
                nz = __ldg(p.pnz + rp); pos = __ldg(p.ppos + rp);
            } else {
                nz = p.mnz[r]; pos = p.mpos[r];
            }
            acc.out(nz, pos, onz, opos);
            p.mnz[r] = onz; p.mpos[r] = opos;
        }
    }
}

template <int NB>
__global__ void __launch_bounds__(128) bec_vn(const BecParams p)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= p.wpr) return;
    const uint32_t run = p.act[w];
    if (run == 0u) return;
    const int v_begin = blockIdx.y * p.per_cta;
    const int v_end = min(p.n, v_begin + p.per_cta);
    uint32_t changed = 0u, haser = 0u;
    for (int v = v_begin; v < v_end; ++v) {
        const int p0 = __ldg(p.var_ptr + v), p1 = __ldg(p.var_ptr + v + 1);
        const size_t rv = (size_t)v * p.wpr + w;
        BsInt<NB> acc;
        acc.set_ternary(__ldg(p.pnz + rv), __ldg(p.ppos + rv));          // marginal = priors + sum_cols(c2v), bec.py:115
        for (int k = p0; k < p1; ++k) {
            const size_t r = (size_t)__ldg(p.var_edges + k) * p.wpr + w;
            acc.add_ternary(p.mnz[r], p.mpos[r]);
        }
        for (int k = p0; k < p1; ++k) {                                  // v2c = sign(marginal[yy] - c2v), bec.py:116
            const size_t r = (size_t)__ldg(p.var_edges + k) * p.wpr + w;
            BsInt<NB> t = acc;
            t.sub_ternary(p.mnz[r], p.mpos[r]);
            uint32_t nz, pos;
            t.sign(nz, pos);
            p.mnz[r] = nz; p.mpos[r] = pos;
        }
        uint32_t nz, pos;
        acc.sign(nz, pos);                                               // x_new = symbols[sign(marginal)], bec.py:119
        const uint32_t xe_new = ~nz, xv_new = pos;
        const uint32_t xe_old = p.xe[rv], xv_old = p.xv[rv];
        changed |= ((xe_new ^ xe_old) | (~xe_new & (xv_new ^ xv_old))) & run;
        haser |= xe_new & run;
        p.xe[rv] = (xe_old & ~run) | (xe_new & run);                     // if nothing changed this is a no-op (bec.py:120-121)
        p.xv[rv] = (xv_old & ~run) | (xv_new & run & ~xe_new);
    }
    if (changed != 0u) atomicOr(p.changed + w, changed);
    if (haser != 0u) atomicOr(p.haser + w, haser);
}

// One warp per flag word, lane = frame.  first: nothing ran yet (only the "no erasures" test applies);
// last: the loop bound was reached (the reference tests max_iter BEFORE erasures, bec.py:96-97).
__global__ void bec_book(uint32_t *act, uint32_t *changed, uint32_t *haser, uint32_t *stopped, int *iters,
                         int wpr, int first, int last, int *any_active)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= wpr) return;
    uint32_t a = act[w];
    const uint32_t ch = changed[w], h = haser[w];
    uint32_t st = 0u;
    if (!first) {
        st = a & ~ch;                                                    // x_new == x_hat: 'stopping', iter_count stays
        a &= ch;
        if ((a >> lane) & 1u) iters[(size_t)w * 32 + lane] += 1;         // bec.py:122
    }
    if (!last) a &= h;                                                   // no erasures left: 'decoded' (bec.py:97)
    __syncwarp();
    if (lane == 0) {
        act[w] = a;
        changed[w] = 0u;
        if (!last) haser[w] = 0u;
        if (st != 0u) stopped[w] |= st;
        if (a != 0u) *any_active = 1;
    }
}

}  // namespace ldpc
