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

// Four flag / plane words at once: a thread then owns 128 frames and every row access is one 128-bit load or store
// (the single-word kernels keep 4 bytes per thread in flight per access and reach 40-45 % of the DRAM rate,
// profiles/README.md).  Needs wpr % 4 == 0, which bec_layout guarantees.
struct W4 {
    uint32_t x, y, z, w;
    __host__ __device__ W4() : x(0u), y(0u), z(0u), w(0u) {}
    __host__ __device__ W4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) : x(a), y(b), z(c), w(d) {}
};
__device__ __forceinline__ W4 operator&(W4 a, W4 b) { return W4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ W4 operator|(W4 a, W4 b) { return W4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ W4 operator^(W4 a, W4 b) { return W4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
__device__ __forceinline__ W4 operator~(W4 a) { return W4(~a.x, ~a.y, ~a.z, ~a.w); }
__device__ __forceinline__ bool nonzero(W4 a) { return (a.x | a.y | a.z | a.w) != 0u; }
__device__ __forceinline__ bool nonzero(uint32_t a) { return a != 0u; }

template <typename U> struct WordIO;
template <> struct WordIO<uint32_t> {
    static constexpr int N = 1;
    static __device__ __forceinline__ uint32_t ld(const uint32_t *p) { return *p; }
    static __device__ __forceinline__ uint32_t ldro(const uint32_t *p) { return __ldg(p); }
    static __device__ __forceinline__ void st(uint32_t *p, uint32_t v) { *p = v; }
    static __device__ __forceinline__ void atomic_or(uint32_t *p, uint32_t v) { if (v != 0u) atomicOr(p, v); }
};
template <> struct WordIO<W4> {
    static constexpr int N = 4;
    static __device__ __forceinline__ W4 ld(const uint32_t *p) { const uint4 v = *reinterpret_cast<const uint4 *>(p); return W4(v.x, v.y, v.z, v.w); }
    static __device__ __forceinline__ W4 ldro(const uint32_t *p) { const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p)); return W4(v.x, v.y, v.z, v.w); }
    static __device__ __forceinline__ void st(uint32_t *p, W4 v) { *reinterpret_cast<uint4 *>(p) = make_uint4(v.x, v.y, v.z, v.w); }
    static __device__ __forceinline__ void atomic_or(uint32_t *p, W4 v)
    {
        if (v.x != 0u) atomicOr(p, v.x);
        if (v.y != 0u) atomicOr(p + 1, v.y);
        if (v.z != 0u) atomicOr(p + 2, v.z);
        if (v.w != 0u) atomicOr(p + 3, v.w);
    }
};

template <typename U>
__global__ void __launch_bounds__(128) bec_cn(const BecParams p)
{
    using IO = WordIO<U>;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) * IO::N;
    if (w >= p.wpr) return;
    if (!nonzero(IO::ld(p.act + w))) return;
    const int c_begin = blockIdx.y * p.per_cta;
    const int c_end = min(p.m, c_begin + p.per_cta);
    for (int c = c_begin; c < c_end; ++c) {
        const int e0 = __ldg(p.chk_ptr + c), e1 = __ldg(p.chk_ptr + c + 1);
        BecCnAccT<U> acc;
        acc.init();
        for (int e = e0; e < e1; ++e) {
            U nz, pos;
            if (p.first) {
                const size_t r = (size_t)__ldg(p.edge_var + e) * p.wpr + w;
                nz = IO::ldro(p.pnz + r); pos = IO::ldro(p.ppos + r);
            } else {
                const size_t r = (size_t)e * p.wpr + w;
                nz = IO::ld(p.mnz + r); pos = IO::ld(p.mpos + r);
            }
            acc.push(nz, pos);
        }
        for (int e = e0; e < e1; ++e) {
            U nz, pos, onz, opos;
            const size_t r = (size_t)e * p.wpr + w;
            if (p.first) {
                const size_t rp = (size_t)__ldg(p.edge_var + e) * p.wpr + w;
                nz = IO::ldro(p.pnz + rp); pos = IO::ldro(p.ppos + rp);
            } else {
                nz = IO::ld(p.mnz + r); pos = IO::ld(p.mpos + r);
            }
            acc.out(nz, pos, onz, opos);
            IO::st(p.mnz + r, onz); IO::st(p.mpos + r, opos);
        }
    }
}

template <int NB, typename U>
__global__ void __launch_bounds__(128) bec_vn(const BecParams p)
{
    using IO = WordIO<U>;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) * IO::N;
    if (w >= p.wpr) return;
    const U run = IO::ld(p.act + w);
    if (!nonzero(run)) return;
    const int v_begin = blockIdx.y * p.per_cta;
    const int v_end = min(p.n, v_begin + p.per_cta);
    U changed = U(), haser = U();
    for (int v = v_begin; v < v_end; ++v) {
        const int p0 = __ldg(p.var_ptr + v), p1 = __ldg(p.var_ptr + v + 1);
        const size_t rv = (size_t)v * p.wpr + w;
        BsInt<NB, U> acc;
        acc.set_ternary(IO::ldro(p.pnz + rv), IO::ldro(p.ppos + rv));    // marginal = priors + sum_cols(c2v), bec.py:115
        for (int k = p0; k < p1; ++k) {
            const size_t r = (size_t)__ldg(p.var_edges + k) * p.wpr + w;
            acc.add_ternary(IO::ld(p.mnz + r), IO::ld(p.mpos + r));
        }
        for (int k = p0; k < p1; ++k) {                                  // v2c = sign(marginal[yy] - c2v), bec.py:116
            const size_t r = (size_t)__ldg(p.var_edges + k) * p.wpr + w;
            BsInt<NB, U> t = acc;
            t.sub_ternary(IO::ld(p.mnz + r), IO::ld(p.mpos + r));
            U nz, pos;
            t.sign(nz, pos);
            IO::st(p.mnz + r, nz); IO::st(p.mpos + r, pos);
        }
        U nz, pos;
        acc.sign(nz, pos);                                               // x_new = symbols[sign(marginal)], bec.py:119
        const U xe_new = ~nz, xv_new = pos;
        const U xe_old = IO::ld(p.xe + rv), xv_old = IO::ld(p.xv + rv);
        changed = changed | (((xe_new ^ xe_old) | (~xe_new & (xv_new ^ xv_old))) & run);
        haser = haser | (xe_new & run);
        IO::st(p.xe + rv, (xe_old & ~run) | (xe_new & run));             // if nothing changed this is a no-op (bec.py:120-121)
        IO::st(p.xv + rv, (xv_old & ~run) | (xv_new & run & ~xe_new));
    }
    IO::atomic_or(p.changed + w, changed);
    IO::atomic_or(p.haser + w, haser);
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
