// stream_bp_tma.cuh — check-node sweep with bulk-async (TMA) row staging.
//
// Why: the register version of cn_sweep (stream_bp.cuh) is latency-bound, not bandwidth-bound — ncu shows
// 80 % long-scoreboard stalls, 45 % occupancy and 4.8 TB/s: the bytes in flight are capped by the registers
// that receive them, and while a warp computes nothing of its own is in flight (profiles/README.md).
// Here the rows of a check land in shared memory through the bulk-copy engine
// (cp.async.bulk ... mbarrier::complete_tx::bytes), S stages deep, so ~(S-1) tiles per CTA are always in
// flight regardless of what the threads are doing; results go back with cp.async.bulk (shared -> global).
//
// Tile = (one check, TF = 128*FPT consecutive frames): dc rows of 2 KB.  CTA = 128 threads, thread = FPT
// frames of the tile (one LDS.128 / STS.128 per row, conflict-free).  A CTA owns one frame tile and walks a
// chunk of checks; thread 0 is the producer: it arms the stage's mbarrier with the byte count, issues the dc
// row copies, and after the compute of a tile issues the dc row stores and refills the stage that was stored
// one tile earlier once the copy engine has finished READING it (cp.async.bulk.wait_group.read).
// Semantics, flags and arithmetic are exactly those of cn_sweep.
#pragma once
#include "stream_bp.cuh"

namespace ldpc {

constexpr int kTmaThreads = 128;
constexpr int kTmaStages = 4;
constexpr int kTmaRowBytes = 2048;           // 128 threads x 16 B

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <typename T, int FPT, int ALGO, int DCMAX, bool UNIFORM>
__global__ void __launch_bounds__(kTmaThreads) cn_sweep_tma(const BpParams<T> p)
{
    using P = Pack<T, FPT>;
    constexpr int TF = kTmaThreads * FPT;                       // frames per tile
    constexpr int WPT = TF / 32;                                // flag words per tile
    constexpr int STAGE_BYTES = DCMAX * kTmaRowBytes;
    static_assert(sizeof(T) * FPT == 16 && TF * sizeof(T) == kTmaRowBytes, "row = 128 threads x 16 bytes");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *stage_base = smem_raw;                                        // [S][DCMAX][2048]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kTmaStages * STAGE_BYTES);
    __shared__ uint32_t s_any;

    const int tid = threadIdx.x, lane = tid & 31;
    const int ftile = blockIdx.x;
    const size_t f0 = (size_t)ftile * TF;
    const int word0 = ftile * WPT;

    // ---- is anything in this frame tile still running?  (act does not change during the sweep)
    if (tid == 0) s_any = 0u;
    __syncthreads();
    if (tid < WPT && p.act[word0 + tid] != 0u) s_any = 1u;
    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1u);
        fence_mbar_init();
    }
    __syncthreads();
    if (s_any == 0u) return;

    const int c_begin = blockIdx.y * p.per_cta;
    const int c_end = min(p.m, c_begin + p.per_cta);
    const int num = c_end - c_begin;
    const bool need_var = p.first || !p.skip_syn;

    auto issue_load = [&](int j) {                                               // thread 0 only
        const int c = c_begin + j;
        int e0, dc;
        if (UNIFORM) { e0 = c * DCMAX; dc = DCMAX; }
        else { e0 = __ldg(p.chk_ptr + c); dc = __ldg(p.chk_ptr + c + 1) - e0; }
        const int s = j % kTmaStages;
        unsigned char *dst = stage_base + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full[s], (uint32_t)dc * kTmaRowBytes);
        for (int k = 0; k < dc; ++k) {
            const T *src = p.first ? p.prior + (size_t)__ldg(p.edge_var + e0 + k) * p.Bp + f0
                                   : p.msg + (size_t)(e0 + k) * p.Bp + f0;
            bulk_g2s(dst + (size_t)k * kTmaRowBytes, src, kTmaRowBytes, &full[s]);
        }
    };

    if (tid == 0) {
        const int pre = min(kTmaStages - 1, num);
        for (int j = 0; j < pre; ++j) issue_load(j);
    }

    // flag word / field of this thread's FPT frames
    const int word = word0 + tid / (32 / FPT);
    const int shift = (tid % (32 / FPT)) * FPT;
    constexpr uint32_t MASK = (1u << FPT) - 1u;
    uint32_t unsat = 0u;

    for (int i = 0; i < num; ++i) {
        const int c = c_begin + i;
        int e0, dc;
        if (UNIFORM) { e0 = c * DCMAX; dc = DCMAX; }
        else { e0 = __ldg(p.chk_ptr + c); dc = __ldg(p.chk_ptr + c + 1) - e0; }
        const int s = i % kTmaStages;
        unsigned char *st = stage_base + (size_t)s * STAGE_BYTES;

        // syndrome words can be fetched while the rows are still landing
        uint32_t syn = 0u;
        if (!p.skip_syn) {
#pragma unroll
            for (int k = 0; k < DCMAX; ++k)
                if (UNIFORM || k < dc) syn ^= __ldg(p.xbits + (size_t)__ldg(p.edge_var + e0 + k) * p.wpr + word);
        }
        (void)need_var;

        mbar_wait(&full[s], (uint32_t)((i / kTmaStages) & 1));

        P v[DCMAX];
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (UNIFORM || k < dc) v[k] = *reinterpret_cast<const P *>(st + (size_t)k * kTmaRowBytes + (size_t)tid * 16);
        unsat |= (syn >> shift) & MASK;
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
            if (UNIFORM || k < dc) *reinterpret_cast<P *>(st + (size_t)k * kTmaRowBytes + (size_t)tid * 16) = v[k];

        fence_async_smem();                                   // generic-proxy writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            for (int k = 0; k < dc; ++k)
                bulk_s2g(p.msg + (size_t)(e0 + k) * p.Bp + f0, st + (size_t)k * kTmaRowBytes, kTmaRowBytes);
            bulk_commit();
            const int j = i + kTmaStages - 1;                 // refills the stage stored one tile ago
            if (j < num) {
                bulk_wait_read<1>();                          // every store group but the newest has left shared memory
                issue_load(j);
            }
        }
    }
    if (tid == 0) bulk_wait_all<0>();                         // stores are complete before the CTA retires

    if (!p.skip_syn) {
        const uint32_t w = Field<FPT>::assemble(unsat, lane);
        if (Field<FPT>::leader(lane) && w != 0u) atomicOr(p.unsat + word, w);
    }
}

template <int DCMAX> constexpr size_t tma_smem_bytes() { return (size_t)kTmaStages * DCMAX * kTmaRowBytes + kTmaStages * sizeof(uint64_t) + 64; }

}  // namespace ldpc
