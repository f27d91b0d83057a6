// common.cuh — handle, launch bookkeeping, 128-bit frame packs, flag-word helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/ldpc_b200.h"
#include "ldpc_math.cuh"

namespace ldpc {

constexpr int kCtaWarps = 8;                 // CN / VN sweeps: 8 warps = 8 frame groups per CTA
constexpr int kCtaThreads = kCtaWarps * 32;
constexpr unsigned kFull = 0xffffffffu;

enum { ALGO_MSA = 0, ALGO_SPA_REF = 1, ALGO_SPA_PHI = 2 };      // check-node arithmetic variants (ldpc_math.cuh)

// Device copies of the graph tables (edge order = np.where(H), /root/reference/src/bpa.py:12).
struct Tables {
    int n = 0, m = 0, E = 0;
    int *chk_ptr = nullptr, *edge_var = nullptr, *var_ptr = nullptr, *var_edges = nullptr;
    int max_dc = 0, max_dv = 0;
    int uni_dc = 0, uni_dv = 0;              // > 0: every check / variable has exactly this degree
};

}  // namespace ldpc

struct HostStage;                            // ldpc_decode_host staging (api file)

struct ResidentInfo {                        // position-indexed tables of the on-chip path (resident_bp.cuh, res_layout.h)
    bool ok = false;
    bool regular36 = false;                  // every check has 6 edges, every variable 3, no holes: the register-resident variant
    int Q = 1;                               // quads per CTA (4 slots, two CTAs per SM)
    int np = 0, mp = 0;                      // variable / check positions
    int planes = 0, threads = 0;
    // [0]: edge planes ordered for conflict-free gathers (min-sum, symmetric check rule); [1]: natural order (sum-product)
    uint16_t *cvar[2] = {nullptr, nullptr}, *vrow[2] = {nullptr, nullptr};
    uint8_t *cdeg = nullptr, *vdeg = nullptr;
    uint16_t *vposmap = nullptr, *vinvmap = nullptr;
    long plan[7] = {0, 0, 0, 0, 0, 0, 0};    // predicted wavefronts: cn ideal/file/plan-natural/plan, vn ideal/file/plan
    // variable-plane variant (resident_vp.cuh, regular codes): one placement per edge order, [0] min-sum, [1] natural
    bool vp = false;
    bool vp_big = false;                     // ... in the one-CTA-per-SM geometry (resident_vp MAXT = 672; codes up to n ≈ 2850)
    uint16_t *vp_cw[2] = {nullptr, nullptr};                    // [mp][8]: (variable position << 4) | (edge rank at the variable + 1)
    uint16_t *vp_vposmap[2] = {nullptr, nullptr}, *vp_vinvmap[2] = {nullptr, nullptr};
    // ... for irregular codes (IRR = true): one word per edge, planes are prefixes of the positions (res_layout.h, VxTables)
    bool vx = false;
    uint32_t *vx_cwx[2] = {nullptr, nullptr};
    int vx_pcnt[2][8] = {}, vx_pbase[2][8] = {}, vx_cells[2] = {0, 0};
    uint8_t *vx_vdeg = nullptr;              // [np] degree of the variable at a position of placement [0], 0xff = hole (resident_bec.cuh)
};

struct ProfEvent {                           // one timed launch (ldpc_profile_*)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int kind = 0;                            // 0 = check-node sweep, 1 = variable-node sweep
};

struct ldpc_handle {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0, smem_per_sm = 0;
    ldpc::Tables t;
    std::string err;
    unsigned long long launches = 0;
    HostStage *stage = nullptr;
    ResidentInfo res;
    double plan_effort = 0.4;                // annealing effort of the shared-memory placement (LDPC_PLAN_EFFORT)
    bool prof = false;
    std::vector<ProfEvent> prof_ev;
    size_t prof_used = 0;
};

namespace ldpc {

// ---- 16-byte pack of FPT consecutive frames (4 x f32 or 2 x f64): one 128-bit load / store per edge row ----
template <typename T, int FPT> struct alignas(16) Pack {
    T x[FPT];
};

template <typename P> __device__ __forceinline__ P ld_stream(const void *ptr)       // read once: evict-first
{
    static_assert(sizeof(P) == 16, "pack must be 128 bits");
    const uint4 r = __ldcs(reinterpret_cast<const uint4 *>(ptr));
    P out;
    *reinterpret_cast<uint4 *>(&out) = r;
    return out;
}
template <typename P> __device__ __forceinline__ P ld_ro(const void *ptr)           // read-only path (priors)
{
    static_assert(sizeof(P) == 16, "pack must be 128 bits");
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(ptr));
    P out;
    *reinterpret_cast<uint4 *>(&out) = r;
    return out;
}
template <typename P> __device__ __forceinline__ void st_stream(void *ptr, const P &v)
{
    static_assert(sizeof(P) == 16, "pack must be 128 bits");
    __stcs(reinterpret_cast<uint4 *>(ptr), *reinterpret_cast<const uint4 *>(&v));
}

// Frame flags are bit-packed naturally: bit l of word w <-> frame 32*w + l.  A thread that owns FPT
// consecutive frames owns an FPT-bit field of one word; 32/FPT adjacent lanes share that word.
template <int FPT> struct Field {
    static constexpr int LPW = 32 / FPT;                 // lanes per word
    static constexpr uint32_t MASK = (1u << FPT) - 1u;
    static __device__ __forceinline__ int word_of(int group, int lane) { return group * FPT + lane / LPW; }
    static __device__ __forceinline__ int shift_of(int lane) { return (lane % LPW) * FPT; }
    static __device__ __forceinline__ uint32_t get(uint32_t word, int lane) { return (word >> shift_of(lane)) & MASK; }
    // OR the fields of the LPW lanes that share a word; every one of them gets the assembled word.
    static __device__ __forceinline__ uint32_t assemble(uint32_t field, int lane)
    {
        uint32_t w = field << shift_of(lane);
#pragma unroll
        for (int d = 1; d < LPW; d <<= 1) w |= __shfl_xor_sync(kFull, w, d);
        return w;
    }
    static __device__ __forceinline__ bool leader(int lane) { return (lane % LPW) == 0; }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *p) : base(static_cast<char *>(p)) {}
    template <typename U> U *take(size_t count)
    {
        off = align_up(off, 256);
        U *r = reinterpret_cast<U *>(base + off);
        off += count * sizeof(U);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

}  // namespace ldpc
