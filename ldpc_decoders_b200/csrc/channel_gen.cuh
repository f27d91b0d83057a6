// channel_gen.cuh — on-device channel simulators: the reference's Channel.send of
//   bec.py:15-18   y = clip(x + 10 * (u < p), 0, 2)          -> 2 where erased, else x
//   bsc.py:15-16   y = (x + (u < p)) % 2
//   biawgn.py:17-18  y = (2x - 1) + normal(0, sqrt(noise_var))   (float32 here)
// with counter-based Philox4x32-10 noise instead of numpy's global Mersenne twister: statistically, not bit-,
// equivalent (parity runs keep uploading numpy draws).  The stream of a received value depends only on
// (seed, global frame index, variable index), so a Monte-Carlo run gives the same frames however it is cut
// into batches or spread over GPUs.
#pragma once
#if defined(__CUDACC__)
#include "common.cuh"
#else
#include "ldpc_math.cuh"      // host build (tests/host_emu): the generators only
#endif

namespace ldpc {

struct Philox4 {
    uint32_t x, y, z, w;
};

LDPC_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1).
LDPC_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// The four 32-bit words of (seed, frame, group of 4 variables).
LDPC_HD Philox4 channel_words(unsigned long long seed, unsigned long long frame, uint32_t group)
{
    return philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), group, 0x4c445043u /* "LDPC" */,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Box-Muller on two words: u1 in (0, 1) with 32 bits of resolution (|z| up to 6.7 sigma), u2 in [0, 1).
// On the device the logarithm and the sine / cosine are the SFU approximations (__logf: 1 ulp-of-2^-21.4 absolute on the
// result's scale, __sincosf on an angle folded into [-pi, pi): 2^-21.4 absolute): the noise changes by ~1e-6 sigma, far
// below anything a Monte-Carlo estimate resolves, and the generator kernel gets ~40 % shorter.  The angle is drawn from
// [-pi, pi) instead of [0, 2 pi): the same uniform distribution on the circle.
LDPC_HD void box_muller(uint32_t r1, uint32_t r2, float *z0, float *z1)
{
    const float u1 = fmaf((float)r1, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float u2 = (float)r2 * 2.3283064365386963e-10f;
    float s, c;
#if defined(__CUDA_ARCH__)
    const float rad = sqrtf(-2.0f * __logf(u1));
    __sincosf(6.283185307179586f * (u2 - 0.5f), &s, &c);
#else
    const float rad = sqrtf(-2.0f * logf(u1));
    s = sinf(6.283185307179586f * (u2 - 0.5f)); c = cosf(6.283185307179586f * (u2 - 0.5f));
#endif
    *z0 = rad * c;
    *z1 = rad * s;
}

enum { GEN_BSC = 1, GEN_BIAWGN = 2, GEN_BEC = 3 };     // == LDPC_CH_*

#if defined(__CUDACC__)
// y [B][n] (reference layout): float32 for BIAWGN, uint8 for BSC / BEC.  Thread = (frame, group of 4 variables).
// x: transmitted word [n] uint8 or NULL (all-zero word).  param: p (BSC, BEC) or sqrt(noise_var) (BIAWGN).
template <int MODE>
__global__ void channel_generate(void *__restrict__ y, const uint8_t *__restrict__ x, int B, int n, double param,
                                 unsigned long long seed, unsigned long long frame0)
{
    const int groups = (n + 3) >> 2;
    const long long total = (long long)B * groups;
    const uint32_t thr = (param >= 1.0) ? 0xffffffffu : (uint32_t)(param * 4294967296.0);   // u < p  <=>  r < p * 2^32
    const float sigma = (float)param;
    const bool small = total < 0x7fffffffLL;                     // 32-bit index arithmetic (the 64-bit division costs ~40 instructions)
    const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0;             // one 128-bit / 32-bit store per thread
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int f, g;
        if (small) { f = (int)((uint32_t)i / (uint32_t)groups); g = (int)((uint32_t)i - (uint32_t)f * (uint32_t)groups); }
        else { f = (int)(i / groups); g = (int)(i % groups); }
        const Philox4 r = channel_words(seed, frame0 + (unsigned long long)f, (uint32_t)g);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == GEN_BIAWGN) {
            box_muller(r.x, r.y, &z[0], &z[1]);
            box_muller(r.z, r.w, &z[2], &z[3]);
        }
        uint32_t xw = 0u;                                        // the four transmitted bits of the group, one per byte
        if (x != nullptr) {
            if ((n & 3) == 0) xw = __ldg(reinterpret_cast<const uint32_t *>(x) + g);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * g + j < n) xw |= (uint32_t)x[4 * g + j] << (8 * j);
            }
        }
        float out_f[4];
        uint32_t out_b = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t xb = ((xw >> (8 * j)) & 0xffu) != 0u ? 1u : 0u;
            if (MODE == GEN_BIAWGN) {
                out_f[j] = fmaf(sigma, z[j], xb ? 1.0f : -1.0f);
            } else {
                const bool hit = (param >= 1.0) || (rr[j] < thr);
                out_b |= ((MODE == GEN_BSC) ? (xb ^ (hit ? 1u : 0u)) : (hit ? 2u : xb)) << (8 * j);
            }
        }
        const size_t o = (size_t)f * n + (size_t)4 * g;
        if (vec) {
            if (MODE == GEN_BIAWGN) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(y) + o) = make_float4(out_f[0], out_f[1], out_f[2], out_f[3]);
            else *reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(y) + o) = out_b;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (4 * g + j >= n) break;
                if (MODE == GEN_BIAWGN) reinterpret_cast<float *>(y)[o + j] = out_f[j];
                else reinterpret_cast<uint8_t *>(y)[o + j] = (uint8_t)(out_b >> (8 * j));
            }
        }
    }
}

// Number of bytes that differ between two words of four symbols in {0, 1, 2}.
__device__ __forceinline__ int sym4_diff(uint32_t a, uint32_t b)
{
    const uint32_t t = a ^ b;
    return __popc((t | (t >> 1)) & 0x01010101u);
}

// One lane's share of #{v : row[v] != x[v]} (x NULL = the all-zero word); 4 symbols per load when the rows are word-aligned.
__device__ __forceinline__ int row_errors(const uint8_t *__restrict__ row, const uint8_t *__restrict__ x, int n, int lane)
{
    int e = 0;
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(x)) & 3u) == 0) {
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(row), *xw = reinterpret_cast<const uint32_t *>(x);
        for (int q = lane; q < (n >> 2); q += 32) e += sym4_diff(rw[q], (x != nullptr) ? __ldg(xw + q) : 0u);
    } else {
        for (int v = lane; v < n; v += 32) e += (row[v] != ((x != nullptr) ? x[v] : (uint8_t)0)) ? 1 : 0;
    }
    return e;
}

// bit_errs[b] = #{v : x_hat[b][v] != x[v]} (src/main.py:41; an undecoded BEC symbol 2 counts as an error).  One warp per frame.
__global__ void count_errors(const uint8_t *__restrict__ x_hat, const uint8_t *__restrict__ x, int B, int n, int *__restrict__ bit_errs)
{
    const int lane = threadIdx.x & 31;
    const int f = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= B) return;
    const uint8_t *row = x_hat + (size_t)f * n;
    int e = row_errors(row, x, n, lane);
    e = __reduce_add_sync(kFull, e);
    if (lane == 0) bit_errs[f] = e;
}

// Monte-Carlo counters of one decoded batch, accumulated ON THE DEVICE (the inner loop of src/main.py:37-45 without a
// host round trip): c[0] += frames (tot), c[1] += #{frames with a bit error} (wec), c[2] += bit errors (bec),
// c[3] += sum of iteration counts, c[4 + min(iters, nhist - 1)] += 1 (the iteration histogram of stats(), admm.py:38-40).
// One warp per frame counts the frame's errors against x (as count_errors); a CTA adds its 8 frames with one set of atomics.
// bit_errs may be NULL.  Counters are unsigned 64-bit (read them as int64).
__global__ void __launch_bounds__(256) count_accumulate(const uint8_t *__restrict__ x_hat, const uint8_t *__restrict__ x,
                                                        const int *__restrict__ iters, int B, int n,
                                                        int *__restrict__ bit_errs, unsigned long long *__restrict__ c, int nhist)
{
    __shared__ int s_e[8], s_it[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * 8 + warp;
    int e = 0, it = -1;
    if (f < B) {
        e = row_errors(x_hat + (size_t)f * n, x, n, lane);
        e = __reduce_add_sync(kFull, e);
        it = iters[f];
        if (lane == 0 && bit_errs != nullptr) bit_errs[f] = e;
    }
    if (lane == 0) { s_e[warp] = e; s_it[warp] = it; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0, wec = 0, bec = 0, its = 0;
        for (int w = 0; w < 8; ++w) {
            if (s_it[w] < 0) continue;
            tot += 1; wec += s_e[w] > 0; bec += (unsigned long long)s_e[w]; its += (unsigned long long)s_it[w];
            if (nhist > 0) {                                     // one atomic per distinct iteration count of the CTA
                bool seen = false;
                unsigned long long same = 0;
                for (int u = 0; u < 8; ++u) {
                    if (s_it[u] != s_it[w]) continue;
                    if (u < w) { seen = true; break; }
                    same += 1;
                }
                if (!seen) atomicAdd(c + 4 + min(s_it[w], nhist - 1), same);
            }
        }
        atomicAdd(c + 0, tot);
        if (wec) atomicAdd(c + 1, wec);
        if (bec) atomicAdd(c + 2, bec);
        atomicAdd(c + 3, its);
    }
}

#endif  // __CUDACC__

}  // namespace ldpc
