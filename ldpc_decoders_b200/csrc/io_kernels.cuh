// io_kernels.cuh — layout changes at the boundary and the channel LLR front ends.
//
// The reference-facing layout is frame-major [B][n] (one numpy row per frame); the kernels want
// variable-major [n][Bp] with frames contiguous.  These kernels transpose through a 32x32 shared
// tile (both sides coalesced) and fuse the LLR map of the channel adapters into the load:
//   BSC     priors = llr * (1 - 2y)            /root/reference/src/bsc.py:25
//   BIAWGN  priors = (-2 y) / noise_var        /root/reference/src/biawgn.py:28
//   BEC     messages[y] = [-1, +1, 0][y]       /root/reference/src/bec.py:76,85
// evaluated in float64 and rounded once to the message type (== reference priors.astype(dtype)).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ldpc {

enum { IN_COPY = 0, IN_BSC = 1, IN_BIAWGN = 2 };

// A received value as the float64 the reference computes with.  binary16 rows (y_dtype LDPC_F16) halve the bytes a
// host batch sends over PCIe; the conversion is exact, so the parity definition is unchanged: priors = (-2 y) / var on
// the values the caller handed over.
template <typename Tin> __device__ __forceinline__ double in_f64(Tin y) { return (double)y; }
template <> __device__ __forceinline__ double in_f64<__half>(__half y) { return (double)__half2float(y); }
template <typename Tin> __device__ __forceinline__ bool in_nonzero(Tin y) { return y != (Tin)0; }
template <> __device__ __forceinline__ bool in_nonzero<__half>(__half y) { return __half2float(y) != 0.0f; }

template <typename Tin, typename T, int MODE> __device__ __forceinline__ T llr_map(Tin y, double param)
{
    if (MODE == IN_BSC) return (T)(param * (double)(1 - 2 * (int)in_nonzero(y)));
    if (MODE == IN_BIAWGN) return (T)((-2.0 * in_f64(y)) / param);
    return (T)in_f64(y);
}

// src [B][n] (frame-major) -> prior [n][Bp]; frames >= B are zero-filled.
// MODE == IN_BSC additionally packs the hard bits y into xbits [n][wpr].
// A CTA transposes kIngestTiles 32 x 32 tiles side by side along the variable axis: all their loads are issued before
// the barrier (16 per thread in flight instead of 4), which is what this latency-bound copy needs.
// grid (ceil(n / (32 * kIngestTiles)), Bp/32), block (32, 8).
constexpr int kIngestTiles = 4;
template <typename Tin, typename T, int MODE>
__global__ void __launch_bounds__(256) ingest_priors(const Tin *__restrict__ src, T *__restrict__ prior, uint32_t *__restrict__ xbits,
                                                     int B, int n, int Bp, int wpr, double param)
{
    __shared__ T tile[kIngestTiles][32][33];
    __shared__ uint8_t hard[MODE == IN_BSC ? kIngestTiles : 1][32][33];
    const int vb = blockIdx.x * 32 * kIngestTiles, f0 = blockIdx.y * 32;
#pragma unroll
    for (int t = 0; t < kIngestTiles; ++t) {
#pragma unroll
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int f = f0 + r, v = vb + 32 * t + threadIdx.x;
            T val = (T)0;
            uint8_t hb = 0;
            if (f < B && v < n) {
                const Tin y = src[(size_t)f * n + v];
                val = llr_map<Tin, T, MODE>(y, param);
                if (MODE == IN_BSC) hb = (uint8_t)in_nonzero(y);
            }
            tile[t][r][threadIdx.x] = val;
            if (MODE == IN_BSC) hard[t][r][threadIdx.x] = hb;
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kIngestTiles; ++t) {
#pragma unroll
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int v = vb + 32 * t + r, f = f0 + threadIdx.x;
            if (v < n) {                                                 // warp-uniform
                prior[(size_t)v * Bp + f] = tile[t][threadIdx.x][r];
                if (MODE == IN_BSC) {
                    const uint32_t w = __ballot_sync(kFull, hard[t][threadIdx.x][r] != 0);
                    if (threadIdx.x == 0) xbits[(size_t)v * wpr + (f0 >> 5)] = w;
                }
            }
        }
    }
}

// y_hard [B][n] uint8 -> xbits [n][wpr]   (the x_hat = y of src/bpa.py:20)
__global__ void pack_hard(const uint8_t *__restrict__ y, uint32_t *__restrict__ xbits, int B, int n, int wpr)
{
    __shared__ uint8_t hard[32][33];
    const int v0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int f = f0 + r, v = v0 + threadIdx.x;
        hard[r][threadIdx.x] = (f < B && v < n) ? (uint8_t)(y[(size_t)f * n + v] != 0) : (uint8_t)0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int v = v0 + r;
        if (v < n) {
            const uint32_t w = __ballot_sync(kFull, hard[threadIdx.x][r] != 0);
            if (threadIdx.x == 0) xbits[(size_t)v * wpr + (f0 >> 5)] = w;
        }
    }
}

// BEC symbols [B][n] uint8 {0,1,2} -> prior planes, x_hat planes and the per-frame "has erasures" flag.
__global__ void ingest_bec(const uint8_t *__restrict__ y, uint32_t *__restrict__ pnz, uint32_t *__restrict__ ppos,
                           uint32_t *__restrict__ xe, uint32_t *__restrict__ xv, uint32_t *__restrict__ haser,
                           int B, int n, int wpr)
{
    __shared__ uint8_t sym[32][33];
    const int v0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int f = f0 + r, v = v0 + threadIdx.x;
        sym[r][threadIdx.x] = (f < B && v < n) ? y[(size_t)f * n + v] : (uint8_t)0;
    }
    __syncthreads();
    const bool valid = (f0 + (int)threadIdx.x) < B;
    uint32_t er_any = 0u;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int v = v0 + r;
        if (v < n) {
            const uint8_t s = sym[threadIdx.x][r];
            const uint32_t e = __ballot_sync(kFull, valid && s >= 2);
            const uint32_t one = __ballot_sync(kFull, valid && s == 1);
            const uint32_t ok = __ballot_sync(kFull, valid);
            if (threadIdx.x == 0) {
                const size_t i = (size_t)v * wpr + (f0 >> 5);
                pnz[i] = ok & ~e; ppos[i] = one;
                xe[i] = e; xv[i] = one;
            }
            er_any |= e;
        }
    }
    if (threadIdx.x == 0 && er_any != 0u) atomicOr(haser + (f0 >> 5), er_any);
}

// ---------------------------------------------------------------------------------------------------------------
// Tiled, vectorised versions of ingest_bec and emit_words for rows whose length is a multiple of 4 (every shipped code).
// The byte-per-thread kernels above move 1 B per lane and write single 4-byte words 4 * wpr bytes apart: 266 GB/s and
// 500 GB/s on 131072 frames of n = 1200, 39 % of an erasure-decoding step (profiles/README.md).  Here a CTA of 256
// threads owns a tile of 256 frames x 32 variables: symbol rows move as 32-byte sectors (uint32 per lane), the bit
// planes as 32-byte runs (the 8 frame-words of one variable), and the transpose happens in shared memory.
// grid (ceil(n/32), ceil(wpr/8)), block 256.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTileFrames = 256, kTileVars = 32, kTileRowWords = kTileVars / 4 + 1;     // 9 words per frame row: conflict-free columns

__global__ void __launch_bounds__(256) ingest_bec_tiled(const uint8_t *__restrict__ y, uint32_t *__restrict__ pnz,
                                                        uint32_t *__restrict__ ppos, uint32_t *__restrict__ xe,
                                                        uint32_t *__restrict__ xv, uint32_t *__restrict__ haser,
                                                        int B, int n, int wpr)
{
    __shared__ uint32_t sym[kTileFrames][kTileRowWords];
    __shared__ uint32_t out_e[kTileVars][8], out_1[kTileVars][8], out_ok[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v0 = blockIdx.x * kTileVars, f0 = blockIdx.y * kTileFrames;
    // ---- symbols in: thread = (frame row, 4 consecutive variables)
#pragma unroll
    for (int i = 0; i < kTileFrames * 8 / 256; ++i) {
        const int idx = tid + i * 256, r = idx >> 3, q = idx & 7;
        const int f = f0 + r, v = v0 + 4 * q;
        uint32_t w = 0u;
        if (f < B && v < n) w = *reinterpret_cast<const uint32_t *>(y + (size_t)f * n + v);      // n % 4 == 0: whole word in range
        sym[r][q] = w;
    }
    __syncthreads();
    // ---- warp = one frame-word (32 frames), lane = frame: two ballots per variable
    const bool valid = (f0 + 32 * warp + lane) < B;
    const uint32_t ok = __ballot_sync(kFull, valid);
    const uint8_t *row = reinterpret_cast<const uint8_t *>(&sym[32 * warp + lane][0]);
    uint32_t er_any = 0u;
#pragma unroll 8
    for (int vv = 0; vv < kTileVars; ++vv) {
        const uint8_t sy = row[vv];
        const uint32_t e = __ballot_sync(kFull, valid && sy >= 2);
        const uint32_t one = __ballot_sync(kFull, valid && sy == 1);
        if (lane == 0) { out_e[vv][warp] = e; out_1[vv][warp] = one; }
        er_any |= e;
    }
    if (lane == 0) {
        out_ok[warp] = ok;
        const int word = (f0 >> 5) + warp;
        if (word < wpr && er_any != 0u) atomicOr(haser + word, er_any);
    }
    __syncthreads();
    // ---- planes out: thread = (variable, frame-word); 8 consecutive threads write one 32-byte run
    const int vv = tid >> 3, w8 = tid & 7, word = (f0 >> 5) + w8;
    if (v0 + vv < n && word < wpr) {
        const size_t i = (size_t)(v0 + vv) * wpr + word;
        const uint32_t e = out_e[vv][w8], one = out_1[vv][w8];
        pnz[i] = out_ok[w8] & ~e; ppos[i] = one;
        xe[i] = e; xv[i] = one;
    }
}

// bit planes -> x_hat [B][n] bytes (symbol 2 where xer is set), same tile.
// orig (optional): column f of the planes holds frame orig[f] (active-frame compaction, stream_bp.cuh); columns whose
// frame index is >= B are padding.
__global__ void __launch_bounds__(256) emit_words_tiled(const uint32_t *__restrict__ xval, const uint32_t *__restrict__ xer,
                                                        uint8_t *__restrict__ x_hat, int B, int n, int wpr,
                                                        const int *__restrict__ orig = nullptr, int extent = 0x7fffffff)
{
    __shared__ uint32_t in_v[kTileVars][8], in_e[kTileVars][8];
    __shared__ uint32_t sym[kTileFrames][kTileRowWords];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v0 = blockIdx.x * kTileVars, f0 = blockIdx.y * kTileFrames;
    {
        const int vv = tid >> 3, w8 = tid & 7, word = (f0 >> 5) + w8;
        uint32_t val = 0u, er = 0u;
        if (v0 + vv < n && word < wpr) {
            val = xval[(size_t)(v0 + vv) * wpr + word];
            if (xer != nullptr) er = xer[(size_t)(v0 + vv) * wpr + word];
        }
        in_v[vv][w8] = val; in_e[vv][w8] = er;
    }
    __syncthreads();
    // warp = frame-word, lane = frame: build the frame's 32 symbols, four per 32-bit word
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint32_t w = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int vv = 4 * q + b;
            const uint32_t bit = (in_v[vv][warp] >> lane) & 1u, er = (in_e[vv][warp] >> lane) & 1u;
            w |= (er ? 2u : bit) << (8 * b);
        }
        sym[32 * warp + lane][q] = w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kTileFrames * 8 / 256; ++i) {
        const int idx = tid + i * 256, r = idx >> 3, q = idx & 7;
        int f = f0 + r;
        const int v = v0 + 4 * q;
        if (orig != nullptr) f = (f < extent) ? orig[f] : B;
        if (f < B && v < n) *reinterpret_cast<uint32_t *>(x_hat + (size_t)f * n + v) = sym[r][q];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Bit-packed rows at the HOST boundary (ldpc_decode_host with LDPC_IN_PACKED / LDPC_OUT_PACKED): a hard bit costs one
// bit on PCIe instead of one byte, an erasure symbol two.  Layout of a packed row (stride = ldpc_packed_row_bytes(n),
// a multiple of 16): bit v of the row = bit (v & 7) of byte (v >> 3), i.e. numpy.packbits(..., bitorder="little");
// BEC rows are two such planes back to back: the value plane (symbol == 1), then the erasure plane (symbol == 2).
// Both kernels are a few tens of microseconds per 32768 frames: the decoders keep their byte interface.
// ---------------------------------------------------------------------------------------------------------------
// src [B][planes * stride] -> dst [B][n] bytes.  Thread = (frame, 8 consecutive variables = one packed byte).
__global__ void unpack_rows(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int B, int n, int stride, int planes)
{
    const int bpr = (n + 7) >> 3;                                    // packed bytes that carry symbols
    const long long total = (long long)B * bpr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / bpr), q = (int)(i % bpr);
        const uint8_t *row = src + (size_t)f * planes * stride;
        const uint32_t val = row[q], er = (planes == 2) ? row[stride + q] : 0u;
        uint8_t *out = dst + (size_t)f * n + (size_t)q * 8;
        const int cnt = min(8, n - q * 8);
        if (cnt == 8 && (n & 7) == 0) {                              // 8-byte aligned: one 64-bit store
            uint32_t lo = 0u, hi = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                lo |= (((er >> b) & 1u) ? 2u : ((val >> b) & 1u)) << (8 * b);
                hi |= (((er >> (b + 4)) & 1u) ? 2u : ((val >> (b + 4)) & 1u)) << (8 * b);
            }
            *reinterpret_cast<uint2 *>(out) = make_uint2(lo, hi);
        } else {
            for (int b = 0; b < cnt; ++b) out[b] = (uint8_t)(((er >> b) & 1u) ? 2u : ((val >> b) & 1u));
        }
    }
}

// src [B][n] bytes -> dst [B][planes * stride] (padding bits and bytes zero).  Warp = one frame, lane = variable.
__global__ void pack_rows(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int B, int n, int stride, int planes)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int words = stride >> 2;
    for (int f = blockIdx.x * wpb + (threadIdx.x >> 5); f < B; f += gridDim.x * wpb) {
        const uint8_t *row = src + (size_t)f * n;
        uint32_t *out = reinterpret_cast<uint32_t *>(dst + (size_t)f * planes * stride);
        for (int w = 0; w < words; ++w) {
            const int v = w * 32 + lane;
            const uint8_t sy = (v < n) ? row[v] : (uint8_t)0;
            const uint32_t one = __ballot_sync(kFull, sy == 1), er = __ballot_sync(kFull, sy >= 2);
            if (lane == 0) {
                out[w] = one;
                if (planes == 2) out[words + w] = er;
            }
        }
    }
}

// act = frames < B; unsat = act (iteration-0 syndrome skipped) or 0; iters = 0.
__global__ void init_flags(uint32_t *act, uint32_t *unsat, int *iters, int B, int Bp, int wpr, int unsat_all)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Bp) iters[i] = 0;
    if (i < wpr) {
        const int lo = i * 32;
        const uint32_t m = (B >= lo + 32) ? 0xffffffffu : (B <= lo ? 0u : ((1u << (B - lo)) - 1u));
        act[i] = m;
        if (unsat != nullptr) unsat[i] = unsat_all ? m : 0u;
    }
}

// xbits [n][wpr] (and, for BEC, the erased plane) -> x_hat [B][n] uint8.  grid (ceil(n/32), wpr), block (32, 8).
__global__ void emit_words(const uint32_t *__restrict__ xval, const uint32_t *__restrict__ xer,
                           uint8_t *__restrict__ x_hat, int B, int n, int wpr,
                           const int *__restrict__ orig = nullptr, int extent = 0x7fffffff)
{
    const int v = blockIdx.x * 32 + threadIdx.x, w = blockIdx.y;
    uint32_t val = 0u, er = 0u;
    if (v < n) {
        val = xval[(size_t)v * wpr + w];
        if (xer != nullptr) er = xer[(size_t)v * wpr + w];
    }
    for (int l = threadIdx.y; l < 32; l += 8) {
        int f = w * 32 + l;
        if (orig != nullptr) f = (f < extent) ? orig[f] : B;
        if (f < B && v < n) x_hat[(size_t)f * n + v] = ((er >> l) & 1u) ? (uint8_t)2 : (uint8_t)((val >> l) & 1u);
    }
}

// iters / exit reason per frame.  still-active frames hit the loop bound: MAXIMUM (or CAP when unlimited).
// orig / extent: see emit_words_tiled (thread = column f, output index = the frame it holds).
__global__ void emit_status(const int *__restrict__ iters_ws, const uint32_t *__restrict__ act,
                            const uint32_t *__restrict__ stopped, int *__restrict__ iters, uint8_t *__restrict__ reason,
                            int B, int bound_reason, const int *__restrict__ orig = nullptr, int extent = 0x7fffffff)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= extent) return;
    const int o = (orig != nullptr) ? orig[f] : f;
    if (o >= B) return;
    iters[o] = iters_ws[f];
    if (reason != nullptr) {
        uint8_t r = LDPC_REASON_DECODED;
        if (stopped != nullptr && ((stopped[f >> 5] >> (f & 31)) & 1u)) r = LDPC_REASON_STOPPING;
        else if ((act[f >> 5] >> (f & 31)) & 1u) r = (uint8_t)bound_reason;
        reason[o] = r;
    }
}

// Plain tiled transposes (debug step, marginals): a [R][C] row-major -> b [C][ldb] row-major.
template <typename T>
__global__ void transpose_tile(const T *__restrict__ a, T *__restrict__ b, int R, int C, int lda, int ldb)
{
    __shared__ T tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int rr = r0 + r, cc = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (rr < R && cc < C) ? a[(size_t)rr * lda + cc] : (T)0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int cc = c0 + r, rr = r0 + threadIdx.x;
        if (cc < C && rr < R) b[(size_t)cc * ldb + rr] = tile[threadIdx.x][r];
    }
}

// Elementwise LLR front ends on flat buffers (ldpc_llr_bsc / ldpc_llr_biawgn).
template <typename Tin, typename T, int MODE>
__global__ void llr_flat(const Tin *__restrict__ y, T *__restrict__ out, size_t count, double param)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = llr_map<Tin, T, MODE>(y[i], param);
}

}  // namespace ldpc
