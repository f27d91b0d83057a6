/* include/ldpc_b200.h — C ABI of libldpc_b200.so (sm_100a).
 *
 * The drop-in boundary for ONE hot path of thadikari/ldpc_decoders: the
 * iterative message-passing decoders
 *     bpa.SPA / bpa.MSA      (/root/reference/src/bpa.py:6-102)
 *     bec.SPA (= bec.MSA)    (/root/reference/src/bec.py:70-125)
 * and their channel LLR front ends (src/bsc.py:19-25, src/biawgn.py:10,21-28,
 * src/bec.py:76,85), decoding a BATCH of frames per call on one B200.
 *
 * The reference has no FFI on this path (it is numpy/scipy Python); its only
 * FFI precedent is src/parity_polytope/exact.py:12-26,41-60 (ctypes,
 * caller-allocated outputs, void returns).  This header is what a ctypes
 * binding on the reference side would load instead of calling
 * bpa.BPA.decode / bec.SPA.decode frame by frame — see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (a CUDA stream
 *     is passed as void*, NULL = the legacy default stream).
 *   - every function returns LDPC_OK (0) or a negative error code and never
 *     throws or exits; ldpc_last_error() gives the message.
 *   - "device" pointers are CUDA device pointers on the handle's device and
 *     are CALLER-OWNED (e.g. torch.empty(...).data_ptr()); the handle owns
 *     only the graph tables (and, for ldpc_decode_host, its staging buffers).
 *   - ldpc_decode / ldpc_decode_channel / ldpc_mc_round / ldpc_llr_* /
 *     ldpc_debug_step are asynchronous on the given stream; ldpc_create /
 *     ldpc_destroy / ldpc_decode_host synchronise.  ONE exception: a decode on
 *     the streaming path whose iteration bound exceeds 32 (max_iter > 32, or
 *     max_iter <= 0 = "unlimited") reads two words back every 2 - 4 iterations
 *     (cudaStreamSynchronize on the given stream): whether any frame still runs
 *     (to stop launching sweeps) and how many (active-frame compaction: once at
 *     most 70 % of the frame columns are live they are packed to the front of
 *     every row, LDPC_NO_COMPACTION=1 in the environment disables it); such a
 *     call cannot be captured into a CUDA graph.
 *     The on-chip path (one launch per batch) never synchronises.
 *   - one handle per (process, device); a handle is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device ldpc_create fails.
 *   - environment variables, read by the library for A/B measurements only
 *     (results are identical with and without them; scripts/r2_*.sh use them):
 *     LDPC_NO_COMPACTION=1 (streaming path without active-frame compaction),
 *     LDPC_RESIDENT_VP=1 (the round-1 on-chip kernels resident_vp / resident_vd
 *     instead of resident_vq), LDPC_RESIDENT_LAYOUT=check (check-major on-chip
 *     layout, resident_bp), LDPC_BEC_WIDE=1 (608-thread geometry of the on-chip
 *     erasure kernel), LDPC_PLAN_EFFORT=<float> (placement annealing effort,
 *     read at ldpc_create).
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_ABI_VERSION 2

/* algo: which reference decoder is replaced */
#define LDPC_MSA 0   /* bpa.MSA.decode_      src/bpa.py:86-102  */
#define LDPC_SPA 1   /* bpa.SPA.decode_      src/bpa.py:71-75   */
#define LDPC_BEC 2   /* bec.SPA.decode       src/bec.py:83-122  */

/* dtype: arithmetic type of the messages (the reference is dtype-polymorphic
 * through priors.dtype).
 *   MSA: F32 / F64 are both bit-exact restatements at that dtype.
 *   SPA: F64 mirrors the reference formula (tanh / log / exp / atanh);
 *        F32 is the production form (hyperbolic-pair rule, cancellation-free).
 *   BEC: dtype is ignored (bit-plane integer arithmetic). */
#define LDPC_F32 0
#define LDPC_F64 1
#define LDPC_F16 2   /* y_dtype ONLY (received values / priors handed over as IEEE binary16): halves the bytes a host
                        batch sends over PCIe.  Converted exactly; messages stay F32 / F64. */

/* per-frame exit reason written by ldpc_decode (the reference's ret('...')
 * strings, src/bpa.py:28-29, src/bec.py:96-97,120) */
#define LDPC_REASON_DECODED  0
#define LDPC_REASON_MAXIMUM  1
#define LDPC_REASON_STOPPING 2   /* BEC only */
#define LDPC_REASON_CAP      4   /* stopped by iter_cap while max_iter <= 0 (reference: unlimited) */

/* error codes */
#define LDPC_OK           0
#define LDPC_EINVAL      -1
#define LDPC_ECUDA       -2
#define LDPC_ENOMEM      -3
#define LDPC_EWORKSPACE  -4   /* workspace too small / misaligned */
#define LDPC_EUNSUPPORTED -5

/* flags for ldpc_decode */
#define LDPC_PATH_AUTO      0u   /* resident kernel when the code fits on chip, else streaming */
#define LDPC_PATH_STREAMING 1u   /* edge-major [E][B] messages in HBM, one CN + one VN sweep per iteration */
#define LDPC_PATH_RESIDENT  2u   /* whole frames kept in shared memory for all iterations (short codes; float32 MSA / SPA,
                                    float64 MSA on regular codes) */
#define LDPC_PATH_MASK      3u
#define LDPC_SPA_ROBUST     4u   /* float32 SPA: do not emulate the reference's float64 tanh saturation (|v| > 38.123
                                    contributes exactly 0, which is what floods a frame with inf/NaN in the reference);
                                    honoured by the resident path */
#define LDPC_CN_REGISTER    8u   /* streaming path: use the register-staged check-node sweep instead of the
                                    bulk-async (TMA) staged one (A/B measurements; it is also the fallback for
                                    check degrees > 8) */
#define LDPC_HOST_ASYNC    16u   /* ldpc_decode_host: return once the work is enqueued; ldpc_host_sync() completes it */
#define LDPC_OUT_PACKED    32u   /* ldpc_decode_host: x_hat is BIT-PACKED, one row of ldpc_packed_row_bytes(n) bytes per frame
                                    (bit v of a word = bit (v & 7) of byte (v >> 3): numpy.packbits(bitorder="little"));
                                    BEC: two such planes per frame, value plane (symbol == 1) then erasure plane (symbol == 2) */
#define LDPC_IN_PACKED     64u   /* ldpc_decode_host, BSC / BEC: y is bit-packed the same way (1 bit per hard bit, 2 planes
                                    per erasure symbol) instead of one byte per symbol */

/* channel kinds for ldpc_decode_host / ldpc_channel_llr */
#define LDPC_CH_PRIORS 0   /* input already is the prior LLR (bpa.*.decode(y, priors)) */
#define LDPC_CH_BSC    1   /* input y in {0,1};  priors = llr * (1 - 2y),  param = llr = log(1-p) - log(p)  (bsc.py:21,25) */
#define LDPC_CH_BIAWGN 2   /* input y real;      priors = (-2y) / param,   param = noise_var = 10**(-snr/10) (biawgn.py:10,28) */
#define LDPC_CH_BEC    3   /* input y in {0,1,2 = erasure} (bec.py:76,85) */

typedef struct ldpc_handle ldpc_t;

int ldpc_abi_version(void);

/* Build a decoder handle for one parity-check matrix H (replaces
 * bpa.BPA.__init__, src/bpa.py:9-15, and bec.SPA.__init__, src/bec.py:73-81).
 * Tables are HOST pointers (copied):
 *   edge e = position in np.where(H) order (check-major, ascending variable)
 *   chk_ptr[m+1], edge_var[E]     check-major CSR of H
 *   var_ptr[n+1], var_edges[E]    per variable, its edge ids in ascending order
 * Validates the tables (monotone pointers, sorted, in range). */
int ldpc_create(ldpc_t **out, int device, int n, int m, int E,
                const int32_t *chk_ptr, const int32_t *edge_var,
                const int32_t *var_ptr, const int32_t *var_edges);

void ldpc_destroy(ldpc_t *h);

/* Last error message of this handle (h == NULL: of the last failed ldpc_create). */
const char *ldpc_last_error(const ldpc_t *h);

/* Bytes of device workspace ldpc_decode needs for B frames (0 on bad arguments). */
size_t ldpc_workspace_bytes(const ldpc_t *h, int algo, int dtype, int B, unsigned flags);

/* Decode B frames (replaces B calls of bpa.BPA.decode(y, priors), src/bpa.py:17-63,
 * or bec.SPA.decode(y), src/bec.py:83-122).
 *   input    device [B,n] row-major: MSA/SPA priors of `dtype`; BEC uint8 symbols {0,1,2}
 *   y_hard   device [B,n] uint8 hard received bits, or NULL.  MSA/SPA only: the
 *            iteration-0 syndrome test (src/bpa.py:29 on x_hat = y) is run on it;
 *            NULL skips that test (BIAWGN: real-valued y never passes it).
 *   max_iter reference semantics: `0 < max_iter <= it` stops; max_iter <= 0 is
 *            unlimited in the reference, here it runs until iter_cap (> 0 required).
 *   x_hat    device [B,n] uint8: decoded bits (BEC: symbols, 2 = still erased).
 *            A frame that exits at iteration 0 returns y_hard / the input symbols.
 *   iters    device [B] int32: the reference's iter_count at return.
 *   reason   device [B] uint8 LDPC_REASON_*, or NULL.
 *   marg_out device [B,n] of `dtype`, or NULL: last marginal of every frame
 *            (before the NaN scrub, src/bpa.py:35), for verification.  MSA/SPA only.
 */
int ldpc_decode(ldpc_t *h, int algo, int dtype,
                const void *input, const uint8_t *y_hard, int B,
                int max_iter, int iter_cap,
                uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *marg_out,
                void *workspace, size_t workspace_bytes, unsigned flags, void *stream);

/* Same as ldpc_decode, with the channel front end fused into the load: `y` is the RECEIVED block on
 * the device ([B,n] row-major; uint8 for BSC / BEC, y_dtype F32 | F64 | F16 for BIAWGN / PRIORS) and the
 * LLR map of the reference's adapters (src/bsc.py:25, src/biawgn.py:28, src/bec.py:85) is applied while
 * it is transposed into the kernels' layout.  channel = LDPC_CH_*, param = llr (BSC) / noise_var
 * (BIAWGN).  BSC uses y as the hard input of the iteration-0 syndrome test. */
int ldpc_decode_channel(ldpc_t *h, int channel, int algo, int dtype, double param,
                        const void *y, int y_dtype, int B, int max_iter, int iter_cap,
                        uint8_t *x_hat, int32_t *iters, uint8_t *reason, void *marg_out,
                        void *workspace, size_t workspace_bytes, unsigned flags, void *stream);

/* Channel LLR front ends on device buffers (elementwise, any shape, `count` elements).
 *   BSC:    y uint8 {0,1}             -> priors(dtype) = llr * (1 - 2y)       src/bsc.py:25
 *   BIAWGN: y of y_dtype (F32 | F64)  -> priors(dtype) = (-2 y) / noise_var   src/biawgn.py:28
 * computed in float64 and rounded once to `dtype` (== reference priors.astype(dtype)). */
int ldpc_llr_bsc(ldpc_t *h, int dtype, double llr, const uint8_t *y, void *priors, size_t count, void *stream);
int ldpc_llr_biawgn(ldpc_t *h, int y_dtype, int dtype, double noise_var, const void *y, void *priors,
                    size_t count, void *stream);

/* On-device channel simulators (the reference's Channel.send, src/bec.py:15-18, src/bsc.py:15-16,
 * src/biawgn.py:17-18) with counter-based Philox4x32-10 noise: statistically, not bit-, equivalent to numpy's global
 * RNG.  The noise of a received value depends only on (seed, frame0 + frame, variable), so a Monte-Carlo run gives
 * the same frames however it is cut into batches or spread over GPUs.
 *   channel  LDPC_CH_BSC / LDPC_CH_BEC: param = p, y uint8 [B,n];  LDPC_CH_BIAWGN: param = noise_var, y float32 [B,n]
 *   x        device [n] uint8 transmitted word, or NULL = the all-zero word */
int ldpc_channel_generate(ldpc_t *h, int channel, double param, const uint8_t *x,
                          unsigned long long seed, unsigned long long frame0, int B, void *y, void *stream);

/* bit_errs[b] = #{v : x_hat[b,v] != x[v]} (src/main.py:41; an undecoded BEC symbol counts); x NULL = all-zero word. */
int ldpc_count_errors(ldpc_t *h, const uint8_t *x_hat, const uint8_t *x, int B, int32_t *bit_errs, void *stream);

/* Monte-Carlo counters of a decoded batch, accumulated on the device (the loop body of src/main.py:37-45, B frames at
 * once, no host round trip): counters is device int64 [4 + nhist]:
 *   [0] += B (tot)   [1] += #{frames with a bit error} (wec)   [2] += bit errors (bec)   [3] += sum of iters
 *   [4 + min(iters, nhist - 1)] += 1   (the iteration histogram of stats(), src/admm.py:38-40; nhist may be 0)
 * bit_errs: device [B] int32 per-frame error counts, or NULL.  x NULL = the all-zero word. */
int ldpc_count_accumulate(ldpc_t *h, const uint8_t *x_hat, const uint8_t *x, const int32_t *iters, int B,
                          int32_t *bit_errs, long long *counters, int nhist, void *stream);

/* One whole Monte-Carlo round on the device: ldpc_channel_generate (frames frame0 .. frame0 + B - 1 of the word x) ->
 * ldpc_decode_channel -> ldpc_count_accumulate into `counters`; nothing is copied to the host.  A fixed-length
 * simulation is a loop of these and ONE read (or one all-reduce over GPUs) of `counters` at the end.
 *   ch_param   channel parameter of ldpc_channel_generate (p, or noise_var for BIAWGN)
 *   dec_param  decoder parameter of ldpc_decode_channel (llr = log(1-p) - log(p) for BSC, noise_var for BIAWGN)
 *   scratch    device, >= ldpc_mc_scratch_bytes(...) bytes, 256-byte aligned (received block, words, workspace) */
size_t ldpc_mc_scratch_bytes(const ldpc_t *h, int channel, int algo, int dtype, int B);
int ldpc_mc_round(ldpc_t *h, int channel, int algo, int dtype, double ch_param, double dec_param,
                  const uint8_t *x, unsigned long long seed, unsigned long long frame0, int B,
                  int max_iter, int iter_cap, long long *counters, int nhist,
                  void *scratch, size_t scratch_bytes, unsigned flags, void *stream);

/* One isolated sweep on caller-supplied messages in the reference's own layout
 * (teacher-forced parity): device [B,E] row-major, edge order of np.where(H).
 *   which = 0: check-node sweep   c2v = CN(v2c)            (src/bpa.py:71-75 / 86-102)
 *   which = 1: variable-node sweep: marg = prior + sum c2v, v2c = marg[yy] - c2v (src/bpa.py:35-37)
 * Runs the same device math as the streaming kernels.  Synchronous-free, on `stream`. */
int ldpc_debug_step(ldpc_t *h, int algo, int dtype, int which, int B,
                    const void *prior /* [B,n], which=1 */, const void *msg_in /* [B,E] */,
                    void *msg_out /* [B,E] */, void *marg /* [B,n], which=1, may be NULL */,
                    void *workspace, size_t workspace_bytes, void *stream);

/* End-to-end call with HOST buffers (what the Python decode_batch uses for numpy
 * input): chunks the batch, overlaps H2D copy / LLR + decode / D2H copy on internal
 * streams, and returns when x_hat / iters are complete in host memory.
 *   channel  LDPC_CH_*; param = llr (BSC) or noise_var (BIAWGN), ignored otherwise
 *   y        host [B,n] row-major; y_dtype: LDPC_F32 / LDPC_F64 / LDPC_F16 for PRIORS and BIAWGN,
 *            ignored (uint8) for BSC and BEC — or, with LDPC_IN_PACKED, [B][planes * ldpc_packed_row_bytes(n)]
 *            bit-packed symbols.  Pinned memory makes the copies asynchronous.
 *   x_hat    host [B,n] uint8 (LDPC_OUT_PACKED: [B][planes * ldpc_packed_row_bytes(n)]);
 *            iters host [B] int32;  reason host [B] uint8 or NULL
 *   chunk    frames per pipeline stage (0 = default)
 * Device staging buffers are owned by the handle and reused across calls (in stream order). */
int ldpc_decode_host(ldpc_t *h, int channel, int algo, int dtype, double param,
                     const void *y, int y_dtype, int B, int max_iter, int iter_cap,
                     uint8_t *x_hat, int32_t *iters, uint8_t *reason,
                     int chunk, unsigned flags);

/* Bytes of one bit-packed row (LDPC_IN_PACKED / LDPC_OUT_PACKED): ceil(n / 8) rounded up to a multiple of 16. */
size_t ldpc_packed_row_bytes(int n);

/* A stream of batches: with LDPC_HOST_ASYNC in `flags`, ldpc_decode_host returns as soon as every chunk is enqueued on
 * the handle's internal streams, so the next call's copies overlap this call's tail (no pipeline ramp between batches).
 * The host buffers of every such call (pinned) must stay valid and untouched until ldpc_host_sync returns, which waits
 * for everything enqueued so far.  Results are the blocking call's. */
int ldpc_host_sync(ldpc_t *h);

/* Per-launch timing of the two sweeps, for bench.py's roofline.  While enabled, ldpc_decode records CUDA
 * events around every check-node and variable-node sweep launch on the stream it runs on;
 * ldpc_profile_read waits for them and returns the accumulated milliseconds and launch counts
 * since the previous read (kind 0 = check-node sweep, 1 = variable-node sweep). */
int ldpc_profile_enable(ldpc_t *h, int on);
int ldpc_profile_read(ldpc_t *h, double *cn_ms, unsigned long long *cn_launches,
                      double *vn_ms, unsigned long long *vn_launches);

/* Frames a CTA of the float32 on-chip path (LDPC_PATH_RESIDENT) keeps in shared memory for this code (the float64
 * min-sum kernel keeps half as many), or 0 when the code does
 * not fit on chip (then LDPC_PATH_AUTO streams and LDPC_PATH_RESIDENT is refused with LDPC_EUNSUPPORTED). */
int ldpc_resident_frames(const ldpc_t *h);

/* Shared-memory placement of the on-chip path (csrc/res_layout.h): predicted 128-bit shared-memory wavefronts per
 * iteration of the two gather phases, out[7] = { check-node ideal, file order, planned positions with natural edge
 * order (sum-product), planned (min-sum), variable-node ideal, file order, planned }.  Returns LDPC_EUNSUPPORTED when
 * the code has no on-chip path.  The annealing effort can be set with the environment variable LDPC_PLAN_EFFORT
 * (default 0.4; 0 keeps file order) before ldpc_create. */
int ldpc_resident_plan(const ldpc_t *h, long *out);

/* Layout family of the on-chip path of this code: "resident_vp" (variable-plane message layout, regular and irregular
 * codes: csrc/resident_vp.cuh and the kernel that runs on it by default, resident_vq, csrc/resident_vq.cuh),
 * "resident_bp" (check-major layout, any degree profile <= 8, csrc/resident_bp.cuh) or "" when the code has no on-chip
 * path.  The string is static. */
const char *ldpc_resident_kernel(const ldpc_t *h);

/* Number of kernel launches issued through this handle since creation (bench.py's gpu_launches). */
unsigned long long ldpc_launch_count(const ldpc_t *h);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
