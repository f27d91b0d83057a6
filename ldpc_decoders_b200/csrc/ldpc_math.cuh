// ldpc_math.cuh — per-frame node arithmetic shared by every kernel.
//
// These are the order-of-operations contracts of the reference restated for
// one frame held in registers:
//   cn_msa      bpa.MSA.decode_   /root/reference/src/bpa.py:86-102, math_utils.py:10,38-43,78-94
//   cn_spa_ref  bpa.SPA.decode_   /root/reference/src/bpa.py:71-75,   math_utils.py:38-60   (formula mirror)
//   cn_spa_sc   same function for float32, cancellation-free hyperbolic-pair rule (the mirror is unusable in fp32)
//   vn_update   bpa.BPA.decode    /root/reference/src/bpa.py:35-37
//   bec_*       bec.SPA.decode    /root/reference/src/bec.py:100-119, 32 frames per machine word (bit planes)
//
// Everything is __host__ __device__ so tests/host_emu can run the very same
// arithmetic on the CPU against the oracle before any GPU time is spent.  The
// product library only ever calls these from device code.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LDPC_HD __host__ __device__ __forceinline__
#else
#define LDPC_HD inline
#endif

namespace ldpc {

template <typename T> struct num;
template <> struct num<float> {
    static LDPC_HD float inf() { return INFINITY; }
    static LDPC_HD float nan() { return NAN; }
    static LDPC_HD float abs(float a) { return fabsf(a); }
#if defined(__CUDA_ARCH__)
    static LDPC_HD float add(float a, float b) { return __fadd_rn(a, b); }   // never contracted / reassociated
    static LDPC_HD float sub(float a, float b) { return __fsub_rn(a, b); }
    static LDPC_HD float mul(float a, float b) { return __fmul_rn(a, b); }
#else
    static LDPC_HD float add(float a, float b) { volatile float r = a + b; return r; }
    static LDPC_HD float sub(float a, float b) { volatile float r = a - b; return r; }
    static LDPC_HD float mul(float a, float b) { volatile float r = a * b; return r; }
#endif
};
template <> struct num<double> {
    static LDPC_HD double inf() { return (double)INFINITY; }
    static LDPC_HD double nan() { return (double)NAN; }
    static LDPC_HD double abs(double a) { return fabs(a); }
#if defined(__CUDA_ARCH__)
    static LDPC_HD double add(double a, double b) { return __dadd_rn(a, b); }
    static LDPC_HD double sub(double a, double b) { return __dsub_rn(a, b); }
#else
    static LDPC_HD double add(double a, double b) { volatile double r = a + b; return r; }
    static LDPC_HD double sub(double a, double b) { volatile double r = a - b; return r; }
#endif
};

// ---------------------------------------------------------------------------------------------
// Min-sum check node.  Two-minimum + sign parity; exact in any IEEE type.
//   par      = #(v < 0) mod 2                    (math_utils.py:40; -0.0 and NaN count as "+")
//   m1, m2   = smallest and second smallest |v|  (first-occurrence arg-min, math_utils.py:91-93)
//   out[k]   = (sall / sown) * (k is the arg-min ? m2 : m1)          (bpa.py:88,94,100,102)
// Value-wise "k is the first arg-min" == "|v[k]| == m1": with a tie m2 == m1.
// ---------------------------------------------------------------------------------------------
template <typename T> struct MsaAcc {
    T m1, m2;
    unsigned par;
    LDPC_HD void init() { m1 = num<T>::inf(); m2 = num<T>::inf(); par = 0u; }
    LDPC_HD void push(T v) {
        const T a = num<T>::abs(v);
        par ^= (v < (T)0) ? 1u : 0u;
        const bool lt1 = a < m1, lt2 = a < m2;
        m2 = lt1 ? m1 : (lt2 ? a : m2);
        m1 = lt1 ? a : m1;
    }
    LDPC_HD T out(T v) const {
        const T a = num<T>::abs(v);
        const T mag = (a == m1) ? m2 : m1;
        const unsigned neg = par ^ ((v >= (T)0) ? 0u : 1u);      // sign(0) = +1 (math_utils.py:10); NaN -> "-"
        return neg ? -mag : mag;                                  // (-1) * 0 = -0.0 like the reference
    }
};

template <typename T, int DCMAX>
LDPC_HD void cn_msa(const T (&v)[DCMAX], int dc, T (&out)[DCMAX])
{
    MsaAcc<T> acc;
    acc.init();
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) acc.push(v[k]);
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) out[k] = acc.out(v[k]);
}

// ---------------------------------------------------------------------------------------------
// The same min-sum rule for float32 on the IEEE bit patterns (fewer issue slots: integer min/max instead
// of compare+select pairs).  For non-NaN inputs it is bit-identical to cn_msa<float>: magnitudes of
// non-negative floats order like their bit patterns, "v < 0" is "bits > 0x80000000" (so -0.0 counts as
// "+", math_utils.py:10,40), and the result's sign bit is par ^ own.  Used by the resident kernel, where
// the check-node arithmetic (not HBM) is the limiter.
// ---------------------------------------------------------------------------------------------
LDPC_HD uint32_t f32_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
LDPC_HD float bits_f32(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

template <int DCMAX>
LDPC_HD void cn_msa_bits(const float (&v)[DCMAX], int dc, float (&out)[DCMAX])
{
    uint32_t m1 = 0x7f800000u, m2 = 0x7f800000u, par = 0u;
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) {
            const uint32_t b = f32_bits(v[k]), a = b & 0x7fffffffu;
            par ^= (b > 0x80000000u) ? 0x80000000u : 0u;
            const uint32_t hi = a > m1 ? a : m1;
            m2 = m2 < hi ? m2 : hi;
            m1 = m1 < a ? m1 : a;
        }
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) {
            const uint32_t b = f32_bits(v[k]), a = b & 0x7fffffffu;
            const uint32_t mag = (a == m1) ? m2 : m1;
            const uint32_t sgn = par ^ ((b > 0x80000000u) ? 0x80000000u : 0u);
            out[k] = bits_f32(mag | sgn);
        }
}

// ---------------------------------------------------------------------------------------------
// Lean float32 min-sum for a check of EXACTLY DC edges (on-chip kernel, where the ALU pipe is the limiter):
//   |out_k| = min_{j != k} |v_j|  built from 3-input minima (FMNMX3 with free |.| operand modifiers),
//   sign    = xor of all sign bits ^ own sign bit, applied as a multiplication by +-1.0f on the FMA pipe
//             (the reference's literal `sign * mag`, bpa.py:102, so 0 * -1 = -0.0 comes out the same).
// Precondition: no input is NaN or -0.0 — then "v < 0" is the sign bit, and the result is bit-identical to
// cn_msa<float>.  The resident kernel guarantees it: v2c = marg - c2v can only be -0.0 if marg is, and
// marg = prior + (0.0 + c ...) never is once -0.0 priors are folded to +0.0 at load (same values everywhere:
// -0.0 and +0.0 priors are indistinguishable to every formula of bpa.py).
// ---------------------------------------------------------------------------------------------
LDPC_HD float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <int DC>
LDPC_HD void cn_msa_lean(const float (&v)[DC], float (&out)[DC])
{
    float a[DC], mag[DC];
#pragma unroll
    for (int k = 0; k < DC; ++k) a[k] = fabsf(v[k]);
    if (DC == 6) {
        const float mL = fmin3(a[0], a[1], a[2]), mR = fmin3(a[3], a[4], a[5]);
        mag[0] = fmin3(a[1], a[2], mR); mag[1] = fmin3(a[0], a[2], mR); mag[2] = fmin3(a[0], a[1], mR);
        mag[3] = fmin3(a[4], a[5], mL); mag[4] = fmin3(a[3], a[5], mL); mag[5] = fmin3(a[3], a[4], mL);
    } else {
        float pre[DC], suf[DC];
        pre[0] = INFINITY;
#pragma unroll
        for (int k = 1; k < DC; ++k) pre[k] = fminf(pre[k - 1], a[k - 1]);
        suf[DC - 1] = INFINITY;
#pragma unroll
        for (int k = DC - 2; k >= 0; --k) suf[k] = fminf(suf[k + 1], a[k + 1]);
#pragma unroll
        for (int k = 0; k < DC; ++k) mag[k] = fminf(pre[k], suf[k]);
    }
    uint32_t x = 0u;
#pragma unroll
    for (int k = 0; k < DC; ++k) x ^= f32_bits(v[k]);
    const uint32_t xs = (x & 0x80000000u) | 0x3f800000u;                    // +-1.0f carrying the check's parity
#pragma unroll
    for (int k = 0; k < DC; ++k) out[k] = mag[k] * bits_f32(xs ^ (f32_bits(v[k]) & 0x80000000u));
}

// ---------------------------------------------------------------------------------------------
// Sum-product check node, formula mirror (float64 verification mode).
//   t = tanh(v / 2); S = sum_k log|t_k| (ordered, from 0); P = (-1)^{#(t<0)} exp(S)
//   q = P / t_k;  out = 2 * (|q| == 1 ? inf * q : atanh(q))
// ---------------------------------------------------------------------------------------------
template <typename T> struct SpaRefAcc {
    T s;
    unsigned par;
    LDPC_HD void init() { s = (T)0; par = 0u; }
    LDPC_HD T push(T v) {                       // returns t = tanh(v/2) for the second pass
        const T t = (T)tanh((double)(v / (T)2));
        s = num<T>::add(s, (T)log((double)num<T>::abs(t)));
        par ^= (t < (T)0) ? 1u : 0u;
        return t;
    }
    LDPC_HD T prod() const { return (par ? (T)-1 : (T)1) * (T)exp((double)s); }
    static LDPC_HD T out(T P, T t) {
        const T q = P / t;                                              // 0/0 -> NaN (bpa.py:74 "TODO")
        const T r = (num<T>::abs(q) == (T)1) ? num<T>::inf() * q : (T)atanh((double)q);
        return (T)2 * r;
    }
};

template <int DCMAX>
LDPC_HD void cn_spa_ref(const double (&v)[DCMAX], int dc, double (&out)[DCMAX])
{
    SpaRefAcc<double> acc;
    acc.init();
    double t[DCMAX];
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) t[k] = acc.push(v[k]);
    const double P = acc.prod();
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) out[k] = SpaRefAcc<double>::out(P, t[k]);
}

// ---------------------------------------------------------------------------------------------
// Saturation of the float64 reference, emulated by the float32 rule below.  In float64 tanh(v/2) rounds to exactly 1
// once |v| > 55 ln 2 = 38.123 (2 e^-|v| < 2^-54), its log is then 0, and a check whose OTHER inputs are all saturated
// emits +-inf (|q| == 1, math_utils.py:57-58), which turns the variable's "total minus own" into inf - inf = NaN
// (bpa.py:37) and floods the frame.  With sat_llr = 38.123 a saturated input contributes u = e^-|v| = 0 here too, so
// the float32 decoder leaves the well-conditioned regime at the same point as the reference (its BER/WER curves
// depend on it: a flooded frame decodes to all-zero).  sat_llr = +inf switches the emulation off (robust decoder).
// ---------------------------------------------------------------------------------------------
constexpr float kSpaSatLlr = 38.1230f;

// ---------------------------------------------------------------------------------------------
// Sum-product check node in float32, hyperbolic-pair form (the kernels' float32 SPA rule).
// With u_j = e^-|v_j| the magnitude of bpa.py:71-75 is
//   |out_k| = 2 atanh( prod_{j != k} tanh(|v_j| / 2) ) = log( S_k / C_k ),
//   S_k = A + B, C_k = A - B,  A = prod_{j != k} (1 + u_j),  B = prod_{j != k} (1 - u_j).
// (S, C) of a set grows by one element with two fused multiply-adds and never subtracts:
//   (S, C) -> (S + u C, C + u S)        [the tanh addition theorem on C / S]
// so both stay accurate to a few ulp whatever the inputs (no 1 - P, no total-minus-own), S >= C holds after
// rounding (both are monotone in the same exact quantities), and an input v_j == 0 (u_j = 1) makes S == C bit for
// bit from then on, i.e. |out| == 0 exactly like the reference's 0 * ... = 0 (bpa.py:74).
// Cost for a degree-6 check: 6 ex2 + 14 pair steps (28 FFMA) + 12 lg2 — 18 MUFU and ~100 instructions, against 36 MUFU
// and ~280 instructions for the phi-domain form phi(sum phi) it replaced (kept as a cross-check in tests/host_emu).  v_k == 0 gives NaN on its own edge (0/0), a check whose OTHER inputs are all
// saturated (u = 0 beyond sat_llr, above) gives C == 0 -> +-inf: the reference's degenerate cases.
// Measured against the float64 formula on the float32-rounded golden states: <= 1e-6 * max(1,|ref|) for |ref| < 20.
// ---------------------------------------------------------------------------------------------
LDPC_HD float spa_ex2(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f(x);
#endif
}
LDPC_HD float spa_lg2(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return log2f(x);
#endif
}
struct SpaPair { float s, c; };
LDPC_HD SpaPair sc_one(float u) { SpaPair r; r.s = 1.0f; r.c = u; return r; }
LDPC_HD SpaPair sc_step(SpaPair a, float u)
{
    SpaPair r;
    r.s = fmaf(u, a.c, a.s);
    r.c = fmaf(u, a.s, a.c);
    return r;
}
// Union of two disjoint sets.  Products and sums are rounded separately so that a side with s == c yields s == c.
LDPC_HD SpaPair sc_join(SpaPair a, SpaPair b)
{
    SpaPair r;
    r.s = num<float>::add(num<float>::mul(a.s, b.s), num<float>::mul(a.c, b.c));
    r.c = num<float>::add(num<float>::mul(a.s, b.c), num<float>::mul(a.c, b.s));
    return r;
}
// xs = (xor of all sign bits) | bits(ln 2): the check's parity riding on the scale factor, applied by one multiplication
// like the reference's literal sign * magnitude.  Sign BITS instead of (v < 0) only changes the sign of zero and
// NaN results (an input of -0.0 makes every other output a zero, a NaN input makes every output NaN).
template <bool CLAMP>
LDPC_HD float sc_out(SpaPair a, float v, uint32_t xs)
{
    float d = spa_lg2(a.s) - spa_lg2(a.c);                         // >= 0: S >= C after rounding when built from steps only
    if (CLAMP) d = (d < 0.0f) ? 0.0f : d;                          // sc_join may round S one ulp below C; NaN stays NaN
    return d * bits_f32(xs ^ (f32_bits(v) & 0x80000000u));          // v == 0: patched to NaN by the caller
}

// The two degenerate inputs — a saturated one (u = 0) and an exact zero (NaN on its own edge) — are found with one
// min / max over the check (3-input FMNMX) and patched on rarely taken branches, instead of a compare + select on
// every edge: the common path is ex2 -> pair steps -> lg2 and nothing else.  A NaN input needs no help: it makes
// every u, S, C and output of the check NaN by itself, like the reference's tanh / log / exp chain.
template <int DCMAX>
LDPC_HD void cn_spa_sc(const float (&v)[DCMAX], int dc, float (&out)[DCMAX], float sat_llr = kSpaSatLlr)
{
    float u[DCMAX];
    uint32_t x = 0u;
    float lo = INFINITY, hi = 0.0f;
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) {
            const float av = fabsf(v[k]);
            u[k] = spa_ex2(av * -1.44269504088896f);
            x ^= f32_bits(v[k]);
            lo = fminf(lo, av);
            hi = fmaxf(hi, av);
        }
    if (hi > sat_llr) {
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc && fabsf(v[k]) > sat_llr) u[k] = 0.0f;
    }
    const bool has_zero = !(lo > 0.0f);
    const uint32_t xs = (x & 0x80000000u) | 0x3f317218u;            // +-ln 2
    if (dc <= 6) {
        // Up to six edges: the all-but-one sets of {0..5} from the two halves, single-element steps only (14 of them).
        // A check with fewer edges is padded with u = 0, the empty element — sc_one(0) is the empty set (1, 0) and a step
        // with u = 0 is the exact identity — so every caller runs the same operations on the same check whatever its
        // DCMAX and however it pads (the on-chip kernel pads with +inf inputs, which saturate to u = 0 above).
        float w[6], vv[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const bool real = (k < DCMAX) && (k < dc);
            w[k] = real ? u[k < DCMAX ? k : 0] : 0.0f;
            vv[k] = real ? v[k < DCMAX ? k : 0] : 0.0f;
        }
        const SpaPair L = sc_step(sc_step(sc_one(w[0]), w[1]), w[2]);
        const SpaPair R = sc_step(sc_step(sc_one(w[3]), w[4]), w[5]);
        const SpaPair R0 = sc_step(R, w[0]), R1 = sc_step(R, w[1]);
        const SpaPair L3 = sc_step(L, w[3]), L4 = sc_step(L, w[4]);
        float o6[6];
        o6[0] = sc_out<false>(sc_step(R1, w[2]), vv[0], xs);
        o6[1] = sc_out<false>(sc_step(R0, w[2]), vv[1], xs);
        o6[2] = sc_out<false>(sc_step(R0, w[1]), vv[2], xs);
        o6[3] = sc_out<false>(sc_step(L4, w[5]), vv[3], xs);
        o6[4] = sc_out<false>(sc_step(L3, w[5]), vv[4], xs);
        o6[5] = sc_out<false>(sc_step(L3, w[4]), vv[5], xs);
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k < DCMAX && k < dc) out[k < DCMAX ? k : 0] = o6[k];
        if (has_zero) {
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (k < DCMAX && k < dc && vv[k] == 0.0f) out[k < DCMAX ? k : 0] = NAN;
        }
        return;
    }
    // general degree: prefix sets by steps, suffix sets by steps, joined
    SpaPair pre[DCMAX];
    SpaPair run = sc_one(0.0f);                                     // the empty set: (1, 0)
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) {
            pre[k] = run;
            run = sc_step(run, u[k]);
        }
    SpaPair suf = sc_one(0.0f);
#pragma unroll
    for (int k = DCMAX - 1; k >= 0; --k)
        if (k < dc) {
            out[k] = sc_out<true>(sc_join(pre[k], suf), v[k], xs);
            suf = sc_step(suf, u[k]);
        }
    if (has_zero) {
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc && v[k] == 0.0f) out[k] = NAN;
    }
}

// ---------------------------------------------------------------------------------------------
// Variable node: s = ((0 + c0) + c1) + ... in ascending edge order; marg = prior + s;
// out_k = marg - c_k ("total minus own", bpa.py:37).  Returns the marginal BEFORE the NaN scrub;
// the hard decision is (marg < 0), which is false for NaN exactly like the scrubbed value (bpa.py:38,62).
// ---------------------------------------------------------------------------------------------
template <typename T, int DVMAX>
LDPC_HD T vn_update(T prior, const T (&c)[DVMAX], int dv, T (&out)[DVMAX])
{
    T s = (T)0;
#pragma unroll
    for (int k = 0; k < DVMAX; ++k)
        if (k < dv) s = num<T>::add(s, c[k]);
    const T marg = num<T>::add(prior, s);
#pragma unroll
    for (int k = 0; k < DVMAX; ++k)
        if (k < dv) out[k] = num<T>::sub(marg, c[k]);
    return marg;
}

// ---------------------------------------------------------------------------------------------
// BIAWGN prior in float32, rounded exactly like the reference's float64 expression cast to float32:
//   priors = (-2 y) / noise_var  (biawgn.py:28), then .astype(float32).
// The float64 product with the reciprocal is within 2 ulp of the quotient, so it rounds to the same float32 unless it
// lies within 4 ulp of a float32 rounding boundary (29 mantissa bits below float32 precision == 0x10000000 +- 4,
// probability 2^-26) or outside the normal float32 range — only then is the division evaluated.
// ---------------------------------------------------------------------------------------------
LDPC_HD uint32_t f64_lo_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(d);
#else
    union { double d; uint64_t u; } c; c.d = d; return (uint32_t)c.u;
#endif
}
LDPC_HD bool llr_biawgn_fast_ok(double pr)
{
    const uint32_t lo = f64_lo_bits(pr) & 0x1fffffffu;
    const double ap = fabs(pr);
    return (lo - 0x0ffffffcu) > 8u && ap >= 2e-38 && ap < 3e38;
}
LDPC_HD float llr_biawgn_f32(double y, double noise_var, double inv_noise_var)
{
    const double t = -2.0 * y;                                           // exact
    const double pr = t * inv_noise_var;
    if (llr_biawgn_fast_ok(pr)) return (float)pr;
    return (float)(t / noise_var);
}

// ---------------------------------------------------------------------------------------------
// BEC, 32 frames per word.  A message in {-1,0,+1} is two bit planes: nz (|msg|) and pos (msg > 0),
// pos is a subset of nz.  Literal restatement of bec.py:100-119 in bit-sliced integer arithmetic.
// ---------------------------------------------------------------------------------------------
// U = the machine word(s) a thread owns: uint32_t (32 frames) or a pack of four (stream_bec.cuh, 128 frames); only
// the bitwise operators are used, so any type providing them works.
template <typename U> struct BecCnAccT {   // per check: erasure count saturating at 2, parity of the +1 votes
    U any, two, par;
    LDPC_HD void init() { any = U(); two = U(); par = U(); }
    LDPC_HD void push(U nz, U pos) {
        const U er = ~nz;                        // 1 - |v2c|  (bec.py:100)
        two = two | (any & er);
        any = any | er;
        par = par ^ pos;                         // (v2c > 0) summed mod 2 (bec.py:110,112)
    }
    // sums == 0: echo v2c; sums > 1: 0; sums == 1: only the erased edge, value 2*(incoming%2)-1  (bec.py:105,112)
    LDPC_HD void out(U nz, U pos, U &onz, U &opos) const {
        const U zero = ~any, one = any & ~two, er = ~nz;
        onz = (zero & nz) | (one & er);
        opos = (zero & pos) | (one & er & par);
    }
};
using BecCnAcc = BecCnAccT<uint32_t>;

// Bit-sliced two's-complement integer of NB bits per frame (range covers |prior + sum of dv votes|).
template <int NB, typename U = uint32_t> struct BsInt {
    U b[NB];
    LDPC_HD void set_ternary(U nz, U pos) {      // +1 = 0..01, -1 = 1..11, 0 = 0
        const U neg = nz & ~pos;
        b[0] = nz;
#pragma unroll
        for (int i = 1; i < NB; ++i) b[i] = neg;
    }
    LDPC_HD void add_ternary(U nz, U pos) {
        const U neg = nz & ~pos;
        U carry = b[0] & nz;
        b[0] = b[0] ^ nz;
#pragma unroll
        for (int i = 1; i < NB; ++i) {
            const U x = b[i];
            b[i] = x ^ neg ^ carry;
            carry = (x & neg) | (carry & (x ^ neg));
        }
    }
    LDPC_HD void sub_ternary(U nz, U pos) { add_ternary(nz, nz & ~pos); }   // -(msg): swap +1/-1
    LDPC_HD void sign(U &nz, U &pos) const {                                // np.sign (bec.py:116,119)
        U any = U();
#pragma unroll
        for (int i = 0; i < NB; ++i) any = any | b[i];
        nz = any;
        pos = any & ~b[NB - 1];
    }
};

// Degree-3 variable node on bit planes WITHOUT integer arithmetic (the on-chip erasure kernel, resident_bec.cuh).
// Inputs: index 0 = prior, 1..3 = the three c2v messages, each a ternary value (nz, pos subset of nz).
// bec.py:115-119:  marginal = prior + sum c2v;  v2c_e = sign(marginal - c2v_e);  x_new = symbols[sign(marginal)].
// marginal - c2v_e is the sum of the OTHER three inputs, and the sign of a sum of three ternary values a, b, c is
//   > 0  <=>  #(+1) > #(-1)  <=>  maj(p_a, p_b, p_c) | (any p & no n)        (p = "is +1", n = "is -1")
// so every output is a handful of 3-input boolean functions (one LOP3 each) instead of ripple adders: ~40 logic
// instructions per variable and word against ~70 with BsInt<4>.  The marginal's sign is recovered from the
// leave-one-out sum of edge 1 (S1 in [-3, 3], known up to {<= -2, -1, 0, 1, >= 2}) plus the left-out message.
// Exhaustively checked against BsInt on all 3^4 inputs (tests/test_host_emu.py).
template <typename U>
LDPC_HD void bec_vn3_pn(const U (&pos)[4], const U (&n)[4], U (&onz)[3], U (&opos)[3], U &mnz, U &mpos)
{
    U ge2 = U(), le2 = U();
#pragma unroll
    for (int e = 1; e <= 3; ++e) {
        const int a = 0, b = (e == 1) ? 2 : 1, c = (e == 3) ? 2 : 3;      // the other three inputs
        const U mp = (pos[a] & pos[b]) | (pos[c] & (pos[a] | pos[b]));
        const U mn = (n[a] & n[b]) | (n[c] & (n[a] | n[b]));
        const U op = pos[a] | pos[b] | pos[c], on = n[a] | n[b] | n[c];
        const U gt = mp | (op & ~on), lt = mn | (on & ~op);
        onz[e - 1] = gt | lt;
        opos[e - 1] = gt;
        if (e == 1) { ge2 = mp & ~on; le2 = mn & ~op; }
    }
    const U z1 = ~onz[0];
    const U ps = ge2 | (opos[0] & ~n[1]) | (z1 & pos[1]);
    const U ng = le2 | ((onz[0] & ~opos[0]) & ~pos[1]) | (z1 & n[1]);
    mnz = ps | ng;
    mpos = ps;
}
template <typename U>
LDPC_HD void bec_vn3(const U (&nz)[4], const U (&pos)[4], U (&onz)[3], U (&opos)[3], U &mnz, U &mpos)
{
    U n[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) n[i] = nz[i] & ~pos[i];
    bec_vn3_pn<U>(pos, n, onz, opos, mnz, mpos);
}

// Variable node WITHOUT CONFLICTING VOTES (no frame of the word has both a +1 and a -1 among the prior and the incoming
// messages of this variable) — always the case for symbols that come from an erasure channel applied to a codeword: every
// vote is the transmitted bit.  Then the sign of any sum of the inputs is "+ if some input is +, - if some input is -",
// and the literal rule of bec.py:115-119 collapses to ORs: v2c_e = OR of the OTHER inputs' planes, marginal = OR of all.
// bec_conflict() is the test; the kernel takes this path when it is zero and the literal rule otherwise, so arbitrary
// (inconsistent) inputs still reproduce the reference bit for bit.  18 logic instructions per word instead of 33.
template <typename U> LDPC_HD U bec_conflict(const U (&pos)[4], const U (&n)[4])
{
    return (pos[0] | pos[1] | pos[2] | pos[3]) & (n[0] | n[1] | n[2] | n[3]);
}
template <typename U>
LDPC_HD void bec_vn3_or(const U (&pos)[4], const U (&n)[4], U (&onz)[3], U (&opos)[3], U &mnz, U &mpos)
{
#pragma unroll
    for (int e = 1; e <= 3; ++e) {
        const int b = (e == 1) ? 2 : 1, c = (e == 3) ? 2 : 3;
        const U gt = pos[0] | pos[b] | pos[c], lt = n[0] | n[b] | n[c];
        onz[e - 1] = gt | lt;
        opos[e - 1] = gt;
    }
    mpos = pos[0] | pos[1] | pos[2] | pos[3];
    mnz = mpos | n[0] | n[1] | n[2] | n[3];
}

// The same for any degree: "some OTHER input is +" = (two or more inputs are +) | (some input is + and this one is not).
template <typename U> struct BecVnOr {
    U pa, p2, na, n2;
    LDPC_HD void init(U nz, U pos) { pa = pos; na = nz & ~pos; p2 = U(); n2 = U(); }        // the prior
    LDPC_HD void push(U nz, U pos) {
        const U n = nz & ~pos;
        p2 = p2 | (pa & pos); pa = pa | pos;
        n2 = n2 | (na & n);   na = na | n;
    }
    LDPC_HD U conflict() const { return pa & na; }
    LDPC_HD void out(U nz, U pos, U &onz, U &opos) const {
        const U n = nz & ~pos;
        const U gt = p2 | (pa & ~pos), lt = n2 | (na & ~n);
        onz = gt | lt;
        opos = gt;
    }
    LDPC_HD void marg(U &mnz, U &mpos) const { mnz = pa | na; mpos = pa; }
};

// Check node of degree 6 on bit planes, erasure count and parity as reduction TREES (three-input functions, one LOP3
// each) instead of the sequential accumulator BecCnAccT::push: 9 logic instructions per word instead of 18.
// Same outputs as BecCnAccT (bec.py:100-112).
template <typename U> struct BecCn6 {
    U zero, one, onepar;
    LDPC_HD void reduce(const U (&nz)[6], const U (&pos)[6]) {
        const U a1 = ~(nz[0] & nz[1] & nz[2]), a2 = ~(nz[3] & nz[4] & nz[5]);                      // some erasure in the triple
        const U t1 = (~nz[0] & ~nz[1]) | (~nz[2] & (~nz[0] | ~nz[1])), t2 = (~nz[3] & ~nz[4]) | (~nz[5] & (~nz[3] | ~nz[4]));   // two or more
        const U any = a1 | a2, two = t1 | t2 | (a1 & a2);
        const U par = (pos[0] ^ pos[1] ^ pos[2]) ^ (pos[3] ^ pos[4] ^ pos[5]);
        zero = ~any;
        one = any & ~two;
        onepar = one & par;
    }
    LDPC_HD void out(U nz, U pos, U &onz, U &opos) const {
        onz = (zero & nz) | (one & ~nz);
        opos = (zero & pos) | (onepar & ~nz);
    }
};

}  // namespace ldpc
