// ldpc_math.cuh — per-frame node arithmetic shared by every kernel.
//
// These are the order-of-operations contracts of the reference restated for
// one frame held in registers:
//   cn_msa      bpa.MSA.decode_   /root/reference/src/bpa.py:86-102, math_utils.py:10,38-43,78-94
//   cn_spa_ref  bpa.SPA.decode_   /root/reference/src/bpa.py:71-75,   math_utils.py:38-60   (formula mirror)
//   cn_spa_phi  same function, evaluated in the phi domain for float32 (the mirror is unusable in fp32)
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
#else
    static LDPC_HD float add(float a, float b) { volatile float r = a + b; return r; }
    static LDPC_HD float sub(float a, float b) { volatile float r = a - b; return r; }
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
// Sum-product check node in float32, phi domain:
//   |out_k| = phi( sum_{j != k} phi(|v_j|) ),  phi(x) = -log tanh(x/2) = log((1+e^-x)/(1-e^-x))
//   sign    = (-1)^{#(v<0)} * sign(v_k)
// phi is evaluated without cancellation: series for 1-e^-x when x is tiny, the
// atanh series 2u(1+u^2/3+u^4/5), u = e^-x, when x > 3.  The sum over the OTHER
// edges is built from prefix/suffix sums (total - own cancels catastrophically).
// Degenerate inputs follow the reference: v_k == 0 gives NaN on that edge
// (0/0, bpa.py:74) and 0 on the others; all others saturated gives +-inf.
// Saturation is the reference's, not float32's: in float64 tanh(v/2) rounds to exactly 1 once
// |v| > 55 ln 2 = 38.123 (2 e^-|v| < 2^-54), its log is then 0, and a check whose OTHER inputs are all
// saturated emits +-inf (|q| == 1, math_utils.py:57-58) which turns the variable's "total minus own"
// into inf - inf = NaN (bpa.py:37) and floods the frame.  With sat_llr = 38.123 a saturated input
// contributes phi = 0 here too, so the float32 decoder leaves the well-conditioned regime at the same
// point as the reference (its BER/WER curves depend on it: a flooded frame decodes to all-zero).
// sat_llr = +inf switches the emulation off (numerically robust decoder).
// Measured against the float64 reference formula: <= 6e-7 * max(1,|ref|) for |ref| < 20.
// ---------------------------------------------------------------------------------------------
LDPC_HD float phi_f32(float x)
{
#if defined(__CUDA_ARCH__)
    const float u = __expf(-x);
#else
    const float u = expf(-x);
#endif
    const float series = x * (1.0f - x * 0.5f * (1.0f - x * (1.0f / 3.0f) * (1.0f - x * 0.25f)));
    const float den = (x < 0.05f) ? series : (1.0f - u);
#if defined(__CUDA_ARCH__)
    const float big = __logf(__fdividef(1.0f + u, den));
#else
    const float big = logf((1.0f + u) / den);
#endif
    const float u2 = u * u;
    const float small = 2.0f * u * (1.0f + u2 * ((1.0f / 3.0f) + u2 * 0.2f));
    return (x > 3.0f) ? small : big;
}

constexpr float kSpaSatLlr = 38.1230f;       // 55 ln 2: float64 tanh(v/2) == 1 beyond this |v|

template <int DCMAX>
LDPC_HD void cn_spa_phi(const float (&v)[DCMAX], int dc, float (&out)[DCMAX], float sat_llr = kSpaSatLlr)
{
    float a[DCMAX], pre[DCMAX];
    unsigned par = 0u;
    float run = 0.0f;
#pragma unroll
    for (int k = 0; k < DCMAX; ++k)
        if (k < dc) {
            const float av = fabsf(v[k]);
            a[k] = (av > sat_llr) ? 0.0f : phi_f32(av);
            par ^= (v[k] < 0.0f) ? 1u : 0u;
            pre[k] = run;
            run += a[k];
        }
    float suf = 0.0f;
#pragma unroll
    for (int k = DCMAX - 1; k >= 0; --k)
        if (k < dc) {
            const float mag = phi_f32(pre[k] + suf);
            suf += a[k];
            const unsigned neg = par ^ ((v[k] < 0.0f) ? 1u : 0u);
            const float r = neg ? -mag : mag;
            out[k] = (v[k] == 0.0f) ? NAN : r;
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
struct BecCnAcc {                 // per check: erasure count saturating at 2, parity of the +1 votes
    uint32_t any, two, par;
    LDPC_HD void init() { any = 0u; two = 0u; par = 0u; }
    LDPC_HD void push(uint32_t nz, uint32_t pos) {
        const uint32_t er = ~nz;                 // 1 - |v2c|  (bec.py:100)
        two |= any & er;
        any |= er;
        par ^= pos;                              // (v2c > 0) summed mod 2 (bec.py:110,112)
    }
    // sums == 0: echo v2c; sums > 1: 0; sums == 1: only the erased edge, value 2*(incoming%2)-1  (bec.py:105,112)
    LDPC_HD void out(uint32_t nz, uint32_t pos, uint32_t &onz, uint32_t &opos) const {
        const uint32_t zero = ~any, one = any & ~two, er = ~nz;
        onz = (zero & nz) | (one & er);
        opos = (zero & pos) | (one & er & par);
    }
};

// Bit-sliced two's-complement integer of NB bits per frame (range covers |prior + sum of dv votes|).
template <int NB> struct BsInt {
    uint32_t b[NB];
    LDPC_HD void set_ternary(uint32_t nz, uint32_t pos) {   // +1 = 0..01, -1 = 1..11, 0 = 0
        const uint32_t neg = nz & ~pos;
        b[0] = nz;
#pragma unroll
        for (int i = 1; i < NB; ++i) b[i] = neg;
    }
    LDPC_HD void add_ternary(uint32_t nz, uint32_t pos) {
        const uint32_t neg = nz & ~pos;
        uint32_t carry = b[0] & nz;
        b[0] ^= nz;
#pragma unroll
        for (int i = 1; i < NB; ++i) {
            const uint32_t x = b[i];
            b[i] = x ^ neg ^ carry;
            carry = (x & neg) | (carry & (x ^ neg));
        }
    }
    LDPC_HD void sub_ternary(uint32_t nz, uint32_t pos) { add_ternary(nz, nz & ~pos); }   // -(msg): swap +1/-1
    LDPC_HD void sign(uint32_t &nz, uint32_t &pos) const {                                 // np.sign (bec.py:116,119)
        uint32_t any = 0u;
#pragma unroll
        for (int i = 0; i < NB; ++i) any |= b[i];
        nz = any;
        pos = any & ~b[NB - 1];
    }
};

}  // namespace ldpc
