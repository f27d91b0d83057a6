// spa_phi_alt.h — TEST-ONLY second float32 formulation of the sum-product check node (phi domain):
//   |out_k| = phi( sum_{j != k} phi(|v_j|) ),  phi(x) = -log tanh(x/2) = log((1+e^-x)/(1-e^-x))
// It was the kernels' rule before the hyperbolic-pair form (csrc/ldpc_math.cuh cn_spa_sc) replaced it at a third of
// the instructions; it stays here as an independent cross-check: two unrelated float32 evaluations of bpa.py:71-75
// must both agree with the float64 formula.  phi is evaluated without cancellation: series for 1-e^-x when x is
// tiny, the atanh series 2u(1+u^2/3+u^4/5), u = e^-x, when x > 3; sums over the OTHER edges are prefix/suffix sums.
#pragma once
#include "../../ldpc_decoders_b200/csrc/ldpc_math.cuh"

namespace ldpc {

inline float phi_f32(float x)
{
    const float u = expf(-x);
    const float series = x * (1.0f - x * 0.5f * (1.0f - x * (1.0f / 3.0f) * (1.0f - x * 0.25f)));
    const float den = (x < 0.05f) ? series : (1.0f - u);
    const float big = logf((1.0f + u) / den);
    const float u2 = u * u;
    const float small = 2.0f * u * (1.0f + u2 * ((1.0f / 3.0f) + u2 * 0.2f));
    return (x > 3.0f) ? small : big;
}

template <int DCMAX>
inline void cn_spa_phi(const float (&v)[DCMAX], int dc, float (&out)[DCMAX], float sat_llr = kSpaSatLlr)
{
    float a[DCMAX], pre[DCMAX];
    unsigned par = 0u;
    float run = 0.0f;
    for (int k = 0; k < dc; ++k) {
        const float av = fabsf(v[k]);
        a[k] = (av > sat_llr) ? 0.0f : phi_f32(av);
        par ^= (v[k] < 0.0f) ? 1u : 0u;
        pre[k] = run;
        run += a[k];
    }
    float suf = 0.0f;
    for (int k = dc - 1; k >= 0; --k) {
        const float mag = phi_f32(pre[k] + suf);
        suf += a[k];
        const unsigned neg = par ^ ((v[k] < 0.0f) ? 1u : 0u);
        const float r = neg ? -mag : mag;
        out[k] = (v[k] == 0.0f) ? NAN : r;
    }
}

}  // namespace ldpc
