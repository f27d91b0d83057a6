// tests/host_emu/emu.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's node arithmetic (ldpc_decoders_b200/csrc/ldpc_math.cuh, the
// __host__ __device__ functions every CUDA kernel calls) with g++ and drives it with plain
// loops, so that tests/test_host_emu.py can compare it with the oracle on the CPU box before any
// GPU time is spent.  It is not a decoder anybody ships or times: the package never loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../ldpc_decoders_b200/csrc/ldpc_math.cuh"
#include "../../ldpc_decoders_b200/csrc/res_layout.h"
#include "../../ldpc_decoders_b200/csrc/channel_gen.cuh"
#include "spa_phi_alt.h"

using namespace ldpc;

namespace {
struct G { int n, m, E; const int32_t *chk_ptr, *edge_var, *var_ptr, *var_edges; };
constexpr int DMAX = 32;

template <typename T, int ALGO> void cn(const G &g, const T *v2c, T *c2v);
template <typename T> void cn_msa_all(const G &g, const T *v2c, T *c2v)
{
    for (int c = 0; c < g.m; ++c) {
        const int e0 = g.chk_ptr[c], dc = g.chk_ptr[c + 1] - e0;
        T a[DMAX], o[DMAX];
        for (int k = 0; k < dc; ++k) a[k] = v2c[e0 + k];
        cn_msa<T, DMAX>(a, dc, o);
        for (int k = 0; k < dc; ++k) c2v[e0 + k] = o[k];
    }
}
void cn_msa_bits_all(const G &g, const float *v2c, float *c2v)
{
    for (int c = 0; c < g.m; ++c) {
        const int e0 = g.chk_ptr[c], dc = g.chk_ptr[c + 1] - e0;
        float a[DMAX], o[DMAX];
        for (int k = 0; k < dc; ++k) a[k] = v2c[e0 + k];
        cn_msa_bits<DMAX>(a, dc, o);
        for (int k = 0; k < dc; ++k) c2v[e0 + k] = o[k];
    }
}
void cn_spa_ref_all(const G &g, const double *v2c, double *c2v)
{
    for (int c = 0; c < g.m; ++c) {
        const int e0 = g.chk_ptr[c], dc = g.chk_ptr[c + 1] - e0;
        double a[DMAX], o[DMAX];
        for (int k = 0; k < dc; ++k) a[k] = v2c[e0 + k];
        cn_spa_ref<DMAX>(a, dc, o);
        for (int k = 0; k < dc; ++k) c2v[e0 + k] = o[k];
    }
}
void cn_spa_phi_all(const G &g, const float *v2c, float *c2v)
{
    for (int c = 0; c < g.m; ++c) {
        const int e0 = g.chk_ptr[c], dc = g.chk_ptr[c + 1] - e0;
        float a[DMAX], o[DMAX];
        for (int k = 0; k < dc; ++k) a[k] = v2c[e0 + k];
        cn_spa_phi<DMAX>(a, dc, o);
        for (int k = 0; k < dc; ++k) c2v[e0 + k] = o[k];
    }
}
int g_spa_rule = 1;            // 1: hyperbolic-pair rule (the kernels'), 0: phi-domain form
// the float32 sum-product rule the kernels use (hyperbolic-pair form)
void cn_spa_sc_all(const G &g, const float *v2c, float *c2v, float sat = kSpaSatLlr)
{
    for (int c = 0; c < g.m; ++c) {
        const int e0 = g.chk_ptr[c], dc = g.chk_ptr[c + 1] - e0;
        float a[DMAX], o[DMAX];
        for (int k = 0; k < dc; ++k) a[k] = v2c[e0 + k];
        cn_spa_sc<DMAX>(a, dc, o, sat);
        for (int k = 0; k < dc; ++k) c2v[e0 + k] = o[k];
    }
}

template <typename T>
void decode_bp(const G &g, int algo, const T *prior, const uint8_t *y_hard, int max_iter,
               uint8_t *x_hat, int32_t *iters, T *marg_out)
{
    std::vector<T> msg((size_t)g.E), tmp((size_t)g.E), marg(prior, prior + g.n);
    for (int e = 0; e < g.E; ++e) msg[e] = prior[g.edge_var[e]];
    bool have_x = y_hard != nullptr;
    if (have_x) memcpy(x_hat, y_hard, (size_t)g.n); else memset(x_hat, 0, (size_t)g.n);
    int it = 0;
    for (; it < max_iter; ++it) {
        if (have_x) {
            bool unsat = false;
            for (int c = 0; c < g.m && !unsat; ++c) {
                unsigned p = 0;
                for (int e = g.chk_ptr[c]; e < g.chk_ptr[c + 1]; ++e) p ^= x_hat[g.edge_var[e]];
                unsat = p & 1u;
            }
            if (!unsat) break;
        }
        if (algo == 0) cn_msa_all<T>(g, msg.data(), tmp.data());
        else if (algo == 2) cn_msa_bits_all(g, (const float *)msg.data(), (float *)tmp.data());
        else if (sizeof(T) == 8) cn_spa_ref_all(g, (const double *)msg.data(), (double *)tmp.data());
        else if (g_spa_rule) cn_spa_sc_all(g, (const float *)msg.data(), (float *)tmp.data());
        else cn_spa_phi_all(g, (const float *)msg.data(), (float *)tmp.data());
        for (int v = 0; v < g.n; ++v) {
            const int p0 = g.var_ptr[v], dv = g.var_ptr[v + 1] - p0;
            T a[DMAX], o[DMAX];
            for (int k = 0; k < dv; ++k) a[k] = tmp[g.var_edges[p0 + k]];
            const T mg = vn_update<T, DMAX>(prior[v], a, dv, o);
            for (int k = 0; k < dv; ++k) msg[g.var_edges[p0 + k]] = o[k];
            marg[v] = mg;
            x_hat[v] = (uint8_t)(mg < (T)0);
        }
        have_x = true;
    }
    *iters = it;
    if (marg_out) memcpy(marg_out, marg.data(), sizeof(T) * (size_t)g.n);
}
}  // namespace

extern "C" {

// Conflict-free variable rules (bec_vn3_or, BecVnOr) against the literal ones, on lanes WITHOUT conflicting votes:
// nz/pos [1 + d] words in (index 0 = prior); out = (onz, opos) per edge then (mnz, mpos); returns the conflict word.
uint32_t emu_bec_vn_or(int d, const uint32_t *nz, const uint32_t *pos, uint32_t *fast3, uint32_t *fastg, uint32_t *ref)
{
    ldpc::BsInt<5, uint32_t> acc;
    acc.set_ternary(nz[0], pos[0]);
    for (int k = 1; k <= d; ++k) acc.add_ternary(nz[k], pos[k]);
    for (int k = 1; k <= d; ++k) {
        ldpc::BsInt<5, uint32_t> t = acc;
        t.sub_ternary(nz[k], pos[k]);
        t.sign(ref[2 * (k - 1)], ref[2 * (k - 1) + 1]);
    }
    acc.sign(ref[2 * d], ref[2 * d + 1]);
    ldpc::BecVnOr<uint32_t> g;
    g.init(nz[0], pos[0]);
    for (int k = 1; k <= d; ++k) g.push(nz[k], pos[k]);
    for (int k = 1; k <= d; ++k) g.out(nz[k], pos[k], fastg[2 * (k - 1)], fastg[2 * (k - 1) + 1]);
    g.marg(fastg[2 * d], fastg[2 * d + 1]);
    uint32_t conflict = g.conflict();
    if (d == 3) {
        uint32_t p4[4], n4[4], onz[3], opos[3], mnz, mpos;
        for (int i = 0; i < 4; ++i) { p4[i] = pos[i]; n4[i] = nz[i] & ~pos[i]; }
        ldpc::bec_vn3_or<uint32_t>(p4, n4, onz, opos, mnz, mpos);
        for (int k = 0; k < 3; ++k) { fast3[2 * k] = onz[k]; fast3[2 * k + 1] = opos[k]; }
        fast3[6] = mnz; fast3[7] = mpos;
        if (ldpc::bec_conflict<uint32_t>(p4, n4) != conflict) return 0xdeadbeefu;
    }
    return conflict;
}


// BecCn6 (tree reduction) against BecCnAccT (sequential), one word: nz/pos [6] in, out [12] = (onz, opos) per edge.
void emu_bec_cn6(const uint32_t *nz, const uint32_t *pos, uint32_t *fast, uint32_t *ref)
{
    uint32_t a[6], b[6];
    for (int k = 0; k < 6; ++k) { a[k] = nz[k]; b[k] = pos[k]; }
    ldpc::BecCn6<uint32_t> t;
    t.reduce(a, b);
    ldpc::BecCnAccT<uint32_t> acc;
    acc.init();
    for (int k = 0; k < 6; ++k) acc.push(a[k], b[k]);
    for (int k = 0; k < 6; ++k) {
        t.out(a[k], b[k], fast[2 * k], fast[2 * k + 1]);
        acc.out(a[k], b[k], ref[2 * k], ref[2 * k + 1]);
    }
}


// bec_vn3 (degree-3 variable node as boolean functions) against the bit-sliced integer form, on caller-supplied planes:
// nz/pos [4] words in, out[8] = {onz0, opos0, onz1, opos1, onz2, opos2, mnz, mpos} for both formulations.
void emu_bec_vn3(const uint32_t *nz, const uint32_t *pos, uint32_t *fast, uint32_t *ref)
{
    uint32_t a[4] = {nz[0], nz[1], nz[2], nz[3]}, b[4] = {pos[0], pos[1], pos[2], pos[3]};
    uint32_t onz[3], opos[3], mnz, mpos;
    ldpc::bec_vn3<uint32_t>(a, b, onz, opos, mnz, mpos);
    for (int k = 0; k < 3; ++k) { fast[2 * k] = onz[k]; fast[2 * k + 1] = opos[k]; }
    fast[6] = mnz; fast[7] = mpos;
    ldpc::BsInt<5, uint32_t> acc;
    acc.set_ternary(a[0], b[0]);
    for (int k = 1; k < 4; ++k) acc.add_ternary(a[k], b[k]);
    for (int k = 1; k < 4; ++k) {
        ldpc::BsInt<5, uint32_t> t = acc;
        t.sub_ternary(a[k], b[k]);
        t.sign(ref[2 * (k - 1)], ref[2 * (k - 1) + 1]);
    }
    acc.sign(ref[6], ref[7]);
}


int emu_bp_f64(int algo, int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
               int B, const double *priors, const uint8_t *y_hard, int max_iter, uint8_t *x_hat, int32_t *iters, double *marg)
{
    G g{n, m, E, cp, ev, vp, ve};
    for (int b = 0; b < B; ++b)
        decode_bp<double>(g, algo, priors + (size_t)b * n, y_hard ? y_hard + (size_t)b * n : nullptr, max_iter,
                          x_hat + (size_t)b * n, iters + b, marg ? marg + (size_t)b * n : nullptr);
    return 0;
}

int emu_bp_f32(int algo, int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
               int B, const float *priors, const uint8_t *y_hard, int max_iter, uint8_t *x_hat, int32_t *iters, float *marg)
{
    G g{n, m, E, cp, ev, vp, ve};
    for (int b = 0; b < B; ++b)
        decode_bp<float>(g, algo, priors + (size_t)b * n, y_hard ? y_hard + (size_t)b * n : nullptr, max_iter,
                         x_hat + (size_t)b * n, iters + b, marg ? marg + (size_t)b * n : nullptr);
    return 0;
}

int emu_cn_msa_bits(int n, int m, int E, const int32_t *cp, const int32_t *ev, const float *v2c, float *c2v)
{
    G g{n, m, E, cp, ev, nullptr, nullptr};
    cn_msa_bits_all(g, v2c, c2v);
    return 0;
}

void emu_set_spa_rule(int r) { g_spa_rule = r; }

int emu_cn_sc(int n, int m, int E, const int32_t *cp, const int32_t *ev, const float *v2c, float *c2v, float sat)
{
    G g{n, m, E, cp, ev, nullptr, nullptr};
    cn_spa_sc_all(g, v2c, c2v, sat);
    return 0;
}

int emu_cn_phi(int n, int m, int E, const int32_t *cp, const int32_t *ev, const float *v2c, float *c2v)
{
    G g{n, m, E, cp, ev, nullptr, nullptr};
    cn_spa_phi_all(g, v2c, c2v);
    return 0;
}

// BEC on bit planes, 32 frames per word, with the same book-keeping as bec_book / bec_cn / bec_vn.
int emu_bec(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
            int B, const uint8_t *y, int max_iter, int nb_bits, uint8_t *x_hat, int32_t *iters, uint8_t *reason)
{
    const int wpr = (B + 31) / 32;
    const int limit = max_iter > 0 ? max_iter : n + 1;
    std::vector<uint32_t> mnz((size_t)E * wpr), mpos((size_t)E * wpr), pnz((size_t)n * wpr, 0), ppos((size_t)n * wpr, 0),
        xe((size_t)n * wpr, 0), xv((size_t)n * wpr, 0), act(wpr, 0), changed(wpr, 0), haser(wpr, 0), stopped(wpr, 0);
    std::vector<int32_t> its((size_t)wpr * 32, 0);
    for (int f = 0; f < B; ++f) {
        act[f >> 5] |= 1u << (f & 31);
        for (int v = 0; v < n; ++v) {
            const uint8_t s = y[(size_t)f * n + v];
            const uint32_t bit = 1u << (f & 31);
            const size_t i = (size_t)v * wpr + (f >> 5);
            if (s >= 2) { xe[i] |= bit; haser[f >> 5] |= bit; }
            else { pnz[i] |= bit; if (s == 1) { ppos[i] |= bit; xv[i] |= bit; } }
        }
    }
    auto book = [&](bool first, bool last) {
        for (int w = 0; w < wpr; ++w) {
            uint32_t a = act[w];
            const uint32_t ch = changed[w], h = haser[w];
            if (!first) {
                stopped[w] |= a & ~ch;
                a &= ch;
                for (int l = 0; l < 32; ++l) if ((a >> l) & 1u) its[(size_t)w * 32 + l] += 1;
            }
            if (!last) { a &= h; haser[w] = 0; }
            act[w] = a; changed[w] = 0;
        }
    };
    int it = 0;
    for (; it < limit; ++it) {
        book(it == 0, false);
        const bool first = it == 0;
        for (int w = 0; w < wpr; ++w) {
            if (!act[w]) continue;
            for (int c = 0; c < m; ++c) {
                BecCnAcc acc; acc.init();
                for (int e = cp[c]; e < cp[c + 1]; ++e) {
                    const size_t r = first ? (size_t)ev[e] * wpr + w : (size_t)e * wpr + w;
                    acc.push(first ? pnz[r] : mnz[r], first ? ppos[r] : mpos[r]);
                }
                for (int e = cp[c]; e < cp[c + 1]; ++e) {
                    const size_t r = (size_t)e * wpr + w, rp = (size_t)ev[e] * wpr + w;
                    uint32_t onz, opos;
                    acc.out(first ? pnz[rp] : mnz[r], first ? ppos[rp] : mpos[r], onz, opos);
                    mnz[r] = onz; mpos[r] = opos;
                }
            }
        }
        for (int w = 0; w < wpr; ++w) {
            const uint32_t run = act[w];
            if (!run) continue;
            for (int v = 0; v < n; ++v) {
                const size_t rv = (size_t)v * wpr + w;
                uint32_t nz, pos;
                auto body = [&](auto acc) {
                    acc.set_ternary(pnz[rv], ppos[rv]);
                    for (int k = vp[v]; k < vp[v + 1]; ++k) { const size_t r = (size_t)ve[k] * wpr + w; acc.add_ternary(mnz[r], mpos[r]); }
                    for (int k = vp[v]; k < vp[v + 1]; ++k) {
                        const size_t r = (size_t)ve[k] * wpr + w;
                        auto t = acc;
                        t.sub_ternary(mnz[r], mpos[r]);
                        uint32_t a, b; t.sign(a, b);
                        mnz[r] = a; mpos[r] = b;
                    }
                    acc.sign(nz, pos);
                };
                if (nb_bits == 5) body(BsInt<5>()); else body(BsInt<8>());
                const uint32_t xe_new = ~nz, xv_new = pos, xe_old = xe[rv], xv_old = xv[rv];
                changed[w] |= ((xe_new ^ xe_old) | (~xe_new & (xv_new ^ xv_old))) & run;
                haser[w] |= xe_new & run;
                xe[rv] = (xe_old & ~run) | (xe_new & run);
                xv[rv] = (xv_old & ~run) | (xv_new & run & ~xe_new);
            }
        }
    }
    if (it == limit) book(false, true);
    for (int f = 0; f < B; ++f) {
        const uint32_t bit = 1u << (f & 31);
        for (int v = 0; v < n; ++v) {
            const size_t i = (size_t)v * wpr + (f >> 5);
            x_hat[(size_t)f * n + v] = (xe[i] & bit) ? 2 : ((xv[i] & bit) ? 1 : 0);
        }
        iters[f] = its[f];
        reason[f] = (stopped[f >> 5] & bit) ? 2 : ((act[f >> 5] & bit) ? (max_iter > 0 ? 1 : 4) : 0);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The on-chip kernel's FORMULATION of a decode (csrc/resident_bp.cuh), one frame, plain loops:
//   state = marg[position], prior[position], c2v[plane][check position]; v2c is never stored:
//   CN: v2c = marg - c2v_old, syndrome from the sign bits of marg, lean min-sum for degree 6 (else the bit-pattern rule),
//       edges of a check visited in the PLANNED order; VN: marg = prior + ordered sum of c2v (reference edge order);
//   -0.0 priors folded to +0.0; exit rules per frame as in the kernel's book-keeping.
// Driven with the placement of res_layout.h, so the permutation tables are exercised on the CPU too.
// ---------------------------------------------------------------------------------------------------------------
// vplane: the variable-plane variant (csrc/resident_vp.cuh): placement with vn_contiguous, messages stored at
// plane[rank of the edge at its variable][variable position], the sum starts from c0 instead of 0 + c0, and the word
// is read off the sign bits of marg.
static int emu_resident_msa_impl(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                                 double effort, int B, const float *priors, const uint8_t *y_hard, int limit,
                                 uint8_t *x_hat, int32_t *iters, uint8_t *decoded, bool vplane)
{
    ResPlanner planner(n, m, E, cp, ev, vp, ve, 8);
    const ResLayout L = planner.plan(12345u, effort, vplane, false);
    std::vector<int> slot_of((size_t)E, 0);                                       // rank of an edge among its variable's edges
    for (int v = 0; v < n; ++v)
        for (int p0 = vp[v], k = 0; p0 < vp[v + 1]; ++p0, ++k) slot_of[ve[p0]] = k;
    std::vector<int> cslot((size_t)L.mp * 8, 0);
    std::vector<int> edge_chk((size_t)E);
    for (int c = 0; c < m; ++c)
        for (int e = cp[c]; e < cp[c + 1]; ++e) edge_chk[e] = c;
    // position-indexed tables, exactly what build_resident uploads
    std::vector<int> cvar((size_t)L.mp * 8, -1), vrow((size_t)L.np * 8, -1), cdeg((size_t)L.mp, 0), vdeg((size_t)L.np, 0);
    for (int c = 0; c < m; ++c) cdeg[L.cpos[c]] = cp[c + 1] - cp[c];
    for (int e = 0; e < E; ++e) {
        cvar[(size_t)L.cpos[edge_chk[e]] * 8 + L.eord[e]] = L.vpos[ev[e]];
        cslot[(size_t)L.cpos[edge_chk[e]] * 8 + L.eord[e]] = slot_of[e];
    }
    for (int v = 0; v < n; ++v) {
        vdeg[L.vpos[v]] = vp[v + 1] - vp[v];
        for (int p0 = vp[v], k = 0; p0 < vp[v + 1]; ++p0, ++k)
            vrow[(size_t)L.vpos[v] * 8 + k] = L.eord[ve[p0]] * L.mp + L.cpos[edge_chk[ve[p0]]];
    }
    std::vector<float> marg((size_t)L.np), prior((size_t)L.np), c2v((size_t)8 * L.mp), plane((size_t)8 * L.np);
    std::vector<uint8_t> hb((size_t)L.np);
    for (int b = 0; b < B; ++b) {
        std::fill(marg.begin(), marg.end(), 0.f);
        std::fill(c2v.begin(), c2v.end(), 0.f);
        std::fill(plane.begin(), plane.end(), 0.f);
        for (int v = 0; v < n; ++v) {
            volatile float val = priors[(size_t)b * n + v] + 0.0f;                 // -0.0 -> +0.0
            marg[L.vpos[v]] = prior[L.vpos[v]] = val;
            hb[L.vpos[v]] = y_hard ? y_hard[(size_t)b * n + v] : 0;
        }
        bool fresh = true, done = false;
        int it = 0;
        if (y_hard) {                                                              // iteration-0 exit on the received bits
            bool unsat = false;
            for (int c = 0; c < L.mp && !unsat; ++c) {
                unsigned s = 0;
                for (int k = 0; k < cdeg[c]; ++k) s ^= hb[cvar[(size_t)c * 8 + k]];
                unsat = s & 1u;
            }
            if (!unsat) {
                for (int v = 0; v < n; ++v) x_hat[(size_t)b * n + v] = hb[L.vpos[v]];
                iters[b] = 0; decoded[b] = 1;
                continue;
            }
        }
        for (;;) {
            bool unsat = false;
            for (int c = 0; c < L.mp; ++c) {
                const int dc = cdeg[c];
                if (dc == 0) continue;
                float a[8], o[8];
                uint32_t sx = 0u;
                for (int k = 0; k < 8; ++k) a[k] = INFINITY;
                for (int k = 0; k < dc; ++k) {
                    const float mv = marg[cvar[(size_t)c * 8 + k]];
                    sx ^= f32_bits(mv);
                    a[k] = num<float>::sub(mv, c2v[(size_t)k * L.mp + c]);
                }
                unsat |= (sx >> 31) != 0u;
                if (dc == 6) {
                    float a6[6], o6[6];
                    for (int k = 0; k < 6; ++k) a6[k] = a[k];
                    cn_msa_lean<6>(a6, o6);
                    for (int k = 0; k < 6; ++k) o[k] = o6[k];
                } else {
                    cn_msa_bits<8>(a, dc, o);
                }
                for (int k = 0; k < dc; ++k) {
                    c2v[(size_t)k * L.mp + c] = o[k];                              // the kernel keeps these in registers
                    plane[(size_t)cslot[(size_t)c * 8 + k] * L.np + cvar[(size_t)c * 8 + k]] = o[k];
                }
            }
            if (!fresh && !unsat) { done = true; break; }                          // syndrome of the last hard decisions
            fresh = false;
            ++it;
            for (int p0 = 0; p0 < L.np; ++p0) {
                if (vplane) {
                    float s = vdeg[p0] > 0 ? plane[p0] : 0.0f;
                    for (int k = 1; k < vdeg[p0]; ++k) s = num<float>::add(s, plane[(size_t)k * L.np + p0]);
                    marg[p0] = num<float>::add(prior[p0], s);
                } else {
                    float s = 0.0f;
                    for (int k = 0; k < vdeg[p0]; ++k) s = num<float>::add(s, c2v[vrow[(size_t)p0 * 8 + k]]);
                    marg[p0] = num<float>::add(prior[p0], s);
                }
            }
            if (it >= limit) break;
        }
        for (int v = 0; v < n; ++v)
            x_hat[(size_t)b * n + v] = vplane ? (uint8_t)(f32_bits(marg[L.vpos[v]]) >> 31) : (uint8_t)(marg[L.vpos[v]] < 0.0f);
        iters[b] = it; decoded[b] = done ? 1 : 0;
    }
    return 0;
}

int emu_resident_msa(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                     double effort, int B, const float *priors, const uint8_t *y_hard, int limit,
                     uint8_t *x_hat, int32_t *iters, uint8_t *decoded)
{
    return emu_resident_msa_impl(n, m, E, cp, ev, vp, ve, effort, B, priors, y_hard, limit, x_hat, iters, decoded, false);
}
int emu_resident_vp_msa(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                        double effort, int B, const float *priors, const uint8_t *y_hard, int limit,
                        uint8_t *x_hat, int32_t *iters, uint8_t *decoded)
{
    return emu_resident_msa_impl(n, m, E, cp, ev, vp, ve, effort, B, priors, y_hard, limit, x_hat, iters, decoded, true);
}

// The variable-plane kernel for IRREGULAR codes (resident_vp.cuh, IRR = true), driven by the very tables the library
// uploads (res_layout.h build_vx_tables): one 32-bit word per edge, planes that are prefixes of the degree-sorted
// positions, checks padded to 6 edges with +inf reads and scratch writes, lean min-sum on the padded check.
// One lane of the kernel's float4 cells; returns -1 when the code has no such layout.
int emu_resident_vx_msa(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                        double effort, int B, const float *priors, const uint8_t *y_hard, int limit,
                        uint8_t *x_hat, int32_t *iters, uint8_t *decoded, int32_t *info)
{
    ResPlanner planner(n, m, E, cp, ev, vp, ve, 8);
    const ResLayout L = planner.plan(12345u, effort, true, false);
    VxTables X;
    if (!build_vx_tables(L, n, m, E, cp, ev, vp, ve, 6, &X)) return -1;
    const int np = L.np, mp = L.mp;
    const int prior0 = vx_planes_offset(np) / 16 + X.plane_cells + 8;              // prior cells follow the scratch cells
    if (info) { info[0] = X.plane_cells; for (int k = 0; k < 8; ++k) info[1 + k] = X.pcnt[k]; }
    // structural checks of the tables: every real edge has its own message cell inside its plane, bank group of the
    // message cell == bank group of the marginal cell it belongs to, padding edges stay in the padding / scratch cells
    {
        std::vector<int> seen((size_t)prior0, 0);
        for (int c = 0; c < mp; ++c) {
            const int dc = (int)(X.cwx[(size_t)c * 8] & 15u);
            for (int k = 0; k < 6; ++k) {
                const uint32_t w = X.cwx[(size_t)c * 8 + k];
                const int g = (int)((w & 0xfff0u) >> 4), cell = (int)(w >> 16);
                if ((g & 7) != (cell & 7)) return -2;
                if (k < dc) { if (g >= np || cell < vx_planes_offset(np) / 16 || cell >= prior0 - 8 || seen[cell]++) return -3; }
                else if (g < np || g >= np + 8 || cell < prior0 - 8 || cell >= prior0) return -4;
            }
        }
    }
    std::vector<float> cell((size_t)prior0 + np, 0.f), old((size_t)mp * 6, 0.f);
    std::vector<uint8_t> hb((size_t)np + 8, 0);
    for (int b = 0; b < B; ++b) {
        std::fill(cell.begin(), cell.end(), 0.f);
        std::fill(old.begin(), old.end(), 0.f);
        std::fill(hb.begin(), hb.end(), 0);
        for (int a = 0; a < 8; ++a) cell[np + a] = INFINITY;
        for (int v = 0; v < n; ++v) {
            volatile float val = priors[(size_t)b * n + v] + 0.0f;                 // -0.0 -> +0.0
            cell[X.vposmap[v]] = cell[prior0 + X.vposmap[v]] = val;
            hb[X.vposmap[v]] = y_hard ? y_hard[(size_t)b * n + v] : 0;
        }
        bool fresh = true, done = false;
        int it = 0;
        if (y_hard) {
            bool unsat = false;
            for (int c = 0; c < mp && !unsat; ++c) {
                unsigned s = 0;
                for (int k = 0; k < 6; ++k) s ^= hb[(X.cwx[(size_t)c * 8 + k] & 0xfff0u) >> 4];
                unsat = s & 1u;
            }
            if (!unsat) {
                for (int pz = 0; pz < np; ++pz)
                    if (X.vinvmap[pz] != 0xffffu) x_hat[(size_t)b * n + X.vinvmap[pz]] = hb[pz];
                iters[b] = 0; decoded[b] = 1;
                continue;
            }
        }
        for (;;) {
            bool unsat = false;
            for (int c = 0; c < mp; ++c) {
                float a[6], o[6];
                uint32_t sx = 0u;
                for (int k = 0; k < 6; ++k) {
                    const float mv = cell[(X.cwx[(size_t)c * 8 + k] & 0xfff0u) >> 4];
                    sx ^= f32_bits(mv);
                    a[k] = num<float>::sub(mv, old[(size_t)c * 6 + k]);
                }
                unsat |= (sx >> 31) != 0u;
                cn_msa_lean<6>(a, o);
                for (int k = 0; k < 6; ++k) {
                    old[(size_t)c * 6 + k] = o[k];
                    cell[X.cwx[(size_t)c * 8 + k] >> 16] = o[k];
                }
            }
            if (!fresh && !unsat) { done = true; break; }
            fresh = false;
            ++it;
            for (int pz = 0; pz < np; ++pz) {
                float s = pz < X.pcnt[0] ? cell[X.pbase[0] / 16 + pz] : 0.0f;
                for (int k = 1; k < 8; ++k) {
                    if (pz >= X.pcnt[k]) break;
                    s = num<float>::add(s, cell[X.pbase[k] / 16 + pz]);
                }
                cell[pz] = num<float>::add(cell[prior0 + pz], s);
            }
            if (it >= limit) break;
        }
        for (int pz = 0; pz < np; ++pz)
            if (X.vinvmap[pz] != 0xffffu) x_hat[(size_t)b * n + X.vinvmap[pz]] = (uint8_t)(f32_bits(cell[pz]) >> 31);
        iters[b] = it; decoded[b] = done ? 1 : 0;
    }
    return 0;
}

// Placement statistics and tables of res_layout.h (stats[7] as ldpc_resident_plan).
static int emu_plan_impl(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                         int G, double effort, long *stats, int32_t *cpos, int32_t *vpos, uint8_t *eord, int mode);
int emu_plan(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
             int G, double effort, long *stats, int32_t *cpos, int32_t *vpos, uint8_t *eord)
{
    return emu_plan_impl(n, m, E, cp, ev, vp, ve, G, effort, stats, cpos, vpos, eord, 0);
}
int emu_plan_vp(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                int G, double effort, long *stats, int32_t *cpos, int32_t *vpos, uint8_t *eord)
{
    return emu_plan_impl(n, m, E, cp, ev, vp, ve, G, effort, stats, cpos, vpos, eord, 1);
}
int emu_plan_vp_natural(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                        int G, double effort, long *stats, int32_t *cpos, int32_t *vpos, uint8_t *eord)
{
    return emu_plan_impl(n, m, E, cp, ev, vp, ve, G, effort, stats, cpos, vpos, eord, 2);
}
static int emu_plan_impl(int n, int m, int E, const int32_t *cp, const int32_t *ev, const int32_t *vp, const int32_t *ve,
                         int G, double effort, long *stats, int32_t *cpos, int32_t *vpos, uint8_t *eord, int mode)
{
    ResPlanner planner(n, m, E, cp, ev, vp, ve, G);
    const ResLayout L = planner.plan(12345u, effort, mode != 0, mode == 2);
    const long st[9] = {L.cn_ideal, L.cn_file, L.cn_plan_natural, L.cn_plan, L.vn_ideal, L.vn_file, L.vn_plan, L.mp, L.np};
    for (int i = 0; i < 9; ++i) stats[i] = st[i];
    for (int c = 0; c < m; ++c) cpos[c] = L.cpos[c];
    for (int v = 0; v < n; ++v) vpos[v] = L.vpos[v];
    for (int e = 0; e < E; ++e) eord[e] = L.eord[e];
    return 0;
}

// llr_biawgn_f32 (fast exact path) next to the plain division, elementwise.
int emu_llr_biawgn(size_t count, const double *y, double noise_var, float *fast, float *exact, int64_t *slow_path)
{
    const double inv = 1.0 / noise_var;
    int64_t slow = 0;
    for (size_t i = 0; i < count; ++i) {
        fast[i] = llr_biawgn_f32(y[i], noise_var, inv);
        exact[i] = (float)((-2.0 * y[i]) / noise_var);
        slow += llr_biawgn_fast_ok((-2.0 * y[i]) * inv) ? 0 : 1;
    }
    *slow_path = slow;
    return 0;
}


// Philox4x32-10 block and the Box-Muller normals built on it (csrc/channel_gen.cuh).
int emu_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    const Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    return 0;
}
int emu_normals(unsigned long long seed, unsigned long long frame0, int frames, int n, float *z)
{
    for (int f = 0; f < frames; ++f)
        for (int g = 0; g < (n + 3) / 4; ++g) {
            const Philox4 r = channel_words(seed, frame0 + f, (uint32_t)g);
            float zz[4];
            box_muller(r.x, r.y, &zz[0], &zz[1]);
            box_muller(r.z, r.w, &zz[2], &zz[3]);
            for (int j = 0; j < 4 && 4 * g + j < n; ++j) z[(size_t)f * n + 4 * g + j] = zz[j];
        }
    return 0;
}

}
