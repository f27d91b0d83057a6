// res_layout.h — bank-conflict-aware placement of a Tanner graph in shared memory (host code, no CUDA).
//
// The on-chip decoder (resident_bp.cuh) gathers 16-byte cells at graph-determined addresses: a check thread reads
// the marginals of its variables, a variable thread reads the messages of its checks.  A 128-bit shared-memory
// access is served one quarter-warp (8 lanes) per wavefront, conflict-free iff the 8 cells lie in 8 different
// 16-byte bank groups, i.e. iff their cell indices differ mod G = 8 (with Q quads per cell row G = 8 / Q row
// colours).  With the graph in file order those indices are random and a gather costs ~2.4 wavefronts.
//
// Everything the kernel touches is position-indexed, so the placement is free:
//   cpos[c]   position of check c    -> colour of all its message rows = cpos mod G; thread item order of the CN phase
//   vpos[v]   position of variable v -> colour of its marginal cell    = vpos mod G; thread item order of the VN phase
//   eord[e]   plane (step) of edge e inside its check (min-sum only: the check rule is symmetric in its edges;
//             the variable-node sum keeps the reference's ascending-edge order, bpa.py:35)
// A group = G consecutive positions = the items of one quarter-warp.  Wanted:
//   VN: for every variable group and every k, the k-th checks of its variables have G different colours;
//   CN: for every check group and every step, the variables read in that step have G different colours.
// Simulated annealing over position swaps minimises the predicted extra wavefronts of both phases, then each
// check group orders its edges by local search.  Any placement is CORRECT (results do not depend on it: the
// arithmetic per check / variable is unchanged); a better one is only faster.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace ldpc {

struct ResLayout {
    int G = 8;                          // colours = cells per conflict-free quarter-warp access
    int mp = 0, np = 0;                 // positions (multiples of G; holes are degree-0 items)
    std::vector<int> cpos, vpos;        // [m], [n]
    std::vector<int> cinv, vinv;        // [mp], [np]: item at a position, -1 = hole
    std::vector<uint8_t> eord;          // [E] plane of every edge within its check (optimised, min-sum)
    std::vector<uint8_t> enat;          // [E] natural plane (position in np.where order) for order-sensitive rules
    // predicted shared-memory wavefronts per iteration of the two gather phases (ideal = one per quarter-warp access)
    long cn_ideal = 0, cn_file = 0, cn_plan = 0, cn_plan_natural = 0;
    long vn_ideal = 0, vn_file = 0, vn_plan = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// Tables of the variable-plane kernel for IRREGULAR codes (resident_vp.cuh, IRR = true), from a vn_contiguous layout.
// Shared memory, in 16-byte cells: marg [np] + 8 padding cells (+inf) | plane 0 .. plane max_dv-1, plane k = the first
// pcnt[k] positions | 8 scratch cells | prior ...   (vx_planes_offset is the one definition of where the planes start).
// ---------------------------------------------------------------------------------------------------------------
inline int vx_planes_offset(int np) { return (np + 8) * 16; }

struct VxTables {
    int plane_cells = 0;
    int pcnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // cells of plane k: 8 * (largest number of variables of one colour with > k edges)
    int pbase[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // byte offset of plane k from the start of shared memory
    std::vector<uint32_t> cwx;                  // [mp][8]: (message cell << 16) | (variable position << 4), word 0 also | degree
    std::vector<uint16_t> vposmap, vinvmap;     // [n] position of a variable, [np] variable at a position (0xffff = hole)
};

// dc_pad: edges per check in the kernel (every check is padded to it).  Returns false when the code does not fit the
// 16-bit fields (np <= 4088, cells < 65536) or a degree is out of range (checks 0 or 2..dc_pad, variables <= 8).
inline bool build_vx_tables(const ResLayout &V, int n, int m, int E, const int32_t *chk_ptr, const int32_t *edge_var,
                            const int32_t *var_ptr, const int32_t *var_edges, int dc_pad, VxTables *out)
{
    const int np = V.np, mp = V.mp, G = 8;
    if (np > 4088 || dc_pad > 8) return false;
    VxTables T;
    int cnt[8][8] = {};                                                 // [k][colour]
    for (int v = 0; v < n; ++v) {
        const int d = var_ptr[v + 1] - var_ptr[v];
        if (d > 8) return false;
        for (int k = 0; k < d; ++k) cnt[k][V.vpos[v] % G] += 1;
    }
    int off = vx_planes_offset(np);
    for (int k = 0; k < 8; ++k) {
        int mx = 0;
        for (int a = 0; a < G; ++a) mx = std::max(mx, cnt[k][a]);
        T.pcnt[k] = mx * G;
        T.pbase[k] = off;
        off += T.pcnt[k] * 16;
        T.plane_cells += T.pcnt[k];
    }
    const int scratch_cell = off / 16;
    if (scratch_cell + 8 > 65535) return false;
    std::vector<int> slot((size_t)E, 0);                                // rank of an edge among its variable's edges
    for (int v = 0; v < n; ++v)
        for (int p0 = var_ptr[v], k = 0; p0 < var_ptr[v + 1]; ++p0, ++k) slot[var_edges[p0]] = k;
    T.cwx.assign((size_t)mp * 8, 0u);
    std::vector<uint8_t> have((size_t)mp * 8, 0);
    for (int c = 0; c < m; ++c) {
        const int d = chk_ptr[c + 1] - chk_ptr[c];
        if (d > dc_pad || d == 1) return false;
        for (int e = chk_ptr[c]; e < chk_ptr[c + 1]; ++e) {
            const int vp = V.vpos[edge_var[e]], k = V.eord[e];
            if (vp >= T.pcnt[slot[e]]) return false;                    // cannot happen: positions are degree-sorted per colour
            T.cwx[(size_t)V.cpos[c] * 8 + k] = ((uint32_t)(T.pbase[slot[e]] / 16 + vp) << 16) | ((uint32_t)vp << 4);
            have[(size_t)V.cpos[c] * 8 + k] = 1;
        }
        T.cwx[(size_t)V.cpos[c] * 8] |= (uint32_t)d;
    }
    // padding edges: per (group of 8 check positions, step) a bank group (colour) no real edge of the step uses
    for (int g = 0; g < mp / G; ++g)
        for (int k = 0; k < dc_pad; ++k) {
            bool used[8] = {false, false, false, false, false, false, false, false};
            for (int i = 0; i < G; ++i)
                if (have[(size_t)(g * G + i) * 8 + k]) used[(T.cwx[(size_t)(g * G + i) * 8 + k] >> 4) & 7u] = true;
            for (int i = 0; i < G; ++i) {
                if (have[(size_t)(g * G + i) * 8 + k]) continue;
                int a = 0;
                while (a < 7 && used[a]) ++a;
                used[a] = true;
                T.cwx[(size_t)(g * G + i) * 8 + k] |= ((uint32_t)(scratch_cell + a) << 16) | ((uint32_t)(np + a) << 4);
            }
        }
    T.vposmap.resize((size_t)n);
    T.vinvmap.assign((size_t)np, 0xffffu);
    for (int v = 0; v < n; ++v) { T.vposmap[v] = (uint16_t)V.vpos[v]; T.vinvmap[V.vpos[v]] = (uint16_t)v; }
    *out = T;
    return true;
}

class ResPlanner {
  public:
    ResPlanner(int n, int m, int E, const int32_t *chk_ptr, const int32_t *edge_var, const int32_t *var_ptr,
               const int32_t *var_edges, int G)
        : n_(n), m_(m), E_(E), G_(G), chk_ptr_(chk_ptr), edge_var_(edge_var), var_ptr_(var_ptr), var_edges_(var_edges)
    {
        edge_chk_.resize(E);
        for (int c = 0; c < m; ++c)
            for (int e = chk_ptr[c]; e < chk_ptr[c + 1]; ++e) edge_chk_[e] = c;
        mp_ = (m + G - 1) / G * G;
        np_ = (n + G - 1) / G * G;
    }

    // vn_contiguous: the variable phase reads contiguous cells whatever the placement (resident_vp.cuh stores a message
    // where its variable reads it), and the check phase scatters to the bank group it gathered from, so only the check
    // gathers count; they are annealed on their exact cost in the NATURAL edge order, which serves the order-sensitive
    // sum-product rule as well as min-sum.
    // natural_order (with vn_contiguous): the edges of a check keep their np.where order (sum-product); the placement is
    // then a balanced colouring of the variables, see colour_variables.
    ResLayout plan(unsigned seed = 12345u, double effort = 1.0, bool vn_contiguous = false, bool natural_order = false)
    {
        vn_free_ = vn_contiguous;
        ResLayout L;
        L.G = G_; L.mp = mp_; L.np = np_;
        cpos_.resize(m_); vpos_.resize(n_);
        cinv_.assign(mp_, -1); vinv_.assign(np_, -1);
        for (int c = 0; c < m_; ++c) { cpos_[c] = c; cinv_[c] = c; }
        for (int v = 0; v < n_; ++v) { vpos_[v] = v; vinv_[v] = v; }
        L.enat.resize(E_);
        for (int c = 0; c < m_; ++c)
            for (int e = chk_ptr_[c]; e < chk_ptr_[c + 1]; ++e) L.enat[e] = (uint8_t)(e - chk_ptr_[c]);
        eord_ = L.enat;
        L.cn_ideal = cn_ideal();
        L.vn_ideal = vn_ideal();
        L.cn_file = L.cn_ideal + cn_extra_exact_all();
        L.vn_file = L.vn_ideal + vn_cost_all();

        if (vn_free_) colour_variables(seed, effort, natural_order);
        else anneal(seed, effort);
        L.cn_plan_natural = L.cn_ideal + cn_extra_exact_all();
        if (!(vn_free_ && natural_order)) {
            if (vn_free_) match_edges();
            order_edges(seed ^ 0x9e3779b9u, effort);
        }
        L.cn_plan = L.cn_ideal + cn_extra_exact_all();
        L.vn_plan = L.vn_ideal + (vn_free_ ? 0 : vn_cost_all());
        L.cpos = cpos_; L.vpos = vpos_; L.cinv = cinv_; L.vinv = vinv_; L.eord = eord_;
        return L;
    }

  private:
    int n_, m_, E_, G_, mp_, np_;
    const int32_t *chk_ptr_, *edge_var_, *var_ptr_, *var_edges_;
    std::vector<int> edge_chk_, cpos_, vpos_, cinv_, vinv_;
    std::vector<uint8_t> eord_;
    std::vector<int> vgc_, cgc_;        // cached group costs
    bool vn_free_ = false;
    uint64_t rng_ = 88172645463325252ull;

    uint32_t rnd()
    {
        rng_ ^= rng_ << 13; rng_ ^= rng_ >> 7; rng_ ^= rng_ << 17;
        return (uint32_t)(rng_ >> 32);
    }
    double rnd01() { return (rnd() >> 8) * (1.0 / 16777216.0); }

    long cn_ideal() const
    {
        long w = 0;
        for (int g = 0; g < mp_ / G_; ++g) {
            int d = 0;
            for (int i = 0; i < G_; ++i) { const int c = cinv_[g * G_ + i]; if (c >= 0) d = std::max(d, chk_ptr_[c + 1] - chk_ptr_[c]); }
            w += d;
        }
        return w;
    }
    long vn_ideal() const
    {
        long w = 0;
        for (int g = 0; g < np_ / G_; ++g) {
            int d = 0;
            for (int i = 0; i < G_; ++i) { const int v = vinv_[g * G_ + i]; if (v >= 0) d = std::max(d, var_ptr_[v + 1] - var_ptr_[v]); }
            w += d;
        }
        return w;
    }

    // Extra wavefronts of one variable group: for every k, (largest colour multiplicity among the k-th checks) - 1.
    int vn_group_cost(int g) const
    {
        int cost = 0;
        for (int k = 0;; ++k) {
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
            bool any = false;
            for (int i = 0; i < G_; ++i) {
                const int v = vinv_[g * G_ + i];
                if (v < 0 || var_ptr_[v + 1] - var_ptr_[v] <= k) continue;
                any = true;
                const int col = cpos_[edge_chk_[var_edges_[var_ptr_[v] + k]]] % G_;
                mx = std::max(mx, ++cnt[col]);
            }
            if (!any) break;
            cost += mx - 1;
        }
        return cost;
    }
    // Proxy for one check group while positions move: colour counts beyond the number of steps cannot be hidden by any
    // edge order (each wavefront serves one cell per colour).
    int cn_group_cost(int g) const
    {
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, steps = 0;
        for (int i = 0; i < G_; ++i) {
            const int c = cinv_[g * G_ + i];
            if (c < 0) continue;
            steps = std::max(steps, chk_ptr_[c + 1] - chk_ptr_[c]);
            for (int e = chk_ptr_[c]; e < chk_ptr_[c + 1]; ++e) ++cnt[vpos_[edge_var_[e]] % G_];
        }
        int cost = 0;
        for (int a = 0; a < G_; ++a) cost += std::max(0, cnt[a] - steps);
        return cost;
    }
    // Exact extra wavefronts of one check group for the current edge planes.
    int cn_group_exact(int g) const
    {
        uint8_t cnt[8][8] = {};
        uint8_t mx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < G_; ++i) {
            const int c = cinv_[g * G_ + i];
            if (c < 0) continue;
            for (int e = chk_ptr_[c]; e < chk_ptr_[c + 1]; ++e) {
                const int k = eord_[e] & 7;
                const uint8_t v = ++cnt[k][vpos_[edge_var_[e]] % G_];
                if (v > mx[k]) mx[k] = v;
            }
        }
        int cost = 0;
        for (int k = 0; k < 8; ++k) cost += mx[k] ? mx[k] - 1 : 0;
        return cost;
    }
    int vn_cost(int g) const { return vn_free_ ? 0 : vn_group_cost(g); }
    int cn_cost(int g) const { return cn_group_cost(g); }
    long vn_cost_all() const { long s = 0; for (int g = 0; g < np_ / G_; ++g) s += vn_group_cost(g); return s; }
    long cn_extra_exact_all() const { long s = 0; for (int g = 0; g < mp_ / G_; ++g) s += cn_group_exact(g); return s; }

    void touched_by_check(int c, std::vector<int> &vg) const
    {
        if (c < 0) return;
        for (int e = chk_ptr_[c]; e < chk_ptr_[c + 1]; ++e) vg.push_back(vpos_[edge_var_[e]] / G_);
    }
    void touched_by_var(int v, std::vector<int> &cg) const
    {
        if (v < 0) return;
        for (int p = var_ptr_[v]; p < var_ptr_[v + 1]; ++p) cg.push_back(cpos_[edge_chk_[var_edges_[p]]] / G_);
    }
    static void uniq(std::vector<int> &a) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }

    void anneal(unsigned seed, double effort)
    {
        rng_ ^= (uint64_t)seed * 0x9e3779b97f4a7c15ull;
        vgc_.resize(np_ / G_); cgc_.resize(mp_ / G_);
        for (int g = 0; g < np_ / G_; ++g) vgc_[g] = vn_cost(g);
        for (int g = 0; g < mp_ / G_; ++g) cgc_[g] = cn_cost(g);
        const long moves = (long)(effort * 60.0 * (mp_ + np_)) * 10;
        const double t0 = 0.8, t1 = 0.05;
        std::vector<int> vg, cg, oldv, oldc;
        for (long it = 0; it < moves; ++it) {
            const double T = t0 * std::pow(t1 / t0, (double)it / (double)moves);
            const bool move_check = (rnd() % (unsigned)(mp_ + np_)) < (unsigned)mp_;
            vg.clear(); cg.clear();
            int a, b;
            if (move_check) {
                a = (int)(rnd() % (unsigned)mp_); b = (int)(rnd() % (unsigned)mp_);
                if (a == b || (cinv_[a] < 0 && cinv_[b] < 0)) continue;
                cg.push_back(a / G_); cg.push_back(b / G_);
                touched_by_check(cinv_[a], vg); touched_by_check(cinv_[b], vg);
            } else {
                a = (int)(rnd() % (unsigned)np_); b = (int)(rnd() % (unsigned)np_);
                if (a == b || (vinv_[a] < 0 && vinv_[b] < 0)) continue;
                vg.push_back(a / G_); vg.push_back(b / G_);
                touched_by_var(vinv_[a], cg); touched_by_var(vinv_[b], cg);
            }
            uniq(vg); uniq(cg);
            int before = 0;
            for (int g : vg) before += vgc_[g];
            for (int g : cg) before += cgc_[g];
            swap_pos(move_check, a, b);
            oldv.clear(); oldc.clear();
            int after = 0;
            for (int g : vg) { oldv.push_back(vgc_[g]); vgc_[g] = vn_cost(g); after += vgc_[g]; }
            for (int g : cg) { oldc.push_back(cgc_[g]); cgc_[g] = cn_cost(g); after += cgc_[g]; }
            const int d = after - before;
            if (d > 0 && rnd01() >= std::exp(-(double)d / T)) {                         // reject: undo
                swap_pos(move_check, a, b);
                for (size_t i = 0; i < vg.size(); ++i) vgc_[vg[i]] = oldv[i];
                for (size_t i = 0; i < cg.size(); ++i) cgc_[cg[i]] = oldc[i];
            }
        }
    }
    // vn_contiguous placement.  Checks stay in file order (group = 8 consecutive checks); what is left is a balanced
    // 8-colouring of the variables (colour = position mod 8), found by annealing colour swaps on a squared deviation:
    //   natural order (sum-product): the variables read in the same step by the checks of a group should all differ
    //     (every colour once per (group, step): 8-cliques, so 8 colours are the bare minimum and a few conflicts stay);
    //   free order (min-sum): every colour should occur exactly `degree` times among the variables of a group; the
    //     edges of each check are then ordered into conflict-free steps by match_edges.
    void colour_variables(unsigned seed, double effort, bool natural)
    {
        rng_ ^= (uint64_t)seed * 0x9e3779b97f4a7c15ull;
        const int ngroups = mp_ / G_;
        const int nsets = natural ? ngroups * 8 : ngroups;
        std::vector<int> cnt((size_t)nsets * 8, 0), target((size_t)nsets, 1);
        std::vector<std::vector<int>> sets_of((size_t)n_);
        for (int c = 0; c < m_; ++c)
            for (int e = chk_ptr_[c]; e < chk_ptr_[c + 1]; ++e)
                sets_of[edge_var_[e]].push_back(natural ? (c / G_) * 8 + ((e - chk_ptr_[c]) & 7) : c / G_);
        if (!natural)
            for (int g = 0; g < ngroups; ++g) {                        // steps of the group = its largest check degree
                int d = 0;
                for (int c = g * G_; c < std::min(m_, (g + 1) * G_); ++c) d = std::max(d, chk_ptr_[c + 1] - chk_ptr_[c]);
                target[g] = d;
            }
        std::vector<int> col((size_t)n_);
        for (int v = 0; v < n_; ++v) col[v] = v % G_;                  // balanced to start with, and swaps keep it so
        auto sq = [](long x) { return x * x; };
        auto add = [&](int v, int a) { long d = 0; for (int st : sets_of[v]) { int &x = cnt[(size_t)st * 8 + a]; d += sq(x + 1 - target[st]) - sq(x - target[st]); ++x; } return d; };
        auto del = [&](int v, int a) { long d = 0; for (int st : sets_of[v]) { int &x = cnt[(size_t)st * 8 + a]; d += sq(x - 1 - target[st]) - sq(x - target[st]); --x; } return d; };
        for (int v = 0; v < n_; ++v) add(v, col[v]);
        auto excess = [&]() { long t = 0; for (int st = 0; st < nsets; ++st) for (int a = 0; a < 8; ++a) t += std::max(0, cnt[(size_t)st * 8 + a] - target[st]); return t; };
        std::vector<std::vector<int>> members((size_t)nsets);
        for (int v = 0; v < n_; ++v)
            for (int st : sets_of[v]) members[st].push_back(v);
        // Annealing with targeted proposals: take a set that violates its target, a member v of an over-represented
        // colour a, an under-represented colour b of the same set, and any variable w of colour b; swap their colours.
        const long moves = (long)(effort * 6000.0 * n_) + 1;
        const double t0 = 0.6, t1 = 0.08;
        for (long it = 0; it < moves; ++it) {
            if ((it & 1023) == 0 && excess() == 0) break;
            const double T = t0 * std::pow(t1 / t0, (double)it / (double)moves);
            const int st = (int)(rnd() % (unsigned)nsets);
            int over[8], under[8], no = 0, nu = 0;
            for (int c = 0; c < 8; ++c) {
                if (cnt[(size_t)st * 8 + c] > target[st]) over[no++] = c;
                if (cnt[(size_t)st * 8 + c] < target[st]) under[nu++] = c;
            }
            if (no == 0 || nu == 0 || members[st].empty()) continue;
            const int a = over[rnd() % (unsigned)no], b = under[rnd() % (unsigned)nu];
            int v = -1, w = -1;
            for (int t = 0; t < 64 && v < 0; ++t) { const int x = members[st][rnd() % members[st].size()]; if (col[x] == a) v = x; }
            for (int t = 0; t < 256 && w < 0; ++t) { const int x = (int)(rnd() % (unsigned)n_); if (col[x] == b) w = x; }
            if (v < 0 || w < 0) continue;
            const long d = del(v, a) + del(w, b) + add(v, b) + add(w, a);
            if (d > 0 && rnd01() >= std::exp(-(double)d / T)) {        // reject: undo
                del(v, b); del(w, a); add(v, a); add(w, b);
            } else {
                col[v] = b; col[w] = a;
            }
        }
        // positions: the i-th variable of colour a sits at 8 * i + a (the np_ - n_ < 8 holes end up at the top); within
        // a colour the variables are ordered by DESCENDING degree (stable: a regular code keeps file order), so that
        // the variables with more than k edges are a prefix of the positions and plane k of resident_vp.cuh (the k-th
        // message of every variable) is a prefix too — irregular codes then need E + a few cells, not max_dv * n.
        std::vector<int> order((size_t)n_);
        for (int v = 0; v < n_; ++v) order[v] = v;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return var_ptr_[a + 1] - var_ptr_[a] > var_ptr_[b + 1] - var_ptr_[b];
        });
        std::vector<int> next((size_t)G_, 0);
        std::fill(vinv_.begin(), vinv_.end(), -1);
        for (int v : order) {
            const int pos = next[col[v]]++ * G_ + col[v];
            vpos_[v] = pos; vinv_[pos] = v;
        }
    }
    void swap_pos(bool check, int a, int b)
    {
        if (check) {
            std::swap(cinv_[a], cinv_[b]);
            if (cinv_[a] >= 0) cpos_[cinv_[a]] = a;
            if (cinv_[b] >= 0) cpos_[cinv_[b]] = b;
        } else {
            std::swap(vinv_[a], vinv_[b]);
            if (vinv_[a] >= 0) vpos_[vinv_[a]] = a;
            if (vinv_[b] >= 0) vpos_[vinv_[b]] = b;
        }
    }

    // Per check group: the edges of its checks against the colours of their variables form a bipartite multigraph; when
    // every colour occurs exactly `steps` times (proxy cost 0) it is regular and splits into `steps` perfect matchings
    // (Koenig), i.e. into conflict-free steps.  Step by step: a maximum matching of checks to colours over the edges not
    // yet placed (Kuhn's augmenting paths on an 8 x 8 graph); a check left unmatched takes any of its remaining edges.
    void match_edges()
    {
        for (int g = 0; g < mp_ / G_; ++g) {
            std::vector<int> rem[8];                                   // unplaced edges per check of the group
            int steps = 0;
            for (int i = 0; i < G_; ++i) {
                const int c = cinv_[g * G_ + i];
                if (c < 0) continue;
                for (int e = chk_ptr_[c]; e < chk_ptr_[c + 1]; ++e) rem[i].push_back(e);
                steps = std::max(steps, chk_ptr_[c + 1] - chk_ptr_[c]);
            }
            // checks of lower degree only take part in the first dc steps (their planes are 0..dc-1)
            for (int k = 0; k < steps; ++k) {
                int owner[8];                                          // colour -> check index matched to it
                int pick[8];                                           // check index -> chosen edge
                std::fill(owner, owner + 8, -1);
                std::fill(pick, pick + 8, -1);
                for (int i = 0; i < G_; ++i) {
                    if (rem[i].empty()) continue;
                    bool seen[8] = {false, false, false, false, false, false, false, false};
                    augment(i, rem, owner, pick, seen);
                }
                for (int i = 0; i < G_; ++i) {
                    if (rem[i].empty()) continue;
                    int e = pick[i];
                    if (e < 0) e = rem[i].back();
                    eord_[e] = (uint8_t)k;
                    rem[i].erase(std::find(rem[i].begin(), rem[i].end(), e));
                }
            }
        }
    }
    bool augment(int i, std::vector<int> (&rem)[8], int (&owner)[8], int (&pick)[8], bool (&seen)[8])
    {
        for (int e : rem[i]) {
            const int a = vpos_[edge_var_[e]] % G_;
            if (seen[a]) continue;
            seen[a] = true;
            if (owner[a] < 0 || augment(owner[a], rem, owner, pick, seen)) {
                owner[a] = i; pick[i] = e;
                return true;
            }
        }
        return false;
    }

    // Per check group: permute the planes of each check's edges to make every step's colours distinct.
    void order_edges(unsigned seed, double effort)
    {
        rng_ ^= (uint64_t)seed << 17;
        for (int g = 0; g < mp_ / G_; ++g) {
            int cost = cn_group_exact(g);
            const int tries = (int)(effort * 4000);
            for (int t = 0; t < tries && cost > 0; ++t) {
                const int c = cinv_[g * G_ + (int)(rnd() % (unsigned)G_)];
                if (c < 0) continue;
                const int dc = chk_ptr_[c + 1] - chk_ptr_[c];
                if (dc < 2) continue;
                const int e1 = chk_ptr_[c] + (int)(rnd() % (unsigned)dc), e2 = chk_ptr_[c] + (int)(rnd() % (unsigned)dc);
                if (e1 == e2) continue;
                std::swap(eord_[e1], eord_[e2]);
                const int nc = cn_group_exact(g);
                if (nc <= cost) cost = nc;
                else std::swap(eord_[e1], eord_[e2]);
            }
        }
    }
};

}  // namespace ldpc
