// Host-side index tables of the sparse grid: the flattened replacement of the reference's hash-keyed element
// map (DGSolution::dg, include/DGSolution.h:219) and of the per-element neighbour sets
// Element::ptr_vol_alpt / ptr_flx_alpt (include/Element.h:152-155, built by the O(N^2) scans of
// source/DGSolution.cpp:675-728).  Everything here is rebuilt only when the grid changes.
//
// Layout produced (per dimension t):
//   fibres     : elements that agree in (level, suppt) in every dim != t, sorted by 1D order inside a fibre
//   slots      : position of an element in that fibre-major order
//   neighbours : per target slot the sources (fibre-local position, canonical 1D pair id), "U" sources
//                (level <= target) first, then "L" sources (level > target), for the vol and the flx relation
// The canonical 1D pair enumeration (Pairs1D) depends only on NMAX; operators are stored as one small
// (edge_from x edge_to) block per pair.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <unordered_map>
#include <vector>

namespace amdg {

static const double ROUND_OFF = 1e-13;   // include/libs.h:28

// ---- Hash::hash_key (source/Hash.cpp:55-114) ------------------------------------------------------------
// The reference keeps binomials in a table of doubles filled through a double-precision factorial
// (source/Hash.cpp:119-157) and updates int accumulators with double operands; the same operations are
// performed here in the same order so that the key is bit-identical.
struct BinomTable
{
    std::vector<std::vector<double>> t;
    static double fact(double n)
    {
        if (n < 2) return 1.0;
        double result = n, pc = n;
        do { result = result * (pc - 1); pc--; } while (pc > 2);
        return result;
    }
    static double bico(double n, double k)
    {
        if (k == 1) return n;
        return fact(n) / (fact(k) * fact(n - k));
    }
    BinomTable() : t(100)
    {
        for (int i = 0; i < 100; ++i)
        {
            t[i].resize(i + 1);
            for (int j = 0; j <= i; ++j) t[i][j] = (j <= i / 2) ? bico(i, j) : t[i][i - j];
        }
    }
};

inline int hash_key(int d, const int * l, const int * j)
{
    static const BinomTable B;
    // signed overflow is what the reference does for large keys; do it in unsigned to stay defined
    uint32_t ind1 = 0;
    for (int t = 0; t < d; ++t) ind1 = ind1 * (uint32_t)(1u << l[t]) + (uint32_t)((j[t] - 1) / 2);
    int sum = l[0];
    int ind2 = 0;
    for (int t = 1; t < d; ++t)
    {
        ind2 = (int)((double)ind2 - B.t[t + sum][t]);
        sum += l[t];
        ind2 = (int)((double)ind2 + B.t[t + sum][t]);
    }
    uint32_t ind2u = (uint32_t)ind2 * (uint32_t)(1u << sum);
    int ind3 = 0;
    for (int s = 0; s < sum; ++s) ind3 = (int)((double)ind3 + B.t[d - 1 + s][d - 1] * (double)(1 << s));
    return (int)(ind1 + ind2u + (uint32_t)ind3);
}

// ---- 1D conventions -------------------------------------------------------------------------------------
inline int order_elem(int n, int j) { return n == 0 ? 0 : ((1 << (n - 1)) + (j - 1) / 2); }   // source/Element.cpp:388-391
inline int level_of_order(int o) { int n = 0; while (o > 0) { o >>= 1; ++n; } return n; }
inline void support(int n, int j, double & xl, double & xr)                                     // source/Basis.cpp:22-36
{
    if (n <= 1) { xl = 0.; xr = 1.; }
    else { xl = std::pow(2., -n + 1.) * (j - 1.) / 2.; xr = std::pow(2., -n + 1.) * (j + 1.) / 2.; }
}
inline bool interval_intersect(double u0, double u1, double v0, double v1)                    // source/Element.cpp:380-386
{
    return !((u0 >= v1) || (u1 <= v0));
}
inline bool interval_intersect_adjacent(double u0, double u1, double v0, double v1)           // source/Element.cpp:337-348
{
    const bool no_intersect = (u0 >= v1) || (u1 <= v0);
    const bool adjacent = (std::abs(u0 - v1) < ROUND_OFF) || (std::abs(u1 - v0) < ROUND_OFF);
    const bool periodic = ((std::abs(u0) < ROUND_OFF) && (std::abs(v1 - 1.0) < ROUND_OFF)) ||
                          ((std::abs(v0) < ROUND_OFF) && (std::abs(u1 - 1.0) < ROUND_OFF));
    return !no_intersect || adjacent || periodic;
}

// ---- canonical enumeration of related 1D element pairs -----------------------------------------------------
struct Pairs1D
{
    int nmax = 0, T = 0, n_pairs = 0;
    std::vector<int> id;          // [src*T + tgt] -> pair id or -1 (flx relation, a superset of vol)
    std::vector<int> src, tgt;    // per pair
    std::vector<uint8_t> vol;     // per pair: also in the vol relation
    std::vector<int> tgt_ptr;     // pairs are grouped by target, sources ascending: tgt_ptr[T+1]

    void build(int nmax_)
    {
        nmax = nmax_; T = 1 << nmax;
        std::vector<double> xl(T), xr(T);
        for (int n = 0; n <= nmax; ++n)
            for (int j = 1; j < std::max(2, 1 << n); j += 2) support(n, j, xl[order_elem(n, j)], xr[order_elem(n, j)]);
        id.assign((size_t)T * T, -1); src.clear(); tgt.clear(); vol.clear(); tgt_ptr.assign(T + 1, 0);
        for (int e = 0; e < T; ++e)
        {
            for (int f = 0; f < T; ++f)
            {
                // test element e (this), solution element f (elem): Element::is_flx_alpt(elem, dim)
                if (!interval_intersect_adjacent(xl[e], xr[e], xl[f], xr[f])) continue;
                id[(size_t)f * T + e] = (int)src.size();
                src.push_back(f); tgt.push_back(e);
                vol.push_back(interval_intersect(xl[e], xr[e], xl[f], xr[f]) ? 1 : 0);
            }
            tgt_ptr[e + 1] = (int)src.size();
        }
        n_pairs = (int)src.size();
    }
};

struct Nbr { int local; int pair; };

// Neighbour lists of a fibre depend only on its SHAPE (the ascending 1D orders of its elements): a refine / coarsen step
// (source/DGAdapt.cpp:1073-1248) changes a few fibres per dimension and leaves the shape of almost all others as it was, so the lists are
// kept per shape across amdg_grid_set calls and a rebuild only renumbers rows and copies templates (the reference re-scans all element
// pairs, source/DGSolution.cpp:675-728, or walks every element per added one, source/DGAdapt.cpp:1108-1136).  One cache per dimension
// (the dimensions are built on parallel host threads).
struct NbrTemplate
{
    std::vector<int> sig;            // 1D orders, ascending
    std::vector<Nbr> nbr[2];         // vol / flx entries of all targets, target after target
    std::vector<int> cnt[2], nu[2];  // per target: entries, leading "U" entries
    mutable int shape_id = -1;       // memo of ShapeTable (mma_items.hpp), valid while shape_epoch equals the table's epoch
    mutable uint64_t shape_epoch = 0;
};
struct NbrCache
{
    std::unordered_multimap<uint64_t, NbrTemplate> map;
    size_t entries = 0;
    int64_t hits = 0, misses = 0;
    void clear() { map.clear(); entries = 0; }
};

struct DimTables
{
    int64_t n_fibre = 0;
    std::vector<int64_t> fibre_ptr;     // [n_fibre+1] into slots
    std::vector<int> slot_elem;         // [n] element row of a slot
    std::vector<int> elem_slot;         // [n]
    std::vector<int> slot_fibre;        // [n] fibre of a slot
    // relation kind k (0 vol, 1 flx): CSR over slots
    std::vector<int64_t> nbr_ptr[2];    // [n+1]
    std::vector<int> nbr_split[2];      // [n] number of leading "U" entries
    std::vector<Nbr> nbr[2];
    int max_fibre_len = 0;
    std::vector<const NbrTemplate *> fibre_tmpl;   // [n_fibre] template of every fibre when built with a cache (valid until that cache is cleared)
};

struct Grid
{
    int dim = 0, nmax = 0;
    int64_t n = 0;
    std::vector<int> level, suppt, ord1d, hash;   // [n*dim] x3, [n]
    std::vector<DimTables> dims;

    // Hash::hash_key of every element (the parity export amdg_grid_keys), computed on first use
    const std::vector<int> & keys()
    {
        if ((int64_t)hash.size() != n)
        {
            hash.resize(n);
            for (int64_t e = 0; e < n; ++e) hash[e] = hash_key(dim, &level[e * dim], &suppt[e * dim]);
        }
        return hash;
    }

    // returns 0, or -1 for an invalid element / duplicate
    int build(int dim_, int nmax_, int64_t n_, const int * level_, const int * suppt_, const Pairs1D & P, std::vector<NbrCache> * caches = nullptr)
    {
        if (caches && (int)caches->size() != dim_) caches->assign(dim_, NbrCache());
        std::vector<uint8_t> lev_of(P.T);
        for (int o = 0; o < P.T; ++o) lev_of[o] = (uint8_t)level_of_order(o);
        dim = dim_; nmax = nmax_; n = n_;
        level.assign(level_, level_ + n * dim); suppt.assign(suppt_, suppt_ + n * dim);
        ord1d.resize(n * dim); hash.clear();          // hash keys are computed on demand (keys()): only the parity export reads them
        for (int64_t e = 0; e < n; ++e)
        {
            for (int t = 0; t < dim; ++t)
            {
                const int l = level[e * dim + t], j = suppt[e * dim + t];
                if (l < 0 || l > nmax || j < 1 || (j % 2) == 0 || (l == 0 && j != 1) || (l >= 1 && (j - 1) / 2 > (1 << (l - 1)) - 1)) return -1;
                ord1d[e * dim + t] = order_elem(l, j);
            }
        }
        dims.resize(dim);   // a Grid object that is built again keeps the capacity of its tables (no fresh pages to fault in)
        // the tables of the dimensions are independent: one host thread per dimension
        auto build_dim = [&](int t) -> int
        {
            std::vector<int> pos_of_ord(P.T, -1);
            DimTables & D = dims[t];
            std::vector<int> & perm = D.slot_elem;
            perm.resize(n);
            std::iota(perm.begin(), perm.end(), 0);
            const int * o = ord1d.data();
            const int d = dim;
            auto same_fibre = [o, d, t](int a, int b)
            {
                for (int k = 0; k < d; ++k) if (k != t && o[(int64_t)a * d + k] != o[(int64_t)b * d + k]) return false;
                return true;
            };
            if (d * nmax <= 64)
            {
                // the lexicographic key (1D orders of the other dims, then of dim t; each below 2^nmax) fits one 64-bit word
                std::vector<std::pair<uint64_t, int>> keyed(n), tmp;
                for (int64_t e = 0; e < n; ++e)
                {
                    uint64_t key = 0;
                    for (int k = 0; k < d; ++k) if (k != t) key = (key << nmax) | (uint64_t)o[e * d + k];
                    key = (key << nmax) | (uint64_t)o[e * d + t];
                    keyed[e] = { key, (int)e };
                }
                // least-significant-digit radix sort, 11 bits per pass over the d * nmax key bits (stable; 2 passes for the 2-D NMAX = 9 grid, 4 for
                // the 6-D NMAX = 7 grid): a grid change is on the critical path of an adaptive run
                const int bits = d * nmax;
                if (n >= 256)
                {
                    tmp.resize(n);
                    for (int shift = 0; shift < bits; shift += 11)
                    {
                        uint32_t count[2049] = { 0 };
                        for (int64_t e = 0; e < n; ++e) count[((keyed[e].first >> shift) & 2047u) + 1]++;
                        for (int b = 0; b < 2048; ++b) count[b + 1] += count[b];
                        for (int64_t e = 0; e < n; ++e) tmp[count[(keyed[e].first >> shift) & 2047u]++] = keyed[e];
                        keyed.swap(tmp);
                    }
                }
                else std::sort(keyed.begin(), keyed.end());
                for (int64_t s = 0; s < n; ++s) perm[s] = keyed[s].second;
            }
            else
            std::sort(perm.begin(), perm.end(), [o, d, t](int a, int b)
            {
                for (int k = 0; k < d; ++k) if (k != t) { const int x = o[(int64_t)a * d + k], y = o[(int64_t)b * d + k]; if (x != y) return x < y; }
                return o[(int64_t)a * d + t] < o[(int64_t)b * d + t];
            });
            D.elem_slot.resize(n); D.slot_fibre.resize(n);
            D.fibre_ptr.clear(); D.fibre_ptr.push_back(0);
            for (int64_t s = 0; s < n; ++s)
            {
                if (s > 0 && !same_fibre(perm[s - 1], perm[s])) D.fibre_ptr.push_back(s);
                else if (s > 0 && o[(int64_t)perm[s - 1] * d + t] == o[(int64_t)perm[s] * d + t]) return -1;   // duplicate element (of this dimension's pass)
                D.elem_slot[perm[s]] = (int)s;
                D.slot_fibre[s] = (int)D.fibre_ptr.size() - 1;
            }
            D.fibre_ptr.push_back(n);
            D.n_fibre = (int64_t)D.fibre_ptr.size() - 1;
            for (int k = 0; k < 2; ++k) { D.nbr_ptr[k].assign(1, 0); D.nbr_ptr[k].reserve(n + 1); D.nbr_split[k].clear(); D.nbr_split[k].reserve(n); D.nbr[k].clear(); }
            D.max_fibre_len = 0;
            NbrCache * cache = caches ? &(*caches)[t] : nullptr;
            if (cache && cache->entries > ((size_t)1 << 24)) cache->clear();   // bound: 16 M entries (128 MB) per dimension
            NbrTemplate scratch;
            std::vector<int> sig;
            std::vector<const NbrTemplate *> & fibre_tmpl = D.fibre_tmpl;
            fibre_tmpl.assign(D.n_fibre, nullptr);
            auto append = [&D](const NbrTemplate * T, int m)
            {
                for (int k = 0; k < 2; ++k)
                {
                    D.nbr[k].insert(D.nbr[k].end(), T->nbr[k].begin(), T->nbr[k].end());
                    int64_t p = D.nbr_ptr[k].back();
                    for (int i = 0; i < m; ++i) { p += T->cnt[k][i]; D.nbr_ptr[k].push_back(p); }
                    D.nbr_split[k].insert(D.nbr_split[k].end(), T->nu[k].begin(), T->nu[k].end());
                }
            };
            for (int64_t f = 0; f < D.n_fibre; ++f)
            {
                const int64_t s0 = D.fibre_ptr[f], s1 = D.fibre_ptr[f + 1];
                const int m = (int)(s1 - s0);
                D.max_fibre_len = std::max(D.max_fibre_len, m);
                sig.resize(m);
                uint64_t h = 1469598103934665603ull;
                for (int i = 0; i < m; ++i) { sig[i] = o[(int64_t)perm[s0 + i] * d + t]; h = (h ^ (uint64_t)sig[i]) * 1099511628211ull; }
                const NbrTemplate * T = nullptr;
                if (cache)
                {
                    auto range = cache->map.equal_range(h);
                    for (auto it = range.first; it != range.second; ++it) if (it->second.sig == sig) { T = &it->second; break; }
                    if (T) cache->hits++; else cache->misses++;
                }
                if (!T)
                {
                    NbrTemplate & N = scratch;
                    N.sig = sig;
                    for (int k = 0; k < 2; ++k) { N.nbr[k].clear(); N.cnt[k].assign(m, 0); N.nu[k].assign(m, 0); }
                    for (int i = 0; i < m; ++i) pos_of_ord[sig[i]] = i;
                    for (int i = 0; i < m; ++i)
                    {
                        const int oe = sig[i];
                        const int le = lev_of[oe];
                        auto visit = [&](int of, int pair)
                        {
                            const int local = pos_of_ord[of];
                            const bool is_u = lev_of[of] <= le;
                            N.nbr[1].push_back({ local, pair }); N.cnt[1][i]++; if (is_u) N.nu[1][i]++;
                            if (P.vol[pair]) { N.nbr[0].push_back({ local, pair }); N.cnt[0][i]++; if (is_u) N.nu[0][i]++; }
                        };
                        const int c0 = P.tgt_ptr[oe], c1 = P.tgt_ptr[oe + 1];
                        if (c1 - c0 <= m)
                        {
                            for (int c = c0; c < c1; ++c) if (pos_of_ord[P.src[c]] >= 0) visit(P.src[c], c);
                        }
                        else
                        {
                            for (int r = 0; r < m; ++r)
                            {
                                const int pair = P.id[(size_t)sig[r] * P.T + oe];
                                if (pair >= 0) visit(sig[r], pair);
                            }
                        }
                        // sources were visited in ascending 1D order, and order grows with level: U entries lead
                    }
                    for (int i = 0; i < m; ++i) pos_of_ord[sig[i]] = -1;
                    if (cache)
                    {
                        cache->entries += N.nbr[0].size() + N.nbr[1].size() + (size_t)m;
                        T = &cache->map.emplace(h, std::move(N))->second;
                        scratch = NbrTemplate();
                    }
                    else T = &N;
                }
                fibre_tmpl[f] = T;
                if (!cache) append(T, m);
            }
            if (!cache) fibre_tmpl.clear();
            if (cache)
            {
                // unordered_multimap never moves its nodes: the pointers collected above are still valid
                size_t tot[2] = { 0, 0 };
                for (int64_t f = 0; f < D.n_fibre; ++f) for (int k = 0; k < 2; ++k) tot[k] += fibre_tmpl[f]->nbr[k].size();
                for (int k = 0; k < 2; ++k) D.nbr[k].reserve(tot[k]);
                for (int64_t f = 0; f < D.n_fibre; ++f) append(fibre_tmpl[f], (int)(D.fibre_ptr[f + 1] - D.fibre_ptr[f]));
            }
                    return 0;
        };
        std::vector<int> rc(dim, 0);
        if (dim > 1 && n >= 2048)
        {
            std::vector<std::thread> th;
            for (int t = 0; t < dim; ++t) th.emplace_back([&, t]() { rc[t] = build_dim(t); });
            for (auto & x : th) x.join();
        }
        else for (int t = 0; t < dim; ++t) rc[t] = build_dim(t);
        for (int t = 0; t < dim; ++t) if (rc[t] != 0) return -1;
        return 0;
    }
};

// ---- DGSolution initial grid (source/DGSolution.cpp:10-57) ---------------------------------------------------
inline int64_t sparse_grid(int dim, int level_init, bool sparse, std::vector<int> * level, std::vector<int> * suppt)
{
    int64_t count = 0;
    std::vector<int> n(dim, 0), j(dim, 0), jmax(dim, 1);
    while (true)
    {
        int sum = 0; for (int t = 0; t < dim; ++t) sum += n[t];
        if (!(sparse && sum > level_init))
        {
            for (int t = 0; t < dim; ++t) { jmax[t] = n[t] == 0 ? 1 : (1 << (n[t] - 1)); j[t] = 0; }
            while (true)
            {
                if (level) for (int t = 0; t < dim; ++t) { level->push_back(n[t]); suppt->push_back(2 * j[t] + 1); }
                ++count;
                int t = dim - 1;
                while (t >= 0 && ++j[t] == jmax[t]) { j[t] = 0; --t; }
                if (t < 0) break;
            }
        }
        int t = dim - 1;
        while (t >= 0 && ++n[t] == level_init + 1) { n[t] = 0; --t; }
        if (t < 0) break;
    }
    return count;
}

// ---- DGSolution with auxiliary dimensions (source/DGSolution.cpp:59-116): full grid of level `level_init` in the first dim - aux_dim dimensions,
// level 0 in the remaining ones -- the grid of a field (E, B) that lives with f in one phase space (example/07_vlasov_*.cpp) ------------------------
inline int64_t aux_grid(int dim, int level_init, int aux_dim, std::vector<int> * level, std::vector<int> * suppt)
{
    const int real_dim = dim - aux_dim;
    int64_t count = 0;
    std::vector<int> n(dim, 0), j(dim, 0), jmax(dim, 1);
    while (true)
    {
        for (int t = 0; t < dim; ++t) { jmax[t] = n[t] == 0 ? 1 : (1 << (n[t] - 1)); j[t] = 0; }
        while (true)
        {
            if (level) for (int t = 0; t < dim; ++t) { level->push_back(n[t]); suppt->push_back(2 * j[t] + 1); }
            ++count;
            int t = dim - 1;
            while (t >= 0 && ++j[t] == jmax[t]) { j[t] = 0; --t; }
            if (t < 0) break;
        }
        int t = real_dim - 1;                       // the auxiliary dimensions stay at level 0
        while (t >= 0 && ++n[t] == level_init + 1) { n[t] = 0; --t; }
        if (t < 0) break;
    }
    return count;
}

}  // namespace amdg
