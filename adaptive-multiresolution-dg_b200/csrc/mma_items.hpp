// Host-side "tile programs" of the tensor-core sweep kernel (kernels.cu: sweep_mma_kernel).
//
// Along dimension t every fibre is a set of 1D elements (its 1D orders, ascending) -- its SHAPE.  All fibres of one
// shape have the same neighbour structure and use the same operator blocks, so the sweep restricted to the
// fibres of a shape is one small block-sparse matrix M (m*KT x m*KF) applied to many right-hand sides
// (fibres x columns).  M is cut into 8x4 tiles for the FP64 tensor-core MMA m8n8k4:
//     rows  = TG consecutive targets x KTP padded output indices (KTP = KT rounded up to a divisor of 8, TG = 8 / KTP)
//     cols  = 4 source indices of one source element (NKP = ceil(KF/4) column tiles per source)
// A ROW TILE keeps the list of column tiles in which any of its targets has an operator block; the kernel walks
// that list with the A fragments (operator values, per operator) streaming from L1/L2 and the B fragments (staged
// source coefficients) from shared memory.  Nothing assumes complete binary trees: a shape is whatever set of 1D
// elements the (possibly adaptive) grid has on a fibre.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

#include "grid.hpp"

namespace amdg {

inline int mma_ktp(int kt) { return kt <= 1 ? 1 : (kt <= 2 ? 2 : (kt <= 4 ? 4 : 8)); }

struct ShapeTable
{
    std::map<std::vector<int>, int> id_of;          // signature (1D orders) -> shape id
    std::vector<std::vector<int>> ords;             // per shape
    // per dimension: shape of every fibre and the fibres of every shape (as first slots)
    std::vector<std::vector<int>> fibre_shape;       // [dim][fibre]
    std::vector<std::map<int, std::vector<int>>> shape_fibres;   // [dim][shape] -> slot0 list

    // shape ids are stable while a shape is known: a grid change (DGAdapt refine / coarsen) only appends the shapes it
    // introduces, so that plans and operator fragments cached per shape id survive it.  retire() forgets shapes the current
    // grid no longer has (their ids are never reused; a shape that comes back gets a fresh id), which bounds the table over a long adaptive run.
    std::vector<char> live_flags() const
    {
        std::vector<char> live(ords.size(), 0);
        for (const auto & per_dim : shape_fibres) for (const auto & kv : per_dim) live[kv.first] = 1;
        return live;
    }
    size_t n_known() const { return id_of.size(); }
    uint64_t epoch = 1;   // memoised ids in the neighbour templates (grid.hpp) are valid for this epoch only
    void retire(const std::vector<char> & live)
    {
        ++epoch;
        for (auto it = id_of.begin(); it != id_of.end();)
        {
            if (!live[it->second]) { std::vector<int>().swap(ords[it->second]); it = id_of.erase(it); } else ++it;
        }
    }
    void build(const Grid & G)
    {
        fibre_shape.assign(G.dim, std::vector<int>()); shape_fibres.assign(G.dim, std::map<int, std::vector<int>>());
        for (int t = 0; t < G.dim; ++t)
        {
            const DimTables & H = G.dims[t];
            fibre_shape[t].resize(H.n_fibre);
            for (int64_t f = 0; f < H.n_fibre; ++f)
            {
                const NbrTemplate * T = H.fibre_tmpl.empty() ? nullptr : H.fibre_tmpl[f];
                int id;
                if (T && T->shape_epoch == epoch && T->shape_id >= 0) id = T->shape_id;
                else
                {
                    std::vector<int> sig;
                    for (int64_t s = H.fibre_ptr[f]; s < H.fibre_ptr[f + 1]; ++s) sig.push_back(G.ord1d[(int64_t)H.slot_elem[s] * G.dim + t]);
                    auto it = id_of.find(sig);
                    if (it == id_of.end()) { id = (int)ords.size(); id_of.emplace(sig, id); ords.push_back(sig); } else id = it->second;
                    if (T) { T->shape_id = id; T->shape_epoch = epoch; }
                }
                fibre_shape[t][f] = id;
                shape_fibres[t][id].push_back((int)H.fibre_ptr[f]);
            }
        }
    }
};

// tile program of one (shape, relation, L/U/full, KF, KT)
struct ShapeProg
{
    int m = 0, n_rt = 0, tg = 1, nkp = 1, ktp = 1;
    std::vector<int> rt_ptr;        // [n_rt+1] into the entries
    std::vector<int> rt_order;      // row tile ids of the piece's row tiles (explicit ids; a whole program lists all of them, longest first)
    std::vector<int> ent_src;       // per entry: source local index * nkp + kp
    std::vector<int> ent_pair;      // per entry: [tg] canonical pair id of (source, target) or -1
    int64_t n_ent() const { return (int64_t)ent_src.size(); }
};

// lu: 0 = L (sources of strictly higher 1D level), 1 = U (same or lower level), 2 = full
inline void build_shape_prog(const Pairs1D & P1, const std::vector<int> & ords, int rel, int lu, int kf, int kt, ShapeProg & out)
{
    const int m = (int)ords.size();
    out.m = m; out.ktp = mma_ktp(kt); out.tg = 8 / out.ktp; out.nkp = (kf + 3) / 4;
    out.n_rt = (m + out.tg - 1) / out.tg;
    out.rt_ptr.assign(1, 0); out.ent_src.clear(); out.ent_pair.clear();
    std::vector<int> lev(m);
    for (int e = 0; e < m; ++e) lev[e] = level_of_order(ords[e]);
    for (int rt = 0; rt < out.n_rt; ++rt)
    {
        for (int f = 0; f < m; ++f)
        {
            int pairs[8]; bool any = false;
            for (int g = 0; g < out.tg; ++g)
            {
                pairs[g] = -1;
                const int e = rt * out.tg + g;
                if (e >= m) continue;
                const int pr = P1.id[(size_t)ords[f] * P1.T + ords[e]];
                if (pr < 0 || (rel == 0 && !P1.vol[pr])) continue;
                const bool is_u = lev[f] <= lev[e];
                if ((lu == 0 && is_u) || (lu == 1 && !is_u)) continue;
                pairs[g] = pr; any = true;
            }
            if (!any) continue;
            for (int kp = 0; kp < out.nkp; ++kp)
            {
                out.ent_src.push_back(f * out.nkp + kp);
                for (int g = 0; g < out.tg; ++g) out.ent_pair.push_back(pairs[g]);
            }
        }
        out.rt_ptr.push_back((int)out.ent_src.size());
    }
    out.rt_order.resize(out.n_rt);
    for (int i = 0; i < out.n_rt; ++i) out.rt_order[i] = i;
    std::stable_sort(out.rt_order.begin(), out.rt_order.end(), [&](int a, int b) { return out.rt_ptr[a + 1] - out.rt_ptr[a] > out.rt_ptr[b + 1] - out.rt_ptr[b]; });
}

// Split a program into `np` pieces with balanced entry counts (longest-processing-time first).  A piece is a program
// over a subset of the row tiles: rt_order holds their ids, rt_ptr/ent_* are piece-local.
inline void split_shape_prog(const ShapeProg & S, int np, std::vector<ShapeProg> & pieces)
{
    pieces.assign(np, ShapeProg());
    std::vector<int64_t> load(np, 0);
    std::vector<std::vector<int>> rts(np);
    // S.rt_order is sorted by decreasing length; rt_ptr of a whole program is indexed by row tile id
    for (int ri = 0; ri < S.n_rt; ++ri)
    {
        const int rt = S.rt_order[ri];
        int best = 0; for (int q = 1; q < np; ++q) if (load[q] < load[best]) best = q;
        rts[best].push_back(rt); load[best] += S.rt_ptr[rt + 1] - S.rt_ptr[rt] + 2;
    }
    for (int q = 0; q < np; ++q)
    {
        ShapeProg & P = pieces[q];
        P.m = S.m; P.tg = S.tg; P.nkp = S.nkp; P.ktp = S.ktp; P.n_rt = (int)rts[q].size();
        P.rt_ptr.assign(1, 0); P.rt_order = rts[q];
        for (int rt : rts[q])
        {
            for (int p = S.rt_ptr[rt]; p < S.rt_ptr[rt + 1]; ++p)
            {
                P.ent_src.push_back(S.ent_src[p]);
                for (int g = 0; g < S.tg; ++g) P.ent_pair.push_back(S.ent_pair[(size_t)p * S.tg + g]);
            }
            P.rt_ptr.push_back((int)P.ent_src.size());
        }
    }
}

// operator values of a program in MMA fragment order: A[entry][lane], lane l <-> row l/4 = (target g, output q), col l%4 = source index
inline void build_shape_A(const ShapeProg & S, const double * blocks, int kf, int kt, std::vector<double> & A)
{
    A.assign((size_t)S.n_ent() * 32, 0.0);
    for (int64_t p = 0; p < S.n_ent(); ++p)
    {
        const int kp = S.ent_src[p] % S.nkp;
        for (int l = 0; l < 32; ++l)
        {
            const int r = l >> 2, kk = l & 3;
            const int g = r / S.ktp, q = r % S.ktp, k = kp * 4 + kk;
            if (g >= S.tg || q >= kt || k >= kf) continue;
            const int pr = S.ent_pair[p * S.tg + g];
            if (pr < 0) continue;
            A[(size_t)p * 32 + l] = blocks[((size_t)pr * kf + k) * kt + q];
        }
    }
}

}  // namespace amdg
