// Host-side plans of the register-direct tensor-core sweep kernel (kernels_dir.cu: sweep_dir_kernel).
//
// The kernel has no shared-memory staging: a warp owns a UNIT = (piece of a fibre shape, a run of fibres of that shape, a range of
// 8-column tiles).  The piece lists a few ROW TILES of the shape's tile program (mma_items.hpp: TG targets x padded outputs) and the
// union of the sources those row tiles read; per source a bit mask says which row tiles have an operator block for it.  The warp
// walks the sources once per column group, loads each source's B fragments straight from global memory into registers and feeds
// them to the row tiles of its mask; accumulators stay in registers until the group is stored.  A source element is therefore read
// once per piece that needs it (coarse ancestors are re-read by the pieces below them, out of L1/L2), never staged.
//
// Variants (row tiles per piece x column tiles per group; the product is the number of accumulator tiles a warp holds):
//   0: 1 x 8   whole short fibres with one row tile
//   1: 2 x 4
//   2: 4 x 2
//   3: 1 x 1, entries spread round-robin over four accumulators: the few coarse row tiles of a long fibre whose source list is long
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

#include "grid.hpp"
#include "mma_items.hpp"

namespace amdg {

static const int DIR_HEAVY_ENT = 40;        // a row tile with more entries than this becomes a narrow piece of its own

struct DirPiece
{
    int variant = 0;
    int n_rt = 0;
    int rt_id[4] = { 0, 0, 0, 0 };  // row tiles of the shape's program (targets rt*TG .. rt*TG+TG-1)
    std::vector<int> src;           // per source entry: fibre-local source index * 2 + k-part
    std::vector<int> mask;          // per source entry: piece-local row tiles with an operator block for it
    ShapeProg prog;                 // entries in (source entry, row tile) order; only ent_src % nkp and ent_pair are used (build_shape_A)
    long long hash = 0;
    int n_ent() const { return (int)prog.ent_src.size(); }
};

inline int dir_variant_rt(int v) { return v == 1 ? 2 : (v == 2 ? 4 : 1); }
inline int dir_variant_g(int v) { return v == 0 ? 8 : (v == 1 ? 4 : (v == 2 ? 2 : 1)); }

// pieces of one (shape, relation, L/U/full, kf, kt); depends on nothing else (no shared-memory capacity enters)
inline void build_dir_plan(const Pairs1D & P1, const std::vector<int> & ords, int nmax, int rel, int lu, int kf, int kt, std::vector<DirPiece> & out)
{
    ShapeProg SP; build_shape_prog(P1, ords, rel, lu, kf, kt, SP);
    out.clear();
    auto make = [&](const std::vector<int> & rts, int variant)
    {
        out.emplace_back();
        DirPiece & P = out.back();
        P.variant = variant; P.n_rt = (int)rts.size();
        for (int r = 0; r < P.n_rt; ++r) P.rt_id[r] = rts[r];
        std::vector<int> codes;
        for (int rt : rts) for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p) codes.push_back(SP.ent_src[p]);
        std::sort(codes.begin(), codes.end()); codes.erase(std::unique(codes.begin(), codes.end()), codes.end());
        ShapeProg & Q = P.prog;
        Q.m = SP.m; Q.tg = SP.tg; Q.nkp = SP.nkp; Q.ktp = SP.ktp; Q.n_rt = P.n_rt;
        for (int code : codes)
        {
            const int f = code / SP.nkp, kp = code % SP.nkp;
            int mk = 0;
            for (int r = 0; r < P.n_rt; ++r)
            {
                const int rt = rts[r];
                for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p)
                    if (SP.ent_src[p] == code)
                    {
                        mk |= 1 << r;
                        Q.ent_src.push_back(kp);
                        for (int g = 0; g < SP.tg; ++g) Q.ent_pair.push_back(SP.ent_pair[(size_t)p * SP.tg + g]);
                    }
            }
            P.src.push_back(f * 2 + kp); P.mask.push_back(mk);
        }
        // content hash: key of the operator fragments of this piece
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&](long long v) { h ^= (unsigned long long)v; h *= 1099511628211ull; };
        mix(SP.m); mix(P.n_rt); mix(SP.tg); mix(SP.nkp); mix(0x5d1);
        for (int v : P.mask) mix(v);
        for (int v : Q.ent_src) mix(v);
        for (int v : Q.ent_pair) mix(v);
        P.hash = (long long)(h >> 1);
    };
    if (SP.n_rt == 1) { make({ 0 }, 0); return; }
    if (SP.n_rt == 2 && SP.n_ent() <= 2 * DIR_HEAVY_ENT) { make({ 0, 1 }, 1); return; }
    // longer fibres: row tiles in depth-first order of the 1D tree (left end of the support, then level) so that the row tiles of a
    // piece share their chain of ancestors; heavy row tiles (coarse targets, long source lists) stand alone
    std::vector<std::pair<int64_t, int>> key(SP.n_rt);
    for (int rt = 0; rt < SP.n_rt; ++rt)
    {
        const int o = ords[rt * SP.tg], n = level_of_order(o);
        const int64_t left = n <= 1 ? 0 : (int64_t)(o - (1 << (n - 1))) << (nmax - (n - 1));
        key[rt] = { left * 64 + n, rt };
    }
    std::sort(key.begin(), key.end());
    std::vector<int> cur;
    auto flush = [&]()
    {
        if (cur.empty()) return;
        make(cur, cur.size() == 1 ? 0 : (cur.size() == 2 ? 1 : 2));
        cur.clear();
    };
    for (auto & kr : key)
    {
        const int rt = kr.second;
        const int n_e = SP.rt_ptr[rt + 1] - SP.rt_ptr[rt];
        if (n_e > DIR_HEAVY_ENT) { make({ rt }, 3); continue; }
        cur.push_back(rt);
        if ((int)cur.size() == 4) flush();
    }
    flush();
}

// per-lane column offset tables of a block shape: B fragment source offsets and C fragment destination offsets of every 8-column
// tile.  B: lane = (k = lane % 4, column n = lane / 4); C: lane = (row = lane / 4 -> output q = row % KTP, columns 2*(lane%4), +1).
// Column c of the (outer x inner) plane is (o, i) = divmod(c, inner); source offset o*KF*inner + k*inner + i, destination offset
// o*KT*inner + q*inner + i.  Tiles are padded to a multiple of 8; columns beyond the plane load a valid (clamped) address and store nowhere (-1).
inline void build_dir_tables(int outer, int inner, int kf, int kt, std::vector<int> & tab_b, std::vector<int> & tab_c, int & nct_pad, bool & vec_ok)
{
    const int W = outer * inner, ktp = mma_ktp(kt);
    const int nct = (W + 7) / 8;
    nct_pad = (nct + 7) & ~7;
    tab_b.assign((size_t)nct_pad * 32, 0); tab_c.assign((size_t)nct_pad * 64, -1);
    vec_ok = true;
    for (int ct = 0; ct < nct_pad; ++ct)
        for (int lane = 0; lane < 32; ++lane)
        {
            const int kk = lane & 3, n = lane >> 2;
            const int c = std::min(ct * 8 + n, W - 1), o = c / inner, i = c % inner;
            tab_b[(size_t)ct * 32 + lane] = o * kf * inner + std::min(kk, kf - 1) * inner + i;
            const int q = (lane >> 2) % ktp;
            for (int h = 0; h < 2; ++h)
            {
                const int c2 = ct * 8 + 2 * (lane & 3) + h;
                if (c2 >= W || q >= kt) continue;
                const int o2 = c2 / inner, i2 = c2 % inner;
                tab_c[((size_t)ct * 32 + lane) * 2 + h] = o2 * kt * inner + q * inner + i2;
            }
            const int a0 = tab_c[((size_t)ct * 32 + lane) * 2], a1 = tab_c[((size_t)ct * 32 + lane) * 2 + 1];
            if (a0 >= 0 && !(a1 == a0 + 1 && (a0 & 1) == 0)) vec_ok = false;
        }
}

}  // namespace amdg
