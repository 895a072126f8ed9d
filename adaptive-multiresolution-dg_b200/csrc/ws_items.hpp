// Host-side plans and work lists of the warp-specialised streaming sweep kernel (kernels_ws.cu: sweep_ws_kernel).
//
// The kernel runs one persistent CTA per SM: a producer warp streams the source rows of ITEM n+1, n+2, ... into a ring of shared-memory
// stages (bulk copies, one instruction per contiguous run; 8-byte async copies where block sizes are odd) while eight consumer warps
// run the FP64 tensor-core MMAs of item n out of shared memory and store from registers.  An item is a PIECE of a fibre shape (a set
// of row tiles of the shape's tile program, mma_items.hpp, together with the union of the sources they read) x a run of fibres of that
// shape x a column rectangle, sized so that its staged rows fit one stage.  Row tiles whose source list is too long to stage (the
// coarse targets of long fibres) are HEAVY items: one row tile x one 8-column tile, not staged; all consumer warps split the entry
// list and add their partial sums through shared memory in warp order.
// Staged rows keep the element's own memory order restricted to the rectangle: slot[o_l][k][i_l]; all index arithmetic of a column
// tile is a look-up in per-rectangle-shape tables (B fragment offsets inside a slot, C fragment offsets in the destination block).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <queue>
#include <vector>

#include "grid.hpp"
#include "mma_items.hpp"

namespace amdg {

static const int WS_SRC_MAX = 24;        // sources staged per piece at most; a row tile that needs more is a heavy item
static const int WS_RT_MAX = 8;          // row tiles per piece at most

struct WsPiece
{
    std::vector<int> rts;            // row tile ids of the shape's program
    std::vector<int> src;            // staged sources (fibre-local indices, ascending); empty for a heavy piece
    ShapeProg prog;                  // entries row tile by row tile; ent_src = slot * 2 + k-part (slot = index into src; fibre-local index if heavy)
    bool heavy = false;
    long long hash = 0;
};

inline void build_ws_plan(const Pairs1D & P1, const std::vector<int> & ords, int nmax, int rel, int lu, int kf, int kt, std::vector<WsPiece> & out)
{
    ShapeProg SP; build_shape_prog(P1, ords, rel, lu, kf, kt, SP);
    out.clear();
    auto make = [&](const std::vector<int> & rts, bool heavy)
    {
        out.emplace_back();
        WsPiece & P = out.back();
        P.heavy = heavy;
        std::vector<int> loc(SP.m, -1);
        if (!heavy)
        {
            for (int rt : rts) for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p) loc[SP.ent_src[p] / SP.nkp] = 0;
            for (int f = 0; f < SP.m; ++f) if (loc[f] == 0) { loc[f] = (int)P.src.size(); P.src.push_back(f); }
        }
        // longest row tiles first: the consumer warps take sub-units round-robin
        std::vector<int> order(rts);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return SP.rt_ptr[a + 1] - SP.rt_ptr[a] > SP.rt_ptr[b + 1] - SP.rt_ptr[b]; });
        P.rts = order;
        ShapeProg & Q = P.prog;
        Q.m = SP.m; Q.tg = SP.tg; Q.nkp = SP.nkp; Q.ktp = SP.ktp; Q.n_rt = (int)order.size(); Q.rt_ptr.assign(1, 0); Q.rt_order = order;
        for (int rt : order)
        {
            for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p)
            {
                const int f = SP.ent_src[p] / SP.nkp, kp = SP.ent_src[p] % SP.nkp;
                Q.ent_src.push_back((heavy ? f : loc[f]) * 2 + kp);
                for (int g = 0; g < SP.tg; ++g) Q.ent_pair.push_back(SP.ent_pair[(size_t)p * SP.tg + g]);
            }
            Q.rt_ptr.push_back((int)Q.ent_src.size());
        }
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&](long long v) { h ^= (unsigned long long)v; h *= 1099511628211ull; };
        mix(SP.m); mix(Q.n_rt); mix(SP.tg); mix(SP.nkp); mix(heavy ? 0x77 : 0x33);
        for (int v : Q.rt_ptr) mix(v);
        for (int v : Q.ent_src) mix(v & 1);
        for (int v : Q.ent_pair) mix(v);
        P.hash = (long long)(h >> 1);
    };
    // build_shape_A reads the k-part as ent_src % nkp: entries carry slot*2 + kp, so nkp == 2 must hold there; ShapeProg::nkp of the
    // piece is forced to 2 below (k-part 1 never occurs when the source edge is <= 4)
    // depth-first order of the 1D tree: consecutive row tiles share their chain of ancestors
    std::vector<std::pair<int64_t, int>> key(SP.n_rt);
    for (int rt = 0; rt < SP.n_rt; ++rt)
    {
        const int o = ords[rt * SP.tg], n = level_of_order(o);
        const int64_t left = n <= 1 ? 0 : (int64_t)(o - (1 << (n - 1))) << (nmax - (n - 1));
        key[rt] = { left * 64 + n, rt };
    }
    std::sort(key.begin(), key.end());
    std::vector<int> cur; std::vector<char> mark(SP.m, 0); int cur_src = 0;
    auto flush = [&]() { if (!cur.empty()) { make(cur, false); cur.clear(); std::fill(mark.begin(), mark.end(), 0); cur_src = 0; } };
    for (auto & kr : key)
    {
        const int rt = kr.second;
        std::vector<int> own;
        for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p) own.push_back(SP.ent_src[p] / SP.nkp);
        std::sort(own.begin(), own.end()); own.erase(std::unique(own.begin(), own.end()), own.end());
        if ((int)own.size() > WS_SRC_MAX) { make({ rt }, true); continue; }
        int add = 0; for (int f : own) if (!mark[f]) ++add;
        if (!cur.empty() && (cur_src + add > WS_SRC_MAX || (int)cur.size() >= WS_RT_MAX)) { flush(); add = (int)own.size(); }
        for (int f : own) mark[f] = 1;
        cur_src += add; cur.push_back(rt);
    }
    flush();
    for (WsPiece & P : out) P.prog.nkp = 2;
}

// tables of one rectangle shape (no x ni columns of an outer x inner plane): B offsets inside a staged slot [o_l][k][i_l], C offsets in
// the destination block relative to the rectangle's origin; 8-column tiles, padded to a multiple of 8 tiles
inline void build_ws_tables(int no, int ni, int inner, int kf, int kt, std::vector<int> & tab_b, std::vector<int> & tab_c, int & nct, int & nct_pad, bool & vec_ok)
{
    const int W = no * ni, ktp = mma_ktp(kt);
    nct = (W + 7) / 8; nct_pad = (nct + 7) & ~7;
    tab_b.assign((size_t)nct_pad * 32, 0); tab_c.assign((size_t)nct_pad * 64, -1);
    vec_ok = true;
    for (int ct = 0; ct < nct_pad; ++ct)
        for (int lane = 0; lane < 32; ++lane)
        {
            const int kk = lane & 3, n = lane >> 2;
            const int c = std::min(ct * 8 + n, W - 1), o = c / ni, i = c % ni;
            tab_b[(size_t)ct * 32 + lane] = o * kf * ni + std::min(kk, kf - 1) * ni + i;
            const int q = (lane >> 2) % ktp;
            for (int h = 0; h < 2; ++h)
            {
                const int c2 = ct * 8 + 2 * (lane & 3) + h;
                if (c2 >= W || q >= kt) continue;
                const int o2 = c2 / ni, i2 = c2 % ni;
                tab_c[((size_t)ct * 32 + lane) * 2 + h] = o2 * kt * inner + q * inner + i2;
            }
            const int a0 = tab_c[((size_t)ct * 32 + lane) * 2], a1 = tab_c[((size_t)ct * 32 + lane) * 2 + 1];
            if (a0 >= 0 && !(a1 == a0 + 1 && (a0 & 1) == 0)) vec_ok = false;
        }
}

}  // namespace amdg
