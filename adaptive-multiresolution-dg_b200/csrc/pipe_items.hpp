// Host-side construction of the work list of the pipelined sweep kernel (kernels.cu: sweep_pipe_kernel).
//
// One work ITEM = a set of target rows of one sweep along dimension t, the source rows they need, the operator
// blocks of the distinct 1D pairs involved and the item-local neighbour lists -- everything the kernel stages in
// shared memory, serialised as one contiguous int RECORD so that it can be prefetched with bulk async copies:
//
//   [0] nsrc  [1] ntgt  [2] npair  [3] nnz  [4] col0  [5] ncol  [6] lcx  [7] pitch  [8] final_idx  [9] offset of ent (even)  [10..15] 0
//   src_elem[nsrc] | tgt_dest[ntgt] | rowptr[ntgt+1] | pair_id[npair] | ent[2*nnz] = (item-local source row, item-local pair)
//   (padded to a multiple of 4 ints)
//
//   tgt_dest >= 0 : element row of the destination array;  tgt_dest < 0 : partial-sum slot -(v+1).
//
// Short fibres are packed several to an item.  A fibre too long for one item is cut at a 1D level k_c: rows of level
// >= k_c are grouped by their level-k_c ancestor ("subtree" items, whose sources are the union of their neighbour
// lists).  The rows above the cut ("top" rows) have sources all over the fibre; their sums are split by the subtree
// the source lives in: every subtree item also produces, as extra rows, the PARTIAL sums of its sources for the top
// rows, one more item does the top-to-top part, and the last of these items to finish (atomic arrival counter per
// (fibre, column chunk)) adds the partials up -- the FINAL step, described by a small table.  Nothing assumes a
// complete binary tree: the grouping only bounds the footprints, the neighbour lists define the sums.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "grid.hpp"

namespace amdg {

enum { PIPE_HDR = 16, PIPE_LU_L = 0, PIPE_LU_U = 1, PIPE_LU_FULL = 2 };

struct PipeBuild
{
    // outputs
    std::vector<int> rec;                  // record pool
    std::vector<int> tab;                  // per item: offset, length (ints)
    std::vector<double> cost;              // per item
    std::vector<int> fin;                  // final table pool: per final: [expected, ntop, col0, ncol, then per top row: elem, nslot, slot...]
    std::vector<int> fin_ofs;              // per final: offset into fin
    int n_slot = 0;                        // partial-sum slots (rows of edge_to^... doubles per column)
    int max_data = 0, max_meta = 0;        // doubles / ints
    int ct = 1;
    int n_packed = 0, n_final = 0;
    bool ok = true;
};

struct PipeParams { int t, W, kf, kt, rel, lu, cap_doubles, meta_cap_ints, item_target, threads; };

inline int pipe_next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// shared-memory row pitch (doubles): even (16-byte vector reads of two adjacent columns) and such that the rows a
// warp touches together fall in distinct banks as far as possible
inline int pipe_pitch(int ncol, int kf, int cx, int ct)
{
    const int step = ct >= 2 ? 2 : 1;
    int base = (ncol + step - 1) / step * step;
    if (cx * ct >= 32) return base;
    int best = base, best_conf = 1 << 30;
    for (int P = base; P < base + 32; P += step)
    {
        // 16 lanes issue 8-byte (ct==1) or 16-byte (ct>=2: 8 lanes per phase) accesses; lanes = (row lane, column lane)
        const int lanes = ct >= 2 ? 8 : 16, unit = ct >= 2 ? 2 : 1, nb = ct >= 2 ? 8 : 16;
        int cnt[16] = { 0 }, conf = 0;
        for (int lane = 0; lane < lanes; ++lane) { const int tx = lane % cx, ty = lane / cx; cnt[((ty * kf * P + tx * ct) / unit) % nb]++; }
        for (int b = 0; b < nb; ++b) conf = std::max(conf, cnt[b]);
        if (conf < best_conf) { best_conf = conf; best = P; }
    }
    return best;
}

inline void build_pipe_list(const Grid & G, const Pairs1D & P1, const PipeParams & pp, PipeBuild & out)
{
    const int t = pp.t, W = pp.W, kf = pp.kf, kt = pp.kt, rel = pp.rel, lu = pp.lu, d = G.dim;
    const DimTables & H = G.dims[t];
    const std::vector<int64_t> & nptr = H.nbr_ptr[rel];
    const std::vector<Nbr> & nbr = H.nbr[rel];
    const std::vector<int> & nsplit = H.nbr_split[rel];
    const int ct = W > 128 ? 4 : (W > 16 ? 2 : 1);
    out.ct = ct;
    auto lcx_of = [&](int ncol) { int l = 0; while ((1 << l) < std::min(pp.threads, pipe_next_pow2((ncol + ct - 1) / ct))) l++; return l; };
    auto cols_of = [&](int j) { return std::max(1, (W + (1 << j) - 1) >> j); };
    const int MAXJ = 6;
    const int pitch_w = pipe_pitch(W, kf, 1 << lcx_of(W), ct);
    const int64_t total = G.n * (int64_t)kf * pitch_w;
    const int64_t pack_cap = std::max<int64_t>(std::min<int64_t>(pp.cap_doubles, total / std::max(1, pp.item_target)), (int64_t)kf * pitch_w);

    auto ord_of = [&](int64_t s) { return G.ord1d[(int64_t)H.slot_elem[s] * d + t]; };
    auto p_range = [&](int64_t s, int64_t & lo, int64_t & hi)
    {
        lo = (lu == PIPE_LU_L) ? nptr[s] + nsplit[s] : nptr[s];
        hi = (lu == PIPE_LU_U) ? nptr[s] + nsplit[s] : nptr[s + 1];
    };

    struct Item { std::vector<int> src, dest, rowptr, pairs, ent; double cost = 0; int final_idx = -1; };
    std::vector<int> sstamp(G.n, -1), slocal(G.n, 0), pstamp(P1.n_pairs, -1), plocal(P1.n_pairs, 0);
    int stamp = 0;
    Item B;
    auto begin_item = [&]() { B = Item(); B.rowptr.push_back(0); ++stamp; };
    // add a target row with the entries [lo,hi) of slot s's list restricted by `keep` (source slot -> bool); dest as given
    auto add_entries = [&](int dest, int64_t fibre_s0, int64_t lo, int64_t hi, const std::vector<char> * keep)
    {
        B.dest.push_back(dest);
        for (int64_t p = lo; p < hi; ++p)
        {
            const int ss = (int)(fibre_s0 + nbr[p].local), pr = nbr[p].pair;
            if (keep && !(*keep)[ss - fibre_s0]) continue;
            if (sstamp[ss] != stamp) { sstamp[ss] = stamp; slocal[ss] = (int)B.src.size(); B.src.push_back(H.slot_elem[ss]); }
            if (pstamp[pr] != stamp) { pstamp[pr] = stamp; plocal[pr] = (int)B.pairs.size(); B.pairs.push_back(pr); }
            B.ent.push_back(slocal[ss]); B.ent.push_back(plocal[pr]);
            B.cost += 1.0;
        }
        B.rowptr.push_back((int)B.ent.size() / 2);
        B.cost += 1.0;
    };
    auto add_row = [&](int64_t s, int64_t fibre_s0) { int64_t lo, hi; p_range(s, lo, hi); add_entries(H.slot_elem[s], fibre_s0, lo, hi, nullptr); };
    auto data_need = [&](int ncol) { const int lcx = lcx_of(ncol); return (int64_t)B.src.size() * kf * pipe_pitch(ncol, kf, 1 << lcx, ct) + 1 + (int64_t)B.pairs.size() * kf * kt; };
    auto meta_need = [&]() { return (int64_t)PIPE_HDR + (int64_t)B.src.size() + 2 * (int64_t)B.dest.size() + 1 + (int64_t)B.pairs.size() + (int64_t)B.ent.size() + 4; };
    auto fits = [&](int ncol, int64_t cap) { return data_need(ncol) <= cap && meta_need() <= pp.meta_cap_ints; };
    auto min_chunk = [&]() { for (int j = 0; j <= MAXJ; ++j) if (fits(cols_of(j), pp.cap_doubles)) return j; return -1; };
    // serialise the current item for every column chunk of level j; final_base >= 0: final index of chunk c is final_base + c
    auto emit = [&](int j, int final_base)
    {
        if (B.dest.empty()) return;
        const int ncol = cols_of(j), lcx = lcx_of(ncol), pitch = pipe_pitch(ncol, kf, 1 << lcx, ct);
        int chunk = 0;
        for (int c0 = 0; c0 < W; c0 += ncol, ++chunk)
        {
            const size_t ofs = out.rec.size();
            int ent_ofs = PIPE_HDR + (int)B.src.size() + 2 * (int)B.dest.size() + 1 + (int)B.pairs.size();
            const int ent_pad = ent_ofs & 1; ent_ofs += ent_pad;           // (source, pair) entries are read as 8-byte words
            const int hdr[PIPE_HDR] = { (int)B.src.size(), (int)B.dest.size(), (int)B.pairs.size(), (int)B.ent.size() / 2, c0, ncol, lcx, pitch,
                                        final_base >= 0 ? final_base + chunk : -1, ent_ofs, 0, 0, 0, 0, 0, 0 };
            out.rec.insert(out.rec.end(), hdr, hdr + PIPE_HDR);
            out.rec.insert(out.rec.end(), B.src.begin(), B.src.end());
            out.rec.insert(out.rec.end(), B.dest.begin(), B.dest.end());
            out.rec.insert(out.rec.end(), B.rowptr.begin(), B.rowptr.end());
            out.rec.insert(out.rec.end(), B.pairs.begin(), B.pairs.end());
            if (ent_pad) out.rec.push_back(0);
            out.rec.insert(out.rec.end(), B.ent.begin(), B.ent.end());
            while (out.rec.size() % 4) out.rec.push_back(0);
            out.tab.push_back((int)ofs); out.tab.push_back((int)(out.rec.size() - ofs));
            out.cost.push_back(B.cost * std::min(ncol, W - c0));
            out.max_meta = std::max(out.max_meta, (int)(out.rec.size() - ofs));
            out.max_data = std::max<int>(out.max_data, (int)((int64_t)B.src.size() * kf * pitch + 1 + (int64_t)B.pairs.size() * kf * kt));
            ++out.n_packed;
        }
        B = Item();
    };

    begin_item();
    std::vector<char> keep;
    for (int64_t f = 0; f < H.n_fibre && out.ok; ++f)
    {
        const int64_t s0 = H.fibre_ptr[f]; const int m = (int)(H.fibre_ptr[f + 1] - s0);
        // (1) the whole fibre into the running packed item (all columns)
        {
            Item saved = B; const int saved_stamp = stamp;
            for (int64_t s = s0; s < s0 + m; ++s) add_row(s, s0);
            if (fits(W, saved.dest.empty() ? pp.cap_doubles : pack_cap)) continue;
            B = saved; (void)saved_stamp;
            emit(0, -1); begin_item();
            for (int64_t s = s0; s < s0 + m; ++s) add_row(s, s0);
            if (fits(W, pp.cap_doubles)) continue;
            begin_item();
        }
        // (2) long fibre: cut at level kc, smallest kc (= largest subtrees) whose items fit with at most 2^MAXJ column chunks
        int lmax = 0; for (int64_t s = s0; s < s0 + m; ++s) lmax = std::max(lmax, level_of_order(ord_of(s)));
        bool done = false;
        for (int kc = 1; kc <= lmax + 1 && !done; ++kc)
        {
            std::map<int, std::vector<int64_t>> groups; std::vector<int64_t> top;
            std::vector<int> group_of(m, -1);
            for (int64_t s = s0; s < s0 + m; ++s)
            {
                const int o = ord_of(s), n = level_of_order(o);
                if (n < kc) { top.push_back(s); continue; }
                const int g = (o - (1 << (n - 1))) >> (n - kc);
                groups[g].push_back(s); group_of[s - s0] = g;
            }
            // which top rows have sources outside the top (those need partial sums)?
            bool need_partials = false;
            for (int64_t s : top) { int64_t lo, hi; p_range(s, lo, hi); for (int64_t p = lo; p < hi; ++p) if (group_of[nbr[p].local] >= 0) { need_partials = true; break; } if (need_partials) break; }
            // dry run: chunk level every item of this cut can live with
            int jmax = 0; bool ok_cut = true;
            std::vector<int> gkeys; for (auto & g : groups) gkeys.push_back(g.first);
            auto fill_group = [&](size_t gi, const std::vector<int> * slots)
            {
                begin_item();
                const std::vector<int64_t> & rows = groups[gkeys[gi]];
                for (int64_t s : rows) add_row(s, s0);
                if (need_partials)
                {
                    keep.assign(m, 0); for (int64_t s : rows) keep[s - s0] = 1;
                    for (size_t ti = 0; ti < top.size(); ++ti)
                    {
                        int64_t lo, hi; p_range(top[ti], lo, hi);
                        bool any = false; for (int64_t p = lo; p < hi; ++p) if (keep[nbr[p].local]) { any = true; break; }
                        if (!any) continue;
                        add_entries(slots ? -((*slots)[ti * (gkeys.size() + 1) + gi] + 1) : -1, s0, lo, hi, &keep);
                    }
                }
            };
            // the top rows go into items of at most TR rows each (item ti0: rows [ti0, ti0+TR))
            auto fill_top = [&](size_t ti0, size_t TR, const std::vector<int> * slots)
            {
                begin_item();
                const size_t ti1 = std::min(top.size(), ti0 + TR);
                if (!need_partials) { for (size_t ti = ti0; ti < ti1; ++ti) add_row(top[ti], s0); return; }
                keep.assign(m, 0); for (int64_t s : top) keep[s - s0] = 1;
                for (size_t ti = ti0; ti < ti1; ++ti)
                {
                    int64_t lo, hi; p_range(top[ti], lo, hi);
                    add_entries(slots ? -((*slots)[ti * (gkeys.size() + 1) + gkeys.size()] + 1) : -1, s0, lo, hi, &keep);
                }
            };
            for (size_t gi = 0; gi < gkeys.size() && ok_cut; ++gi) { fill_group(gi, nullptr); const int j = min_chunk(); if (j < 0) ok_cut = false; else jmax = std::max(jmax, j); }
            size_t TR = 32;
            if (ok_cut && !top.empty())
            {
                for (; TR >= 1; TR /= 2)
                {
                    bool all = true; int jt = 0;
                    for (size_t ti0 = 0; ti0 < top.size() && all; ti0 += TR) { fill_top(ti0, TR, nullptr); const int j = min_chunk(); if (j < 0) all = false; else jt = std::max(jt, j); }
                    if (all) { jmax = std::max(jmax, jt); break; }
                    if (TR == 1) { ok_cut = false; break; }
                }
            }
            begin_item();
            if (!ok_cut) continue;
            const int n_top_items = top.empty() ? 0 : (int)((top.size() + TR - 1) / TR);
            // commit this cut
            const int nchunk = (W + cols_of(jmax) - 1) / cols_of(jmax);
            std::vector<int> slots;
            int final_base = -1;
            if (need_partials)
            {
                // slot numbering: (top row, contributor) -> slot, only for contributors that have entries
                slots.assign(top.size() * (gkeys.size() + 1), -1);
                std::vector<std::vector<int>> contrib(top.size());
                for (size_t ti = 0; ti < top.size(); ++ti)
                {
                    int64_t lo, hi; p_range(top[ti], lo, hi);
                    std::vector<char> has(gkeys.size() + 1, 0);
                    for (int64_t p = lo; p < hi; ++p)
                    {
                        const int g = group_of[nbr[p].local];
                        if (g < 0) has[gkeys.size()] = 1;
                        else has[std::lower_bound(gkeys.begin(), gkeys.end(), g) - gkeys.begin()] = 1;
                    }
                    has[gkeys.size()] = 1;     // the top item always writes its slot (possibly an empty sum) so that every top row is covered
                    for (size_t gi = 0; gi <= gkeys.size(); ++gi) if (has[gi]) { slots[ti * (gkeys.size() + 1) + gi] = out.n_slot; contrib[ti].push_back(out.n_slot); ++out.n_slot; }
                }
                final_base = out.n_final;
                for (int chunk = 0; chunk < nchunk; ++chunk)
                {
                    out.fin_ofs.push_back((int)out.fin.size());
                    out.fin.push_back((int)gkeys.size() + n_top_items);   // expected arrivals: every subtree item + the top items (per column chunk)
                    out.fin.push_back((int)top.size());
                    out.fin.push_back(chunk * cols_of(jmax));
                    out.fin.push_back(cols_of(jmax));
                    for (size_t ti = 0; ti < top.size(); ++ti)
                    {
                        out.fin.push_back(H.slot_elem[top[ti]]);
                        out.fin.push_back((int)contrib[ti].size());
                        out.fin.insert(out.fin.end(), contrib[ti].begin(), contrib[ti].end());
                    }
                    ++out.n_final;
                }
            }
            for (size_t gi = 0; gi < gkeys.size(); ++gi) { fill_group(gi, need_partials ? &slots : nullptr); emit(jmax, final_base); }
            for (size_t ti0 = 0; ti0 < top.size(); ti0 += TR) { fill_top(ti0, TR, need_partials ? &slots : nullptr); emit(jmax, final_base); }
            begin_item();
            done = true;
        }
        if (!done) out.ok = false;
    }
    if (out.ok) emit(0, -1);
}

}  // namespace amdg
