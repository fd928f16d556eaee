// circall_b200 — per-mate / per-record bodies of the two tools, on top of cg_core.cuh.
//   wt_mate()     one mate of  src/Circall_wt.cpp:275-365 / 283-540   
//   bsj_record()  one record of src/Circall_bsj.cpp:300-399           
// Written as CG_HD so the kernels are thin wrappers and tests can run the same code on the host.
#pragma once
#include "cg_core.cuh"

namespace cg {

static const int MAXC = 1000;         // an interval kept by the collector spans < maxInterval = 1000
static const int MAX_READ_OCCS = 200; // sfOpts.maxReadOccs (Circall_wt.cpp:1515)

// per-thread scratch; lives in local memory on the device
template <int MAXINT, int MAXKS, bool WITH_POS>
struct Scratch {
    CollectState<MAXINT, MAXKS> st;
    uint32_t tidF[MAXC], tidR[MAXC];
    uint8_t actF[MAXC], actR[MAXC];
    uint32_t posF[WITH_POS ? MAXC : 1], posR[WITH_POS ? MAXC : 1];
    uint16_t qpF[WITH_POS ? MAXC : 1], qpR[WITH_POS ? MAXC : 1];
    int nhF, nhR;
};

// run the collector and turn each strand's intervals into its tid-sorted hit list
template <int MAXINT, int MAXKS, bool WITH_POS>
CG_HD bool collect_hits(const IndexView& ix, const ReadView& rv, bool strict, int lceMode,
                        Scratch<MAXINT, MAXKS, WITH_POS>& sc) {
    sc.nhF = sc.nhR = 0;
    bool found = collect<MAXINT, MAXKS>(ix, rv, strict, lceMode, sc.st);
    if (!found) return false;
    HitScratch<WITH_POS> hf; hf.tid = sc.tidF; hf.act = sc.actF; hf.pos = sc.posF; hf.qp = sc.qpF;
    HitScratch<WITH_POS> hr; hr.tid = sc.tidR; hr.act = sc.actR; hr.pos = sc.posR; hr.qp = sc.qpR;
    sc.nhF = hit_tids<WITH_POS>(ix, sc.st.fwd, sc.st.nF, hf, MAXC, sc.st.err);
    sc.nhR = hit_tids<WITH_POS>(ix, sc.st.rc, sc.st.nR, hr, MAXC, sc.st.err);
    return true;
}

// size of the fwd/rc lists merged by tid, fwd first on ties, duplicates dropped (:567-579).
// emit(tid, fromFwd, indexInList) is called for every surviving hit in tid order.
template <int MAXINT, int MAXKS, bool WITH_POS, class Emit>
CG_HD int merged_hits(const Scratch<MAXINT, MAXKS, WITH_POS>& sc, Emit emit) {
    int i = 0, j = 0, n = 0;
    while (i < sc.nhF || j < sc.nhR) {
        if (j >= sc.nhR || (i < sc.nhF && sc.tidF[i] <= sc.tidR[j])) {
            uint32_t t = sc.tidF[i];
            emit(t, true, i);
            ++i;
            if (j < sc.nhR && sc.tidR[j] == t) ++j;
        } else {
            emit(sc.tidR[j], false, j);
            ++j;
        }
        ++n;
    }
    return n;
}
struct NoEmit { CG_HD void operator()(uint32_t, bool, int) const {} };

enum { MATE_SEEDED = 1, MATE_FULL = 2, MATE_ERR = 0x80 };

// one mate of Circall_wt: flags bit0 = isKmerMapped (:311,486), bit1 = !isUnmapped (:318-342,493-517,
// compared against mate 1's length — `readLen` is set once per pair at :242); nhits = hits.size()
template <int MAXINT, int MAXKS>
CG_HD void wt_mate(const IndexView& ix, const ReadView& rv, int pairReadLen, int lceMode,
                   Scratch<MAXINT, MAXKS, false>& sc, uint8_t& flags, uint32_t& nhits) {
    flags = 0; nhits = 0;
    bool found = collect_hits<MAXINT, MAXKS, false>(ix, rv, true, lceMode, sc);
    if (sc.st.err) flags |= MATE_ERR;
    if (!found) return;
    if (sc.st.nF + sc.st.nR > 0) flags |= MATE_SEEDED;
    bool full = false;
    for (int i = 0; i < sc.st.nF; ++i) if ((int)sc.st.fwd[i].len == pairReadLen) { full = true; break; }
    for (int i = 0; i < sc.st.nR && !full; ++i) if ((int)sc.st.rc[i].len == pairReadLen) full = true;
    if (full) flags |= MATE_FULL;
    nhits = (uint32_t)merged_hits(sc, NoEmit());
}

// one record of Circall_bsj.  sink.line(dir, queryPos, len, hitPos, tid) receives every junction-
// table row; tids[0..min(nhits,200)) receives jointHits' transcript ids in order.
template <int MAXINT, int MAXKS, class Sink>
CG_HD void bsj_record(const IndexView& ix, const ReadView& rv, int lceMode, Scratch<MAXINT, MAXKS, false>& sc,
                      Sink& sink, uint32_t& nhits, bool& isSplit, uint32_t* tids, uint8_t& err) {
    nhits = 0; isSplit = false;
    bool found = collect_hits<MAXINT, MAXKS, false>(ix, rv, true, lceMode, sc);
    err = (uint8_t)sc.st.err;
    if (!found) return;
    int nt = 0;
    nhits = (uint32_t)merged_hits(sc, [&](uint32_t t, bool, int) { if (nt < MAX_READ_OCCS) tids[nt] = t; ++nt; });
    if (nhits == 0) return;                                                         // :319
    for (int dir = 0; dir < 2; ++dir) {
        const Interval* ints = dir == 0 ? sc.st.fwd : sc.st.rc;
        int n = dir == 0 ? sc.st.nF : sc.st.nR;
        for (int x = 0; x < n; ++x) {
            const Interval it = ints[x];
            for (uint32_t i0 = it.begin; i0 < it.end; i0 += 4) {                    // :326 / :355
                int nn = (int)cmin<uint32_t>(4u, it.end - i0);
                uint32_t tt[4], ll[4], o0[4], o1[4];
                sa_tids4(ix, i0, nn, tt, ll);
#pragma unroll
                for (int z = 0; z < 4; ++z) if (z < nn) { o0[z] = CG_LDG(&ix.offsets[tt[z]]); o1[z] = CG_LDG(&ix.offsets[tt[z] + 1]); }
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    if (z < nn) {
                        uint32_t refLen = o1[z] - o0[z] - 1;
                        int hitPos = (int)ll[z];
                        int junctPos = (int)(refLen / 2);                           // :333
                        if (hitPos < junctPos - 5 && hitPos + (int)it.len > junctPos + 5) { // :336 / :364
                            sink.line((uint32_t)dir, it.qpos, it.len, hitPos, tt[z]);
                            isSplit = true;
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path: the same collector with small fixed capacities so that all of its state fits in shared
// memory (no local-memory scratch).  A read that needs more than FI intervals per strand, FK k-mer
// scores, HC hit candidates, or keeps both strands after the vote is reported back (return false,
// nothing emitted) and redone by the general path above — results are identical either way.
// ---------------------------------------------------------------------------------------------
template <int FI, int FK, int HC>
struct FastScratch {
    CollectState<FI, FK> st;
    uint32_t tid[HC];
    uint8_t act[HC];
};

template <int FI, int FK, int HC>
CG_HD bool fast_collect_hits(const IndexView& ix, const ReadView& rv, int lceMode, FastScratch<FI, FK, HC>& sc, bool& found, int& nh) {
    nh = 0;
    found = collect<FI, FK>(ix, rv, true, lceMode, sc.st);
    if (sc.st.err) return false;
    if (!found) return true;
    if (sc.st.nF > 0 && sc.st.nR > 0) return false;
    HitScratch<false> hs; hs.tid = sc.tid; hs.act = sc.act; hs.pos = nullptr; hs.qp = nullptr;
    int err = 0;
    if (sc.st.nF) nh = hit_tids<false>(ix, sc.st.fwd, sc.st.nF, hs, HC, err);
    else if (sc.st.nR) nh = hit_tids<false>(ix, sc.st.rc, sc.st.nR, hs, HC, err);
    return err == 0;
}

template <int FI, int FK, int HC>
CG_HD bool wt_mate_fast(const IndexView& ix, const ReadView& rv, int pairReadLen, int lceMode, FastScratch<FI, FK, HC>& sc,
                        uint8_t& flags, uint32_t& nhits) {
    flags = 0; nhits = 0;
    bool found; int nh;
    if (!fast_collect_hits(ix, rv, lceMode, sc, found, nh)) return false;
    if (!found) return true;
    if (sc.st.nF + sc.st.nR > 0) flags |= MATE_SEEDED;
    bool full = false;
    for (int i = 0; i < sc.st.nF; ++i) if ((int)sc.st.fwd[i].len == pairReadLen) { full = true; break; }
    for (int i = 0; i < sc.st.nR && !full; ++i) if ((int)sc.st.rc[i].len == pairReadLen) full = true;
    if (full) flags |= MATE_FULL;
    nhits = (uint32_t)nh;
    return true;
}

// tids of jointHits are left in sc.tid[0..nhits)
template <int FI, int FK, int HC, class Sink>
CG_HD bool bsj_record_fast(const IndexView& ix, const ReadView& rv, int lceMode, FastScratch<FI, FK, HC>& sc, Sink& sink,
                           uint32_t& nhits, bool& isSplit) {
    nhits = 0; isSplit = false;
    bool found; int nh;
    if (!fast_collect_hits(ix, rv, lceMode, sc, found, nh)) return false;
    nhits = (uint32_t)nh;
    if (!found || nh == 0) return true;                                             // :319
    const int dir = sc.st.nF ? 0 : 1;
    const Interval* ints = dir == 0 ? sc.st.fwd : sc.st.rc;
    const int n = dir == 0 ? sc.st.nF : sc.st.nR;
    for (int x = 0; x < n; ++x) {
        const Interval it = ints[x];
        for (uint32_t i0 = it.begin; i0 < it.end; i0 += 4) {                        // :326 / :355
            int nn = (int)cmin<uint32_t>(4u, it.end - i0);
            uint32_t tt[4], ll[4], o0[4], o1[4];
            sa_tids4(ix, i0, nn, tt, ll);
#pragma unroll
            for (int z = 0; z < 4; ++z) if (z < nn) { o0[z] = CG_LDG(&ix.offsets[tt[z]]); o1[z] = CG_LDG(&ix.offsets[tt[z] + 1]); }
#pragma unroll
            for (int z = 0; z < 4; ++z) {
                if (z < nn) {
                    uint32_t refLen = o1[z] - o0[z] - 1;
                    int hitPos = (int)ll[z];
                    int junctPos = (int)(refLen / 2);                               // :333
                    if (hitPos < junctPos - 5 && hitPos + (int)it.len > junctPos + 5) { // :336 / :364
                        sink.line((uint32_t)dir, it.qpos, it.len, hitPos, tt[z]);
                        isSplit = true;
                    }
                }
            }
        }
    }
    return true;
}

} // namespace cg
