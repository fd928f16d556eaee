// TEST SCAFFOLDING — host emulation of the device code.
//
// Compiles circall_b200/csrc/cg_core.cuh + cg_pipeline.cuh with g++ (CG_HOST_EMU) and runs the very
// same per-thread functions the CUDA kernels call, one "thread" per read, on a CPU-only box.  It is
// how the device logic is checked against the oracle where no GPU exists (the `-m "not gpu"` tier);
// the `-m gpu` tests call libcircall_gpu.so itself.  It is not part of the product: nothing in
// circall_b200/ links or loads it, and libcircall_gpu.so has no CPU path.
//
// The index arrays are built here on the host with the same packing primitives the GPU build
// kernels use (make_block / text_kmer / hmix); the suffix array is passed in by the test.
#define CG_HOST_EMU 1
#include "../../circall_b200/csrc/cg_pipeline.cuh"
#include "../../circall_b200/csrc/cg_pre.cuh"
#include "../../circall_b200/csrc/cg_thr.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

using namespace cg;

struct EmuIndex {
    std::vector<u32x4> text, hslots;
    std::vector<uint32_t> sa, offsets, fbits, satid;
    std::vector<uint8_t> lcp;
    IndexView v;
    uint64_t n_kmers = 0;
};


// the front kernel's per-mate function on the host (false: declined)
static bool pre_host(const IndexView& ix, const std::vector<uint64_t>& pk, int L, int pairLen, std::vector<uint32_t>& scratch, uint8_t& fl, uint32_t& nh) {
    int W2f = std::max(1, (L + 31) / 32), WMf = (W2f + 1) / 2;
    for (int j = 0; j < WMf; ++j) {
        int lo = 64 * j;
        if (lo >= L) break;
        uint64_t mk = pk[W2f + j];
        if (L - lo < 64) mk &= (1ULL << (L - lo)) - 1;
        if (mk) return false;
    }
    cgt::TMem m; m.stride = 1; m.W2 = W2f + 1;
    scratch.assign(cgp::pre_words(m.W2), 0xdeadbeefu);
    m.base = scratch.data();
    for (int j = 0; j < W2f; ++j) { m.st(2 * j, (uint32_t)pk[j]); m.st(2 * j + 1, (uint32_t)(pk[j] >> 32)); }
    m.st(2 * W2f, 0); m.st(2 * W2f + 1, 0);
    uint8_t cls;
    return cgp::perfect_mate(ix, m, L, pairLen, fl, nh, cls);
}

// the lockstep thread tier's per-read functions on the host (false: declined).  As in the kernels, a read whose first k-mer
// and its reverse complement are both in the table goes to the two-strand (vote) instantiation, the rest to the lean one.
template <bool VOTE>
static bool thr_stage(const std::vector<uint64_t>& pk, int L, std::vector<uint32_t>& scratch, cgq::QMemT<VOTE>& m) {
    int W2f = std::max(1, (L + 31) / 32), WMf = (W2f + 1) / 2;
    for (int j = 0; j < WMf; ++j) {
        int lo = 64 * j;
        if (lo >= L) break;
        uint64_t mk = pk[W2f + j];
        if (L - lo < 64) mk &= (1ULL << (L - lo)) - 1;
        if (mk) return false;
    }
    if (L < 1) return false;
    m.stride = 1; m.W2 = W2f + 1;
    scratch.assign(cgq::QMemT<VOTE>::words(m.W2), 0xdeadbeefu);
    m.base = scratch.data();
    for (int j = 0; j < W2f; ++j) { m.st(2 * j, (uint32_t)pk[j]); m.st(2 * j + 1, (uint32_t)(pk[j] >> 32)); }
    m.st(2 * W2f, 0); m.st(2 * W2f + 1, 0);
    cgt::make_rc(m, L);
    return true;
}
// look-up class of a read as k_wt_pre / k_bsj_cls compute it: bit 0 / 1 = read[0,k) / its reverse complement in the table
static bool both_class(const IndexView& ix, const std::vector<uint64_t>& pk, int L) {
    if (L <= ix.k) return false;
    uint64_t mer0 = pk[0] & ix.km;
    return present(ix, mer0) && present(ix, revcomp(mer0, ix.k));
}
// the look-up class minus one, as k_wt_pre / k_bsj_cls leave it for the lockstep tier (which of the four end k-mers are in the table)
static int known_of(const IndexView& ix, const std::vector<uint64_t>& pk, int L) {
    if (L <= ix.k || getenv("CG_EMU_NO_KNOWN")) return -1;
    int W2f = std::max(1, (L + 31) / 32);
    std::vector<uint64_t> w(pk.begin(), pk.begin() + W2f); w.push_back(0);
    const int j = (L - ix.k) >> 5, sh = ((L - ix.k) & 31) * 2;
    const uint64_t mer0 = w[0] & ix.km, merL = (sh ? (w[j] >> sh) | (w[j + 1] << (64 - sh)) : w[j]) & ix.km;
    return (present(ix, mer0) ? 1 : 0) | (present(ix, revcomp(mer0, ix.k)) ? 2 : 0) | (present(ix, merL) ? 4 : 0) | (present(ix, revcomp(merL, ix.k)) ? 8 : 0);
}
static uint64_t g_thr_why[2][16];
static uint64_t g_thr_vote[2];
static std::vector<uint32_t> g_cost_log[2];      // per lockstep-tier call: class, ok, 8 cost counters  (0 wt, 1 bsj)
static void cost_begin() { memset(cgq::g_cost, 0, sizeof cgq::g_cost); }
static void cost_end(int which, int cls, bool ok) { auto& v = g_cost_log[which]; v.push_back((uint32_t)cls); v.push_back(ok); for (int i = 0; i < 8; ++i) v.push_back(cgq::g_cost[i]); }
static bool thr_wt_host(const IndexView& ix, const std::vector<uint64_t>& pk, int L, int pairLen, int lceMode, std::vector<uint32_t>& scratch, uint8_t& fl, uint32_t& nh) {
    int why = cgq::QD_N, nF = 0, nR = 0;
    bool ok;
    cost_begin();
    struct CostEnd { const IndexView& ix; const std::vector<uint64_t>& pk; int L; bool* ok; ~CostEnd() { cost_end(0, known_of(ix, pk, L) + (both_class(ix, pk, L) ? 16 : 0), *ok); } } ce{ix, pk, L, &ok};
    if (both_class(ix, pk, L) && !getenv("CG_EMU_NO_VOTE")) {
        cgq::QMemT<true> m;
        ok = thr_stage<true>(pk, L, scratch, m) && cgq::q_wt_mate<true>(ix, m, L, pairLen, lceMode, fl, nh, why, nF, nR, known_of(ix, pk, L));
        if (ok) ++g_thr_vote[0];
    } else {
        cgq::QMemT<false> m;
        ok = thr_stage<false>(pk, L, scratch, m) && cgq::q_wt_mate<false>(ix, m, L, pairLen, lceMode, fl, nh, why, nF, nR, known_of(ix, pk, L));
    }
    if (!ok) ++g_thr_why[0][(-why) & 15];
    return ok;
}
template <class Sink>
static bool thr_bsj_host(const IndexView& ix, const std::vector<uint64_t>& pk, int L, int lceMode, std::vector<uint32_t>& scratch, Sink& sink, uint32_t& nh, bool& split, uint32_t* tids) {
    int why = cgq::QD_N, at = 0, nF = 0, nR = 0;
    bool ok;
    cost_begin();
    struct CostEnd { const IndexView& ix; const std::vector<uint64_t>& pk; int L; bool* ok; ~CostEnd() { cost_end(1, known_of(ix, pk, L) + (both_class(ix, pk, L) ? 16 : 0), *ok); } } ce{ix, pk, L, &ok};
    if (both_class(ix, pk, L) && !getenv("CG_EMU_NO_VOTE")) {
        cgq::QMemT<true> m;
        ok = thr_stage<true>(pk, L, scratch, m) && cgq::q_bsj_record<true>(ix, m, L, lceMode, sink, nh, split, at, why, nF, nR, known_of(ix, pk, L));
        if (ok) { ++g_thr_vote[1]; for (uint32_t t = 0; t < nh && t < (uint32_t)MAX_READ_OCCS; ++t) tids[t] = m.ld(at + t); }
    } else {
        cgq::QMemT<false> m;
        ok = thr_stage<false>(pk, L, scratch, m) && cgq::q_bsj_record<false>(ix, m, L, lceMode, sink, nh, split, at, why, nF, nR, known_of(ix, pk, L));
        if (ok) for (uint32_t t = 0; t < nh && t < (uint32_t)MAX_READ_OCCS; ++t) tids[t] = m.ld(at + t);
    }
    if (!ok) ++g_thr_why[1][(-why) & 15];
    return ok;
}

extern "C" {

void* emu_index_create(const char* text, uint64_t n, const int32_t* sa, int k) {
    auto* E = new EmuIndex();
    for (uint64_t i = 0; i < n; ++i) if (text[i] == '$') E->offsets.push_back(0);
    uint32_t ntx = (uint32_t)E->offsets.size();
    E->offsets.assign(ntx + 1, 0);
    std::vector<uint32_t> ends;
    { uint32_t t = 0, st = 0; for (uint64_t i = 0; i < n; ++i) if (text[i] == '$') { E->offsets[t++] = st; ends.push_back((uint32_t)i); st = (uint32_t)i + 1; } }
    E->offsets[ntx] = (uint32_t)n;
    uint64_t nb = (n + 63) / 64 + 1;
    E->text.resize(2 * nb);
    for (uint64_t b = 0; b < nb; ++b) {
        uint64_t c0, c1, mask; bool bad;
        unsigned char buf[64];
        int nv = (int)std::min<int64_t>(64, std::max<int64_t>(0, (int64_t)n - (int64_t)b * 64));
        memset(buf, '$', 64);
        if (nv > 0) memcpy(buf, text + b * 64, nv);
        make_block(buf, nv, c0, c1, mask, bad);
        uint32_t tid0 = (uint32_t)(std::lower_bound(ends.begin(), ends.end(), (uint32_t)std::min<uint64_t>(b * 64, 0xffffffffu)) - ends.begin());
        uint32_t start0 = tid0 < ntx ? E->offsets[tid0] : (uint32_t)n;
        E->text[2 * b] = u32x4{(uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32)};
        E->text[2 * b + 1] = u32x4{(uint32_t)mask, (uint32_t)(mask >> 32), tid0, start0};
    }
    E->sa.assign(sa, sa + n);
    E->v.text = E->text.data(); E->v.sa = E->sa.data(); E->v.offsets = E->offsets.data();
    E->v.n = (uint32_t)n; E->v.n_tx = ntx; E->v.k = k; E->v.hslots = nullptr; E->v.hmask = 0; E->v.hshift = 0;
    E->v.km = kmask(k); E->v.km3 = E->v.km / 3;
    E->v.fbits = nullptr; E->v.fm = 0; E->v.satid = nullptr; E->v.lcp = nullptr;
    if (!getenv("CG_NO_LCP")) {        // min(255, lcp(SA[i-1], SA[i])), as cg_index.cu::build_lcp makes it
        E->lcp.assign(n, 0);
        for (uint64_t i = 1; i < n; ++i) E->lcp[i] = (uint8_t)lce(E->v, (uint32_t)i - 1, (uint32_t)i, 0, 255, 1);
        E->v.lcp = E->lcp.data();
    }
    if (!getenv("CG_NO_SATID")) {      // transcriptAtPosition(SA[i]) per suffix-array index, as cg_index.cu::build_satid makes it
        E->satid.resize(n);
        for (uint64_t i = 0; i < n; ++i) { uint32_t tid, local; tid_of(E->v, E->sa[i], tid, local); E->satid[i] = tid; }
        E->v.satid = E->satid.data();
    }
    {   // seed filter, as cg_index.cu::build_fbits makes it: every '$'-free fm-mer of the text and its reverse complement
        int want = 16;
        if (const char* e = getenv("CG_FM")) want = atoi(e);
        const int fm = fm_for_k(k, want);
        if (fm) {
            E->fbits.assign((size_t)((1ULL << (2 * fm)) / 32), 0u);
            for (uint64_t p = 0; p < n; ++p) {
                uint64_t w;
                if (!text_word(E->v, (uint32_t)p, fm, w)) continue;
                const uint32_t x = (uint32_t)w, y = fm_revcomp(x, fm);
                E->fbits[x >> 5] |= 1u << (x & 31); E->fbits[y >> 5] |= 1u << (y & 31);
            }
            E->v.fbits = E->fbits.data(); E->v.fm = fm;
        }
    }
    // k-mer table: heads of runs of equal valid k-mers in SA order
    std::vector<uint64_t> keys; std::vector<uint32_t> begins, endsI;
    bool have = false; uint64_t prev = 0;
    for (uint64_t j = 0; j <= n; ++j) {
        uint64_t w = 0; bool v = j < n && text_kmer(E->v, E->sa[j], w);
        bool same = v && have && w == prev;
        if (have && !same) { endsI.push_back((uint32_t)j); have = false; }
        if (v && !same) { keys.push_back(w); begins.push_back((uint32_t)j); have = true; }
        if (v) prev = w;
    }
    uint64_t cap = 16; int hshift = 60; while (cap < keys.size() * 4 + 2) { cap <<= 1; --hshift; }
    E->hslots.assign(cap, u32x4{0xffffffffu, 0xffffffffu, 0, 0});
    for (size_t i = 0; i < keys.size(); ++i) {
        uint64_t h = home_of(keys[i], hshift);
        while (!(E->hslots[h].x == 0xffffffffu && E->hslots[h].y == 0xffffffffu)) h = (h + 1) & (cap - 1);
        E->hslots[h] = u32x4{(uint32_t)keys[i], (uint32_t)(keys[i] >> 32), begins[i], endsI[i]};
    }
    E->v.hslots = E->hslots.data(); E->v.hmask = cap - 1; E->v.hshift = hshift; E->n_kmers = keys.size();
    return E;
}
void emu_index_free(void* h) { delete (EmuIndex*)h; }
uint64_t emu_index_num_kmers(void* h) { return ((EmuIndex*)h)->n_kmers; }

static bool pack_read(const char* s, int L, int W2f, int WMf, std::vector<uint64_t>& pk) {
    pk.assign(W2f + WMf, 0);
    bool anybad = false;
    for (int j = 0; j < W2f; ++j) {
        uint64_t code; uint32_t nm; bool bad;
        int nb = std::min(32, std::max(0, L - 32 * j));
        pack_word((const unsigned char*)s + 32 * j, nb, code, nm, bad);
        anybad |= bad;
        pk[j] = code;
        pk[W2f + j / 2] |= (uint64_t)nm << (32 * (j & 1));
    }
    if (W2f & 1) pk[W2f + W2f / 2] |= 0xffffffff00000000ULL;
    return !anybad;
}

static const int EMAXINT = 232, EMAXKS = 928;
static const int EFI = 4, EFK = 16, EHC = 24;     // fast-path capacities, as in cg_session.cu

// Same outputs as cg_collect.  Caller provides capacities; returns -1 on overflow, -4 on a bad base.
int emu_collect(void* h, const char* bases, const uint64_t* off, uint64_t n_reads, int strict, int lceMode,
                uint8_t* found, uint64_t* fwd_off, int32_t* fwd_ints, uint64_t* rc_off, int32_t* rc_ints,
                uint64_t cap_ints, uint64_t* hits_off, uint32_t* hit_tid, int32_t* hit_pos, uint8_t* hit_fwd,
                uint64_t cap_hits) {
    EmuIndex* E = (EmuIndex*)h;
    auto* sc = new Scratch<EMAXINT, EMAXKS, true>();
    std::vector<uint64_t> pk, strands;
    uint64_t nf = 0, nr = 0, nh = 0;
    int rcode = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        int L = (int)(off[r + 1] - off[r]);
        int W2f = (L + 31) / 32, WMf = (W2f + 1) / 2;
        if (W2f == 0) { W2f = 1; WMf = 1; }
        if (!pack_read(bases + off[r], L, W2f, WMf, pk)) { rcode = -4; break; }
        strands.assign(view_words(W2f, WMf), 0);
        ReadView rv = build_view(pk.data(), 1, L, W2f, WMf, strands.data());
        bool f = collect_hits<EMAXINT, EMAXKS, true>(E->v, rv, strict != 0, lceMode, *sc);
        if (sc->st.err) { rcode = -6; break; }
        found[r] = f;
        fwd_off[r] = nf; rc_off[r] = nr; hits_off[r] = nh;
        if (f) {
            if (nf + sc->st.nF > cap_ints || nr + sc->st.nR > cap_ints) { rcode = -1; break; }
            for (int i = 0; i < sc->st.nF; ++i, ++nf) { auto& I = sc->st.fwd[i]; fwd_ints[4*nf] = I.begin; fwd_ints[4*nf+1] = I.end; fwd_ints[4*nf+2] = I.len; fwd_ints[4*nf+3] = I.qpos; }
            for (int i = 0; i < sc->st.nR; ++i, ++nr) { auto& I = sc->st.rc[i]; rc_ints[4*nr] = I.begin; rc_ints[4*nr+1] = I.end; rc_ints[4*nr+2] = I.len; rc_ints[4*nr+3] = I.qpos; }
            bool ovf = false;
            merged_hits(*sc, [&](uint32_t t, bool fw, int i) {
                if (nh >= cap_hits) { ovf = true; return; }
                hit_tid[nh] = t; hit_pos[nh] = (int32_t)(fw ? sc->posF[i] : sc->posR[i]); hit_fwd[nh] = fw; ++nh;
            });
            if (ovf) { rcode = -1; break; }
        }
    }
    fwd_off[n_reads] = nf; rc_off[n_reads] = nr; hits_off[n_reads] = nh;
    delete sc;
    return rcode;
}

struct VecSink {
    std::vector<int64_t>* v; uint64_t rec;
    void line(uint32_t dir, uint32_t q, uint32_t len, int hitPos, uint32_t tid) {
        v->push_back((int64_t)rec); v->push_back(dir); v->push_back(q); v->push_back(len); v->push_back(hitPos); v->push_back(tid);
    }
};

struct EmuRun {
    std::vector<uint8_t> mateFlags; std::vector<uint16_t> mateNHits;
    std::vector<uint64_t> expPair; std::vector<uint8_t> expCode;
    std::vector<int64_t> junc; std::vector<uint64_t> cand;
    std::vector<uint32_t> recNHits; std::vector<uint8_t> recSplit;
    std::vector<uint32_t> eqLens, eqTids;     // one entry per candidate record
    uint64_t ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t fallbacks[2] = {0, 0};      // reads the fast path handed to the general path (wt mates, bsj records)
    uint64_t pre_done = 0;               // wt mates resolved by the thread-per-mate front kernel's function
    uint64_t thr_done[2] = {0, 0};       // wt mates / bsj records resolved by the lockstep thread tier
    int err = 0;
};

static bool view_of(const char* s, int L, std::vector<uint64_t>& pk, std::vector<uint64_t>& strands, ReadView& rv) {
    int W2f = std::max(1, (L + 31) / 32), WMf = (W2f + 1) / 2;
    bool ok = pack_read(s, L, W2f, WMf, pk);
    strands.assign(view_words(W2f, WMf), 0);
    rv = build_view(pk.data(), 1, L, W2f, WMf, strands.data());
    return ok;
}


// fused wt -> bsj on the host, mirroring what the session's kernels + commit step compute
void* emu_run(void* txh, void* bsjh, const char* r1, const uint64_t* off1, const char* r2, const uint64_t* off2,
              uint64_t n, int lceMode) {
    EmuIndex* TX = (EmuIndex*)txh; EmuIndex* BS = (EmuIndex*)bsjh;
    auto* R = new EmuRun();
    auto* sc = new Scratch<EMAXINT, EMAXKS, false>();
    auto* fs = new FastScratch<EFI, EFK, EHC>();
    std::vector<uint64_t> pk, strands;
    std::vector<uint32_t> tscr;
    const bool use_pre = getenv("CG_EMU_NO_PRE") == nullptr;
    const bool use_thr = getenv("CG_EMU_NO_THR") == nullptr;
    for (uint64_t i = 0; i < n; ++i) {
        int l1 = (int)(off1[i + 1] - off1[i]), l2 = (int)(off2[i + 1] - off2[i]);
        uint8_t fl[2]; uint32_t nh[2];
        for (int m = 0; m < 2; ++m) {
            ReadView rv;
            if (!view_of(m ? r2 + off2[i] : r1 + off1[i], m ? l2 : l1, pk, strands, rv)) R->err = -4;
            // same tiers as the kernels: clean-match front function, lockstep thread tier, small-capacity fast path, general path on overflow
            if (use_pre && pre_host(TX->v, pk, m ? l2 : l1, l1, tscr, fl[m], nh[m])) ++R->pre_done;
            else if (use_thr && thr_wt_host(TX->v, pk, m ? l2 : l1, l1, lceMode, tscr, fl[m], nh[m])) ++R->thr_done[0];
            else if (!wt_mate_fast<EFI, EFK, EHC>(TX->v, rv, l1, lceMode, *fs, fl[m], nh[m])) {
                ++R->fallbacks[0];
                wt_mate<EMAXINT, EMAXKS>(TX->v, rv, l1, lceMode, *sc, fl[m], nh[m]);
            }
            if (fl[m] & MATE_ERR) R->err = -6;
            R->mateFlags.push_back(fl[m] & 3);
        }
        // pair-level counters (Circall_wt.cpp:369-372,467-470,544-547,642-645)
        R->ctr[3] += (nh[0] > 0) + (nh[1] > 0);
        uint32_t a = nh[0] > (uint32_t)MAX_READ_OCCS ? 0 : nh[0], b = nh[1] > (uint32_t)MAX_READ_OCCS ? 0 : nh[1];
        R->mateNHits.push_back((uint16_t)a); R->mateNHits.push_back((uint16_t)b);
        R->ctr[1] += (a > 0) + ((a > 0) || (b > 0));
        R->ctr[2] += a + b;
        R->ctr[0] += 2;
        for (int m = 0; m < 2; ++m)
            if ((fl[m] & MATE_SEEDED) && !(fl[m] & MATE_FULL)) { R->expPair.push_back(i); R->expCode.push_back((uint8_t)(m + 1)); }
    }
    if (BS) {
        uint32_t tids[MAX_READ_OCCS];
        for (uint64_t rec = 0; rec < R->expPair.size(); ++rec) {
            uint64_t i = R->expPair[rec]; int m = R->expCode[rec] - 1;
            ReadView rv;
            view_of(m ? r2 + off2[i] : r1 + off1[i], (int)(m ? off2[i + 1] - off2[i] : off1[i + 1] - off1[i]), pk, strands, rv);
            VecSink sink{&R->junc, rec};
            uint32_t nh; bool split; uint8_t err = 0;
            std::vector<int64_t> tmpj; VecSink tsink{&tmpj, rec};
            int Lr = (int)(m ? off2[i + 1] - off2[i] : off1[i + 1] - off1[i]);
            if (use_thr && thr_bsj_host(BS->v, pk, Lr, lceMode, tscr, tsink, nh, split, tids)) {
                ++R->thr_done[1];
                R->junc.insert(R->junc.end(), tmpj.begin(), tmpj.end());
            } else if (bsj_record_fast<EFI, EFK, EHC>(BS->v, rv, lceMode, *fs, sink, nh, split)) {
                for (uint32_t t = 0; t < nh && t < (uint32_t)MAX_READ_OCCS; ++t) tids[t] = fs->tid[t];
            } else {
                ++R->fallbacks[1];
                bsj_record<EMAXINT, EMAXKS>(BS->v, rv, lceMode, *sc, sink, nh, split, tids, err);
            }
            if (err) R->err = -6;
            R->recNHits.push_back(nh); R->recSplit.push_back(split);
            R->ctr[7] += nh > 0;
            uint32_t c = nh > (uint32_t)MAX_READ_OCCS ? 0 : nh;
            bool cand = c > 0 && split;
            if (cand) { R->cand.push_back(rec); R->eqLens.push_back(c); for (uint32_t t = 0; t < c; ++t) R->eqTids.push_back(tids[t]); }
            R->ctr[5] += cand; R->ctr[6] += c; R->ctr[4] += 1;
        }
    }
    delete sc; delete fs;
    return R;
}
void emu_run_free(void* r) { delete (EmuRun*)r; }
int emu_run_err(void* r) { return ((EmuRun*)r)->err; }
void emu_run_fallbacks(void* r, uint64_t* out) { out[0] = ((EmuRun*)r)->fallbacks[0]; out[1] = ((EmuRun*)r)->fallbacks[1]; }
// which tier answered: [0] wt mates by the clean-match front function, [1] wt mates / [2] bsj records by the lockstep thread tier
void emu_run_tiers(void* r, uint64_t* out) { EmuRun* R = (EmuRun*)r; out[0] = R->pre_done; out[1] = R->thr_done[0]; out[2] = R->thr_done[1]; }
void emu_run_sizes(void* r, uint64_t* out) {
    EmuRun* R = (EmuRun*)r;
    out[0] = R->expPair.size(); out[1] = R->junc.size() / 6; out[2] = R->cand.size(); out[3] = R->eqTids.size();
    out[4] = R->mateFlags.size();
}
void emu_run_fill(void* r, uint8_t* mateFlags, uint16_t* mateNHits, uint64_t* expPair, uint8_t* expCode, int64_t* junc,
                  uint64_t* cand, uint32_t* recNHits, uint8_t* recSplit, uint32_t* eqLens, uint32_t* eqTids, uint64_t* ctr) {
    EmuRun* R = (EmuRun*)r;
    auto cp = [](void* d, const void* s, size_t b) { if (b) memcpy(d, s, b); };
    cp(mateFlags, R->mateFlags.data(), R->mateFlags.size()); cp(mateNHits, R->mateNHits.data(), R->mateNHits.size() * 2);
    cp(expPair, R->expPair.data(), R->expPair.size() * 8); cp(expCode, R->expCode.data(), R->expCode.size());
    cp(junc, R->junc.data(), R->junc.size() * 8); cp(cand, R->cand.data(), R->cand.size() * 8);
    cp(recNHits, R->recNHits.data(), R->recNHits.size() * 4); cp(recSplit, R->recSplit.data(), R->recSplit.size());
    cp(eqLens, R->eqLens.data(), R->eqLens.size() * 4); cp(eqTids, R->eqTids.data(), R->eqTids.size() * 4);
    memcpy(ctr, R->ctr, sizeof R->ctr);
}

void emu_thr_why(uint64_t* out, int reset) { memcpy(out, g_thr_why, sizeof g_thr_why); out[15] = g_thr_vote[0]; out[31] = g_thr_vote[1]; if (reset) { memset(g_thr_why, 0, sizeof g_thr_why); memset(g_thr_vote, 0, sizeof g_thr_vote); } }

// seed filter of an index (0 = none) and what the lockstep tier's walks cost since the last reset: filter look-ups, table look-ups, windows the filter settled
int emu_index_fm(void* h) { return ((EmuIndex*)h)->v.fm; }
void emu_walk_stats(uint64_t* out, int reset) { out[0] = cgq::g_walk.filter_loads; out[1] = cgq::g_walk.table_loads; out[2] = cgq::g_walk.dead_windows; if (reset) cgq::g_walk = cgq::WalkStats(); }
// per-call cost log of the lockstep tier (which: 0 wt, 1 bsj): 10 uint32 per call; returns the number of values, copies when out is given, clears when asked
uint64_t emu_cost_log(int which, uint32_t* out, int clear) { auto& v = g_cost_log[which]; uint64_t n = v.size(); if (out) memcpy(out, v.data(), n * 4); if (clear) v.clear(); return n; }
// geometry of the walks' filter stage (tests sweep it; cg_thr.cuh)
void emu_walk_geometry(int np, int d0, int sp, int tail) { cgq::Q_NPR = np; cgq::Q_D0X = d0; cgq::Q_SPX = sp; cgq::Q_TAIL = tail; }
// the same with the number of walks, their table rounds and the histogram of rounds per walk (out: 6 + 24 values)
void emu_walk_stats_ex(uint64_t* out, int reset) { const auto& g = cgq::g_walk; out[0] = g.filter_loads; out[1] = g.table_loads; out[2] = g.dead_windows; out[3] = g.calls; out[4] = g.rounds; out[5] = g.found; for (int i = 0; i < 24; ++i) out[6 + i] = g.hist[i]; if (reset) cgq::g_walk = cgq::WalkStats(); }

// std::sort replica check: sorts packed kmerScores entries with cg::StdSort
void emu_stdsort(uint32_t* a, int n) { StdSort s; s.a = a; s.sort<true>(n); }

} // extern "C"
