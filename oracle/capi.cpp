// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Plain-C surface over the restatement so tests/ and bench.py can drive it through ctypes.
#include "oracle_tools.hpp"
#include "u_mer.hpp"
#include <cstring>
#include <stdexcept>

using namespace orc;
namespace orc {
void extendSearchNaive(const Searcher&, OffsetT, OffsetT, OffsetT, const Query&, OffsetT&, OffsetT&, OffsetT&);
void extendSearchBrute(const Searcher&, OffsetT, OffsetT, OffsetT, const Query&, OffsetT&, OffsetT&, OffsetT&);
int lce(const Searcher&, OffsetT, OffsetT, OffsetT, OffsetT);
bool collect(const Searcher&, const char*, int, std::vector<QuasiAlignment>&,
             std::vector<SAIntervalHit>&, std::vector<SAIntervalHit>&, MateStatus, bool);
void build_sa_sais(const std::string&, std::vector<OffsetT>&);
void build_sa_compare(const std::string&, std::vector<OffsetT>&, int);
void mergeLeftRightHitsFuzzy(bool, bool, std::vector<QuasiAlignment>&, std::vector<QuasiAlignment>&,
                             std::vector<QuasiAlignment>&, uint32_t, uint32_t, bool&);
}

static thread_local std::string g_err;
struct Run { WtResult wt; BsjResult bsj; std::string e1, e2; std::vector<uint64_t> eo1, eo2; };

#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_index_from_text(const char* seq, uint64_t n, int k, int nthreads) {
    ORC_TRY
    auto* ix = new Index();
    try { index_from_text(*ix, seq, n, k, nthreads); } catch (...) { delete ix; throw; }
    return ix;
    ORC_CATCH(nullptr)
}
void* orc_index_from_text_sa(const char* seq, uint64_t n, int k, const int32_t* sa, int nthreads) {
    ORC_TRY
    auto* ix = new Index();
    try { index_from_text_with_sa(*ix, seq, n, k, sa, nthreads); } catch (...) { delete ix; throw; }
    return ix;
    ORC_CATCH(nullptr)
}
// the suffix array alone: algo 0 = induced sorting (u_sais.cpp, what the index builder uses), 1 = bucketed comparison sort
int orc_suffix_array(const char* seq, uint64_t n, int algo, int nthreads, int32_t* out) {
    ORC_TRY
    std::string s(seq, n);
    std::vector<OffsetT> SA;
    if (algo == 0) build_sa_sais(s, SA); else build_sa_compare(s, SA, nthreads);
    std::memcpy(out, SA.data(), n * sizeof(int32_t));
    return 0;
    ORC_CATCH(-1)
}
// FASTA records (names, sequences as NUL-terminated strings) with SURVEY B6 normalisation
void* orc_index_from_records(const char** names, const char** seqs, uint32_t nrec, int k, int nthreads) {
    ORC_TRY
    std::vector<std::pair<std::string, std::string>> recs;
    for (uint32_t i = 0; i < nrec; ++i) recs.emplace_back(names[i], seqs[i]);
    auto* ix = new Index();
    try { index_from_records(*ix, recs, k, nthreads); } catch (...) { delete ix; throw; }
    return ix;
    ORC_CATCH(nullptr)
}
void orc_index_free(void* h) { delete (Index*)h; }
uint64_t orc_index_text_len(void* h) { return ((Index*)h)->seq.size(); }
uint32_t orc_index_num_tx(void* h) { return (uint32_t)((Index*)h)->txpOffsets.size(); }
uint64_t orc_index_num_kmers(void* h) { return ((Index*)h)->khash.size; }
const char* orc_index_text(void* h) { return ((Index*)h)->seq.data(); }
const int32_t* orc_index_sa(void* h) { return ((Index*)h)->SA.data(); }
const uint32_t* orc_index_offsets(void* h) { return ((Index*)h)->txpOffsets.data(); }
const uint32_t* orc_index_lens(void* h) { return ((Index*)h)->txpLens.data(); }
const char* orc_index_name(void* h, uint32_t t) { return ((Index*)h)->txpNames[t].c_str(); }
void orc_index_set_name(void* h, uint32_t t, const char* nm) { ((Index*)h)->txpNames[t] = nm; }
uint32_t orc_index_tid_at(void* h, uint64_t p) { return ((Index*)h)->transcriptAtPosition(p); }

int orc_khash_find(void* h, const char* kmer, int32_t* b, int32_t* e) {
    Index* ix = (Index*)h;
    auto s = ix->khash.find(mer_from_chars(kmer, ix->k));
    if (!s) return 0;
    *b = s->begin; *e = s->end;
    return 1;
}

int orc_extend(void* h, int brute, int32_t lbIn, int32_t ubIn, int startAt, const char* read, int L, int start,
               int rc, int32_t* out3) {
    Searcher ss{(Index*)h};
    Query q{read, L, start, rc != 0};
    OffsetT lb, ub, len;
    if (brute) extendSearchBrute(ss, lbIn, ubIn, startAt, q, lb, ub, len);
    else extendSearchNaive(ss, lbIn, ubIn, startAt, q, lb, ub, len);
    out3[0] = lb; out3[1] = ub; out3[2] = len;
    return 0;
}

int orc_lce(void* h, int mode, int32_t p1, int32_t p2, int startAt, int stopAt) {
    Searcher ss{(Index*)h}; ss.lceMode = mode;
    return lce(ss, p1, p2, startAt, stopAt);
}

// SACollectorCircall::operator() on one read.  Intervals come back as 4 ints each
// (begin, end, len, queryPos).  Returns foundHit (0/1) or -1 when a capacity is too small.
int orc_collect(void* h, const char* read, int len, int strict, int lceMode,
                int32_t* fwd, int* nfwd, int32_t* rc, int* nrc,
                uint32_t* hit_tid, int32_t* hit_pos, uint8_t* hit_fwd, int* nhits, int cap_ints, int cap_hits) {
    Searcher ss{(Index*)h}; ss.lceMode = lceMode;
    std::vector<QuasiAlignment> hits; std::vector<SAIntervalHit> f, r;
    bool found = collect(ss, read, len, hits, f, r, SINGLE_END, strict != 0);
    if ((int)f.size() > cap_ints || (int)r.size() > cap_ints || (int)hits.size() > cap_hits) return -1;
    for (size_t i = 0; i < f.size(); ++i) { fwd[4*i] = f[i].begin; fwd[4*i+1] = f[i].end; fwd[4*i+2] = f[i].len; fwd[4*i+3] = (int32_t)f[i].queryPos; }
    for (size_t i = 0; i < r.size(); ++i) { rc[4*i] = r[i].begin; rc[4*i+1] = r[i].end; rc[4*i+2] = r[i].len; rc[4*i+3] = (int32_t)r[i].queryPos; }
    for (size_t i = 0; i < hits.size(); ++i) { hit_tid[i] = hits[i].tid; hit_pos[i] = hits[i].pos; hit_fwd[i] = hits[i].fwd; }
    *nfwd = (int)f.size(); *nrc = (int)r.size(); *nhits = (int)hits.size();
    return found ? 1 : 0;
}

// mode: 0 = Circall_wt only, 1 = Circall_bsj only (r1 = read mapped, r2 ignored), 2 = wt then bsj on the exported records
void* orc_run(void* txh, void* bsjh, int mode, const char* r1, const uint64_t* off1, const char* r2,
              const uint64_t* off2, uint64_t n, int nthreads, int lceMode) {
    ORC_TRY
    auto* R = new Run();
    ReadBatch rb{r1, off1, r2, off2, n};
    try {
        if (mode == 0 || mode == 2) run_wt(*(Index*)txh, rb, nthreads, lceMode, R->wt);
        if (mode == 1) run_bsj(*(Index*)bsjh, rb, nthreads, lceMode, R->bsj);
        if (mode == 2) {
            export_records(rb, R->wt, R->e1, R->eo1, R->e2, R->eo2);
            ReadBatch eb{R->e1.data(), R->eo1.data(), R->e2.data(), R->eo2.data(), R->wt.expPair.size()};
            run_bsj(*(Index*)bsjh, eb, nthreads, lceMode, R->bsj);
        }
    } catch (...) { delete R; throw; }
    return R;
    ORC_CATCH(nullptr)
}
void orc_run_free(void* r) { delete (Run*)r; }

// out[0..4] = numObserved, numMapped, numHits, upperBound, readLen ; out[5..9] = P, S, T, H, collector calls
void orc_run_counters(void* r, int which, uint64_t* out) {
    Run* R = (Run*)r;
    if (which == 0) { auto& w = R->wt; uint64_t v[10] = {w.numObserved, w.numMapped, w.numHits, w.upperBound, w.readLen, w.wc.P, w.wc.S, w.wc.T, w.wc.H, w.wc.calls}; memcpy(out, v, sizeof v); }
    else { auto& w = R->bsj; uint64_t v[10] = {w.numObserved, w.numMapped, w.numHits, w.upperBound, w.readLen, w.wc.P, w.wc.S, w.wc.T, w.wc.H, w.wc.calls}; memcpy(out, v, sizeof v); }
}
uint64_t orc_wt_n_export(void* r) { return ((Run*)r)->wt.expPair.size(); }
void orc_wt_export(void* r, uint64_t* pair, uint8_t* code) {
    auto& w = ((Run*)r)->wt;
    memcpy(pair, w.expPair.data(), w.expPair.size() * 8);
    memcpy(code, w.expCode.data(), w.expCode.size());
}
void orc_wt_mates(void* r, uint8_t* flags, uint16_t* nhits) {
    auto& w = ((Run*)r)->wt;
    memcpy(flags, w.mateFlags.data(), w.mateFlags.size());
    memcpy(nhits, w.mateNHits.data(), w.mateNHits.size() * 2);
}
uint64_t orc_bsj_n_junc(void* r) { return ((Run*)r)->bsj.junc.size(); }
// 6 int64 per line: rec, dir, queryPos, len, hitPos, tid
void orc_bsj_junc(void* r, int64_t* out) {
    auto& b = ((Run*)r)->bsj;
    for (size_t i = 0; i < b.junc.size(); ++i) {
        auto& j = b.junc[i];
        out[6*i] = (int64_t)j.rec; out[6*i+1] = j.dir; out[6*i+2] = j.queryPos; out[6*i+3] = j.len; out[6*i+4] = j.hitPos; out[6*i+5] = j.tid;
    }
}
uint64_t orc_bsj_n_cand(void* r) { return ((Run*)r)->bsj.cand.size(); }
void orc_bsj_cand(void* r, uint64_t* out) { auto& b = ((Run*)r)->bsj; memcpy(out, b.cand.data(), b.cand.size() * 8); }
uint64_t orc_bsj_n_rec(void* r) { return ((Run*)r)->bsj.recNHits.size(); }
void orc_bsj_recs(void* r, uint32_t* nhits, uint8_t* split) {
    auto& b = ((Run*)r)->bsj;
    memcpy(nhits, b.recNHits.data(), b.recNHits.size() * 4);
    memcpy(split, b.recSplit.data(), b.recSplit.size());
}
// equivalence classes, flattened: sizes[0] = number of classes, sizes[1] = total tids
void orc_eq_sizes(void* r, int which, uint64_t* sizes) {
    auto& eq = which == 0 ? ((Run*)r)->wt.eq : ((Run*)r)->bsj.eq;
    uint64_t nt = 0; for (auto& kv : eq) nt += kv.first.size();
    sizes[0] = eq.size(); sizes[1] = nt;
}
void orc_eq_fill(void* r, int which, uint32_t* lens, uint32_t* tids, uint64_t* counts) {
    auto& eq = which == 0 ? ((Run*)r)->wt.eq : ((Run*)r)->bsj.eq;
    size_t c = 0, t = 0;
    for (auto& kv : eq) {
        lens[c] = (uint32_t)kv.first.size(); counts[c] = kv.second; ++c;
        for (auto x : kv.first) tids[t++] = x;
    }
}

// mate-hit intersection as Circall_pseudo.cpp:271-280 runs it (Fuzzy variant,
// maxNumHits = 200).  Inputs are the two mates' tid-sorted hit lists; output joint hits as
// 7 int32 each: tid, pos, fwd, matePos, mateIsFwd, fragLen, mateStatus.
int orc_merge_fuzzy(int leftMatches, int rightMatches, const uint32_t* ltid, const int32_t* lpos, const uint8_t* lfwd, int nl,
                    const uint32_t* rtid, const int32_t* rpos, const uint8_t* rfwd, int nr, uint32_t readLen,
                    uint32_t maxNumHits, int32_t* out, int cap, int* tooMany) {
    std::vector<QuasiAlignment> L(nl), R(nr), J;
    for (int i = 0; i < nl; ++i) { L[i].tid = ltid[i]; L[i].pos = lpos[i]; L[i].fwd = lfwd[i]; L[i].readLen = readLen; L[i].mateStatus = PAIRED_END_LEFT; }
    for (int i = 0; i < nr; ++i) { R[i].tid = rtid[i]; R[i].pos = rpos[i]; R[i].fwd = rfwd[i]; R[i].readLen = readLen; R[i].mateStatus = PAIRED_END_RIGHT; }
    bool tm = false;
    mergeLeftRightHitsFuzzy(leftMatches != 0, rightMatches != 0, L, R, J, readLen, maxNumHits, tm);
    if ((int)J.size() > cap) return -1;
    for (size_t i = 0; i < J.size(); ++i) {
        out[7*i] = (int32_t)J[i].tid; out[7*i+1] = J[i].pos; out[7*i+2] = J[i].fwd; out[7*i+3] = J[i].matePos;
        out[7*i+4] = J[i].mateIsFwd; out[7*i+5] = (int32_t)J[i].fragLen; out[7*i+6] = (int32_t)J[i].mateStatus;
    }
    *tooMany = tm;
    return (int)J.size();
}

} // extern "C"

// libstdc++ std::sort on entries compared by (entry >> 8) only — the reference behaviour the product's
// StdSort replica is tested against (SACollectorCircall.hpp:450; SURVEY §7 hard part 2).
#include <algorithm>
namespace { struct KsEntry { uint32_t v; bool operator<(const KsEntry& o) const { return (v >> 8) < (o.v >> 8); } }; }
extern "C" void orc_stdsort(uint32_t* a, int n) {
    std::vector<KsEntry> v(n);
    for (int i = 0; i < n; ++i) v[i].v = a[i];
    std::sort(v.begin(), v.end());
    for (int i = 0; i < n; ++i) a[i] = v[i].v;
}
