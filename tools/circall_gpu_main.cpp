// circall_b200 — drop-in executables over libcircall_gpu.so (C ABI, include/circall_gpu.h).
//
// One binary, dispatched on its name (or on the first argument):
//   Circall_wt   -i TXINDEX  -1 R1.. -2 R2.. -o OUT [-p N] [--gpu D] [--gpus N]   src/Circall_wt.cpp main :1415-1739, called at Circall_source.sh:273
//   Circall_bsj  -i BSJINDEX -1 U1.. -2 U2.. -o OUT [-p N] [--gpu D]      src/Circall_bsj.cpp,         called at Circall_source.sh:291
//   Circall_pseudo -i PSEUDOINDEX -1 B1.. -2 B2.. -o OUT [-p N]           src/Circall_pseudo.cpp,      called at Circall_source.sh:340
//   TxIndexer    -t FASTA -o INDEX [-k 31] [-f]                           src/TxIndexer.cpp:81-94,192-227
//   circall_gpu fused -i TXINDEX -b BSJINDEX -1 .. -2 .. -o OUTWT -O OUTBSJ   (extension: both steps in one pass, no
//                intermediate Unmapped_readM_{1,2}.fa; the two output directories receive what the two tools write)
//   circall_gpu parse -1 .. -2 ..          (no GPU: parses the inputs and prints record counts; used by the CPU tests)
//   circall_gpu fld LEN..                  (no GPU: the fragment-length prior's fragmentInfo.txt columns and effective lengths)
// Options of the reference that do not apply (-l libtype, bias flags, ...) are accepted and ignored; a value that
// follows an unknown option is skipped.  Errors go to stderr and exit(1) like the reference (Circall_wt.cpp:1726-1737).
//
// The host side streams the reads into page-locked staging buffers and rotates over two sessions per GPU, so the copy
// and the kernels of one batch overlap the parsing and output writing of the others.  --gpus N (extension, SURVEY.md section 8 b / e)
// replicates the indices on N devices (--gpu D .. D + N - 1) and deals the batches round-robin over them; the batches are fetched
// in the order they were submitted, so the output files are the same as with one GPU, and the per-device counters meet in one
// all-reduce at the end of the run (cg_sessions_allreduce: NCCL over NVLink).
#include <algorithm>
#include <cstdlib>
#include <sys/stat.h>
#include "../include/circall_gpu.h"
#include "cg_host_io.hpp"

namespace {

[[noreturn]] void die(const std::string& m) { fprintf(stderr, "[circall_gpu] error: %s\n", m.c_str()); exit(1); }
void ck(int rc, const char* what) { if (rc) die(std::string(what) + ": " + cg_last_error()); }

struct Args {
    std::string index, bsj_index, out, out_bsj, fasta;
    std::vector<std::string> m1, m2;
    int k = 31, gpu = 0, gpus = 1, threads = 4, lce = 0;
    uint64_t batch = 1u << 20;
    bool force = false;
    bool wt_classes = false;      // also write Circall_wt's own rawCount.txt (its equivalence classes; nothing downstream reads it)
};

Args parse_args(int argc, char** argv, int first) {
    Args a;
    for (int i = first; i < argc; ++i) {
        std::string o = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) die("missing value after " + o); return argv[++i]; };
        auto vals = [&](std::vector<std::string>& dst) { while (i + 1 < argc && argv[i + 1][0] != '-') dst.push_back(argv[++i]); if (dst.empty()) die("missing value after " + o); };
        if (o == "-i" || o == "--index") a.index = val();
        else if (o == "-b" || o == "--bsjindex") a.bsj_index = val();
        else if (o == "-1" || o == "--mates1") vals(a.m1);
        else if (o == "-2" || o == "--mates2") vals(a.m2);
        else if (o == "-o" || o == "--output" || o == "--out") a.out = val();
        else if (o == "-O" || o == "--output-bsj") a.out_bsj = val();
        else if (o == "-p" || o == "--threads") a.threads = atoi(val().c_str());
        else if (o == "-t" || o == "--transcripts") a.fasta = val();
        else if (o == "-k" || o == "--kmerSize") a.k = atoi(val().c_str());
        else if (o == "-f" || o == "--force") a.force = true;
        else if (o == "--gpu") a.gpu = atoi(val().c_str());
        else if (o == "--gpus") a.gpus = atoi(val().c_str());
        else if (o == "--batch") a.batch = strtoull(val().c_str(), nullptr, 10);
        else if (o == "--lce-mode") a.lce = atoi(val().c_str());
        else if (o == "--wt-classes") a.wt_classes = true;
        else if (o.size() > 1 && o[0] == '-') { if (i + 1 < argc && argv[i + 1][0] != '-') ++i; }   // accepted and ignored
        else die("unexpected argument " + o);
    }
    return a;
}

void make_dir(const std::string& d) {
    if (d.empty()) die("no output directory given (-o)");
    std::string cur;
    for (size_t i = 0; i <= d.size(); ++i) {
        if (i == d.size() || d[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0777); }
        if (i < d.size()) cur.push_back(d[i]);
    }
    struct stat st;
    if (stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) die("cannot create directory " + d);
}
FILE* open_append(const std::string& path) {            // "a": outputs accumulate across runs, as in the reference (Circall_wt.cpp:231)
    FILE* f = fopen(path.c_str(), "a");
    if (!f) die("cannot open " + path);
    static std::vector<std::vector<char>> bufs;
    bufs.emplace_back(1 << 22);
    setvbuf(f, bufs.back().data(), _IOFBF, bufs.back().size());
    return f;
}

struct Outputs {
    // wt directory
    FILE* un1 = nullptr; FILE* un2 = nullptr;
    // bsj directory
    FILE* jt = nullptr; FILE* c1 = nullptr; FILE* c2 = nullptr;
    cgh::EqClasses eq;
};

const char* SUF1[3] = {"", "___11", "___21"};     // header suffix of the record written to UN_read1 (mate code 1 / 2)
const char* SUF2[3] = {"", "___22", "___12"};

// write what one fetched batch produced.  stage 1: wt files; stage 2: bsj files (rec = pair); stage 0: both
void emit(int stage, const cgh::Batch& B, const cg_batch_result& r, const cg_index* bsj, Outputs& o) {
    auto first_seq = [&](uint64_t rec, const char*& s, size_t& n, const char*& s2, size_t& n2, std::string& hA, std::string& hB) {
        // the (header, sequence) pair Circall_bsj sees for export record `rec`
        if (stage == 2) { s = B.seq1(rec); n = B.len1(rec); s2 = B.seq2(rec); n2 = B.len2(rec); hA = ">" + B.h1[rec]; hB = ">" + B.h2[rec]; return; }
        uint32_t i = r.exp_pair[rec]; int code = r.exp_code[rec];
        hA = ">" + B.h1[i] + SUF1[code]; hB = ">" + B.h1[i] + SUF2[code];      // both use mate 1's header (Circall_wt.cpp:526-527)
        if (code == 1) { s = B.seq1(i); n = B.len1(i); s2 = B.seq2(i); n2 = B.len2(i); }
        else { s = B.seq2(i); n = B.len2(i); s2 = B.seq1(i); n2 = B.len1(i); }
    };
    const char *s, *s2; size_t n, n2; std::string hA, hB;
    if (stage != 2 && o.un1) {
        for (uint64_t k = 0; k < r.n_export; ++k) {                              // Circall_wt.cpp:350-365,525-540
            first_seq(k, s, n, s2, n2, hA, hB);
            cgh::put_record(o.un1, hA, s, n);
            cgh::put_record(o.un2, hB, s2, n2);
        }
    }
    if (stage != 1) {
        for (uint64_t k = 0; k < r.n_junc; ++k) {                                // Circall_bsj.cpp:339-341,366-368
            const cg_junction& j = r.junc[k];
            first_seq(j.rec, s, n, s2, n2, hA, hB);
            fprintf(o.jt, "%s\t%u\t%u\t%u\t%d\t%s\t%u\n", hA.c_str(), j.dir, j.query_pos, j.len, j.hit_pos, cg_index_tx_name(bsj, j.tid), cg_index_tx_len(bsj, j.tid));
        }
        for (uint64_t k = 0; k < r.n_cand; ++k) {                                // Circall_bsj.cpp:392-399
            first_seq(r.cand_rec[k], s, n, s2, n2, hA, hB);
            cgh::put_record(o.c1, hA, s, n);
            cgh::put_record(o.c2, hB, s2, n2);
        }
        for (uint64_t c = 0; c < r.n_eq; ++c)                                    // multi-transcript classes (single ones are counted on the device)
            ++o.eq[std::vector<uint32_t>(r.eq_tids + r.eq_off[c], r.eq_tids + r.eq_off[c + 1])];
    }
}

int run_mapper(int stage, const Args& a) {
    if (a.m1.empty() || a.m2.empty()) die("both -1 and -2 are required (paired-end reads)");
    int nd = 0;
    cg_device_count(&nd);
    if (nd == 0) die("no CUDA device: libcircall_gpu has no CPU path");
    if (a.gpus < 1 || a.gpu < 0 || a.gpu + a.gpus > nd) die("--gpu / --gpus ask for devices this machine does not have (" + std::to_string(nd) + " visible)");
    const int ND = a.gpus, NU = 2 * ND;                  // devices, work units (two sessions per device)
    std::vector<cg_index*> txs(ND, nullptr), bsjs(ND, nullptr);
    for (int d = 0; d < ND; ++d) {                        // the indices are replicated in every GPU's memory
        if (stage != 2) ck(cg_index_open(a.index.c_str(), a.gpu + d, &txs[d]), "opening the transcript index");
        if (stage != 1) ck(cg_index_open((stage == 2 ? a.index : a.bsj_index).c_str(), a.gpu + d, &bsjs[d]), "opening the BSJ index");
    }
    cg_index *tx = txs[0], *bsj = bsjs[0];
    const std::string out_wt = a.out, out_bsj = stage == 0 ? a.out_bsj : a.out;
    Outputs o;
    const std::string tag = std::to_string(1804289383);          // std::rand() without srand (Circall_wt.cpp:228): any unique tag works
    if (stage != 2) { make_dir(out_wt); o.un1 = open_append(out_wt + "/UN_read1_" + tag + ".fa"); o.un2 = open_append(out_wt + "/UN_read2_" + tag + ".fa"); }
    if (stage != 1) {
        make_dir(out_bsj);
        o.jt = open_append(out_bsj + "/UN_read1_" + tag + ".fa");
        o.c1 = open_append(out_bsj + "/BSJ_read1_" + tag + ".fa"); o.c2 = open_append(out_bsj + "/BSJ_read2_" + tag + ".fa");
    }
    cg_opts opts; memset(&opts, 0, sizeof opts);
    opts.lce_mode = a.lce; opts.stage = stage; opts.max_batch_pairs = (uint32_t)a.batch;
    std::vector<cg_session*> S(NU, nullptr);
    std::vector<cgh::Batch> B(NU);
    const uint64_t cap = std::max<uint64_t>(a.batch * 160, 1u << 20);
    for (int x = 0; x < NU; ++x) {                       // unit x lives on device x % ND: consecutive batches go to different GPUs
        ck(cg_session_create(txs[x % ND], bsjs[x % ND], &opts, &S[x]), "creating a session");
        void *p1, *p2;
        ck(cg_host_alloc(cap, &p1), "allocating pinned memory"); ck(cg_host_alloc(cap, &p2), "allocating pinned memory");
        B[x].b1 = (char*)p1; B[x].b2 = (char*)p2; B[x].cap = cap;
    }
    cgh::PairStream in(a.m1, a.m2);
    uint64_t pairs = 0;
    int cur = 0;
    std::vector<char> pending(NU, 0);
    // Circall_wt's own equivalence classes (src/Circall_wt.cpp:369-465,544-640; opt-in, --wt-classes): the transcripts of mate 1's hits
    // form a group, and mate 2's are APPENDED to them for a second group — txpIDsCompat is not cleared between the mates (:247-252).
    // Nothing downstream reads the file, and the mapping kernels do not keep hit lists, so the lists come from the collector surface
    // (cg_collect) in a second pass over the batch: a switch for completeness, not the fast path.
    cgh::EqClasses wt_eq;
    auto wt_classes_of = [&](int x) {
        const cgh::Batch& b = B[x];
        std::vector<std::vector<uint32_t>> left((size_t)b.n);
        cg_collect_result cr;
        ck(cg_collect(txs[x % ND], b.b1, b.off1.data(), b.n, 1, a.lce, &cr), "collecting mate 1's hits");
        for (uint64_t i = 0; i < b.n; ++i) {
            const uint64_t h0 = cr.hits_off[i], h1 = cr.hits_off[i + 1];
            if (h1 > h0 && h1 - h0 <= 200) left[(size_t)i].assign(cr.hit_tid + h0, cr.hit_tid + h1);      // maxReadOccs (:372)
        }
        ck(cg_collect(txs[x % ND], b.b2, b.off2.data(), b.n, 1, a.lce, &cr), "collecting mate 2's hits");
        for (uint64_t i = 0; i < b.n; ++i) {
            std::vector<uint32_t>& g = left[(size_t)i];
            if (!g.empty()) ++wt_eq[g];                                                                   // :450-451
            const uint64_t h0 = cr.hits_off[i], h1 = cr.hits_off[i + 1];
            if (h1 > h0 && h1 - h0 <= 200) { g.insert(g.end(), cr.hit_tid + h0, cr.hit_tid + h1); ++wt_eq[g]; }   // :625-626
        }
    };
    auto finish = [&](int x) {
        cg_batch_result r;
        ck(cg_session_fetch(S[x], 1, &r), "mapping a batch");
        emit(stage, B[x], r, bsj, o);
        if (a.wt_classes && stage != 2) wt_classes_of(x);
        pending[x] = 0;
    };
    try {
        for (;;) {
            if (pending[cur]) finish(cur);
            if (!in.fill(B[cur], a.batch)) break;
            // reads of one length (the usual case): no offsets across PCIe
            uint64_t len = B[cur].n ? B[cur].off1[1] : 0;
            bool fixed = len > 0;
            for (uint64_t i = 0; fixed && i < B[cur].n; ++i)
                fixed = B[cur].off1[i + 1] - B[cur].off1[i] == len && B[cur].off2[i + 1] - B[cur].off2[i] == len;
            if (fixed) ck(cg_session_submit_fixed(S[cur], B[cur].b1, B[cur].b2, B[cur].n, (uint32_t)len), "submitting a batch");
            else ck(cg_session_submit(S[cur], B[cur].b1, B[cur].off1.data(), B[cur].b2, B[cur].off2.data(), B[cur].n), "submitting a batch");
            pending[cur] = 1;
            pairs += B[cur].n;
            cur = (cur + 1) % NU;
        }
    } catch (const std::string& e) { die(e); }
    for (int x = 0; x < NU; ++x) if (pending[(x + cur) % NU]) finish((x + cur) % NU);      // oldest first: the files keep the input order

    // totals -> fragmentInfo.txt / rawCount.txt (ExportFeq.cpp:46-105): one all-reduce over every session of every device, after
    // which any session holds them
    cg_index_info bi; memset(&bi, 0, sizeof bi);
    if (bsj) cg_index_get_info(bsj, &bi);
    std::vector<uint64_t> per(bi.n_tx, 0);
    cg_counters tot; memset(&tot, 0, sizeof tot);
    ck(cg_sessions_allreduce(S.data(), NU), "reducing the counters");
    ck(cg_session_counters(S[0], &tot, bsj ? per.data() : nullptr, bsj ? bi.n_tx : 0), "reading the counters");
    cgh::FragPrior fp;
    cg_index_info ti; memset(&ti, 0, sizeof ti);
    if (tx) cg_index_get_info(tx, &ti);
    if (stage != 2 && !cgh::write_fragment_info(out_wt + "/fragmentInfo.txt", tot.wt_read_len, fp, tot.wt_num_observed, tot.wt_num_mapped, tot.wt_num_hits, ti.k))
        die("cannot write fragmentInfo.txt");
    if (stage != 2 && a.wt_classes) {
        FILE* f = fopen((out_wt + "/rawCount.txt").c_str(), "w");
        if (!f) die("cannot write rawCount.txt");
        fprintf(f, "Transcript\tWeight\tCount\teffLens\tRefLength\tEffectiveLength\teqClass\n");       // ExportFeq.cpp:88
        uint32_t cls = 0;
        for (auto& kv : wt_eq) {
            ++cls;
            for (uint32_t t : kv.first) {                                                    // weights 1/classSize (EquivalenceClassBuilder.hpp:32-40)
                const uint32_t len = cg_index_tx_len(tx, t);
                const double e = fp.effective_length(len);
                fprintf(f, "%s\t%g\t%llu\t%g\t%u\t%g\t%u\n", cg_index_tx_name(tx, t), 1.0 / (double)kv.first.size(), (unsigned long long)kv.second, e, len, e, cls);
            }
        }
        fclose(f);
    }
    if (stage != 1) {
        if (!cgh::write_fragment_info(out_bsj + "/fragmentInfo.txt", tot.bsj_read_len, fp, tot.bsj_num_observed, tot.bsj_num_mapped, tot.bsj_num_hits, bi.k))
            die("cannot write fragmentInfo.txt");
        FILE* f = fopen((out_bsj + "/rawCount.txt").c_str(), "w");
        if (!f) die("cannot write rawCount.txt");
        fprintf(f, "Transcript\tWeight\tCount\teffLens\tRefLength\tEffectiveLength\teqClass\n");       // ExportFeq.cpp:88
        uint32_t cls = 0;
        auto row = [&](uint32_t t, double w, uint64_t n) {
            uint32_t len = cg_index_tx_len(bsj, t);
            double e = fp.effective_length(len);
            fprintf(f, "%s\t%g\t%llu\t%g\t%u\t%g\t%u\n", cg_index_tx_name(bsj, t), w, (unsigned long long)n, e, len, e, cls);
        };
        for (uint32_t t = 0; t < bi.n_tx; ++t) if (per[t]) { ++cls; row(t, 1.0, per[t]); }
        for (auto& kv : o.eq) { ++cls; for (uint32_t t : kv.first) row(t, 1.0 / (double)kv.first.size(), kv.second); }   // weights 1/classSize (EquivalenceClassBuilder.hpp:32-40)
        fclose(f);
    }
    for (FILE* f : {o.un1, o.un2, o.jt, o.c1, o.c2}) if (f) fclose(f);
    fprintf(stderr, "[circall_gpu] %llu pairs", (unsigned long long)pairs);
    if (stage != 2) fprintf(stderr, "; wt: observed %llu mapped %llu hits %llu", (unsigned long long)tot.wt_num_observed, (unsigned long long)tot.wt_num_mapped, (unsigned long long)tot.wt_num_hits);
    if (stage != 1) fprintf(stderr, "; bsj: observed %llu mapped %llu hits %llu", (unsigned long long)tot.bsj_num_observed, (unsigned long long)tot.bsj_num_mapped, (unsigned long long)tot.bsj_num_hits);
    fprintf(stderr, "\n");
    for (int x = 0; x < NU; ++x) { cg_session_destroy(S[x]); cg_host_free(B[x].b1); cg_host_free(B[x].b2); }
    for (int d = 0; d < ND; ++d) { if (txs[d]) cg_index_close(txs[d]); if (bsjs[d]) cg_index_close(bsjs[d]); }
    return 0;
}

// Circall_pseudo (src/Circall_pseudo.cpp:216-555, defaults: allowOrphans, ignoreLibCompat, fuzzy intersection): both
// mates through the collector, mergeLeftRightHitsFuzzy (:271-280, maxNumHits = maxReadOccs = 200), more than 200 joint
// hits -> none (:296); a pair with joint hits writes its header to Mapped_read_<tag>.rh and the names of the hit
// transcripts, in transcript order, each followed by a space, to mapInfo_<tag>.txt (:474-508).
int run_pseudo(const Args& a) {
    if (a.m1.empty() || a.m2.empty()) die("both -1 and -2 are required (paired-end reads)");
    int nd = 0;
    cg_device_count(&nd);
    if (nd == 0) die("no CUDA device: libcircall_gpu has no CPU path");
    cg_index* idx = nullptr;
    ck(cg_index_open(a.index.c_str(), a.gpu, &idx), "opening the pseudo-sequence index");
    make_dir(a.out);
    const std::string tag = std::to_string(1804289383);
    FILE* rh = open_append(a.out + "/Mapped_read_" + tag + ".rh");
    FILE* mi = open_append(a.out + "/mapInfo_" + tag + ".txt");
    cgh::Batch B;
    const uint64_t maxp = std::min<uint64_t>(a.batch, 1u << 18);
    std::vector<char> b1(std::max<uint64_t>(maxp * 320, 1u << 20)), b2(b1.size());
    B.b1 = b1.data(); B.b2 = b2.data(); B.cap = b1.size();
    cgh::PairStream in(a.m1, a.m2);
    uint64_t pairs = 0, mapped = 0, hits = 0;
    std::vector<uint64_t> loff, joff;
    std::vector<uint32_t> ltid; std::vector<int32_t> lpos, joint; std::vector<uint8_t> lfwd, lfound, tm;
    try {
        while (in.fill(B, maxp)) {
            const uint64_t n = B.n;
            cg_collect_result L, R;
            ck(cg_collect(idx, B.b1, B.off1.data(), n, 1, a.lce, &L), "mapping mate 1");
            loff.assign(L.hits_off, L.hits_off + n + 1);                       // the buffers are reused by the next call
            ltid.assign(L.hit_tid, L.hit_tid + loff[n]); lpos.assign(L.hit_pos, L.hit_pos + loff[n]); lfwd.assign(L.hit_fwd, L.hit_fwd + loff[n]);
            lfound.assign(L.found, L.found + n);
            ck(cg_collect(idx, B.b2, B.off2.data(), n, 1, a.lce, &R), "mapping mate 2");
            const uint64_t cap = loff[n] + R.hits_off[n] + n + 1;
            joff.assign(n + 1, 0); joint.assign(cap * 7, 0); tm.assign(n, 0);
            ck(cg_merge_pairs(a.gpu, n, loff.data(), ltid.data(), lpos.data(), lfwd.data(), lfound.data(), R.hits_off, R.hit_tid, R.hit_pos, R.hit_fwd, R.found,
                              (uint32_t)B.len1(0), 200, joff.data(), joint.data(), cap, tm.data()), "intersecting the mates' hits");
            for (uint64_t i = 0; i < n; ++i) {
                const uint64_t nj = joff[i + 1] - joff[i];
                if (nj == 0 || nj > 200) continue;                                // :296
                std::vector<uint32_t> t;
                for (uint64_t h = joff[i]; h < joff[i + 1]; ++h) t.push_back((uint32_t)joint[7 * h]);
                if (joint[7 * joff[i] + 6] != 3) std::stable_sort(t.begin(), t.end());   // orphans: left then right lists merged by transcript (:304-319)
                fprintf(rh, ">%s\n", B.h1[i].c_str());
                for (uint32_t x : t) { fputs(cg_index_tx_name(idx, x), mi); fputc(' ', mi); }
                fputc('\n', mi);
                ++mapped; hits += nj;
            }
            pairs += n;
        }
    } catch (const std::string& e) { die(e); }
    cg_collect_free();
    fclose(rh); fclose(mi);
    fprintf(stderr, "[circall_gpu] %llu pairs; pseudo: mapped %llu hits %llu\n", (unsigned long long)pairs, (unsigned long long)mapped, (unsigned long long)hits);
    cg_index_close(idx);
    return 0;
}

int run_indexer(const Args& a) {
    if (a.fasta.empty() || a.out.empty()) die("usage: TxIndexer -t transcripts.fa -o indexDir [-k 31] [-f]");
    struct stat st;
    if (!a.force && stat((a.out + "/header.json").c_str(), &st) == 0) {          // TxIndexer.cpp:192-200
        fprintf(stderr, "[circall_gpu] index %s exists (use -f to rebuild)\n", a.out.c_str());
        return 0;
    }
    make_dir(a.out);
    cg_index* idx = nullptr;
    ck(cg_index_build_fasta(a.gpu, a.fasta.c_str(), a.k, &idx), "building the index");
    ck(cg_index_save(idx, a.out.c_str()), "writing the index");
    cg_index_info i;
    cg_index_get_info(idx, &i);
    fprintf(stderr, "[circall_gpu] indexed %u sequences, %llu characters, %llu distinct %d-mers in %.2f s\n", i.n_tx, (unsigned long long)i.text_len,
            (unsigned long long)i.n_kmers, i.k, i.build_seconds);
    cg_index_close(idx);
    return 0;
}

// circall_gpu fld LEN...: the fragment-length prior's columns of fragmentInfo.txt and the effective length of references of the given
// lengths (no GPU; tests/test_cli.py pins them against a restatement of EmpiricalDistribution.cpp)
int run_fld(int argc, char** argv, int first) {
    cgh::FragPrior fp;
    printf("median %.9g mean %.9g sd %.9g\n", (double)fp.median, (double)fp.mean, (double)fp.sd);
    for (int i = first; i < argc; ++i) printf("efflen %s %.17g\n", argv[i], fp.effective_length((uint32_t)strtoul(argv[i], nullptr, 10)));
    return 0;
}

int run_parse(const Args& a) {
    cgh::Batch B;
    std::vector<char> b1(1 << 24), b2(1 << 24);
    B.b1 = b1.data(); B.b2 = b2.data(); B.cap = b1.size();
    cgh::PairStream in(a.m1, a.m2);
    uint64_t pairs = 0, bases = 0, hsum = 0;
    try {
        while (in.fill(B, a.batch)) {
            pairs += B.n; bases += B.off1.back() + B.off2.back();
            for (uint64_t i = 0; i < B.n; ++i) { for (char c : B.h1[i]) hsum = hsum * 131 + (unsigned char)c; for (char c : B.h2[i]) hsum = hsum * 131 + (unsigned char)c; }
        }
    } catch (const std::string& e) { die(e); }
    printf("pairs %llu bases %llu headers %llu\n", (unsigned long long)pairs, (unsigned long long)bases, (unsigned long long)hsum);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    std::string me = argv[0];
    size_t sl = me.find_last_of('/');
    if (sl != std::string::npos) me = me.substr(sl + 1);
    int first = 1;
    std::string cmd;
    if (me == "Circall_wt") cmd = "wt";
    else if (me == "Circall_bsj") cmd = "bsj";
    else if (me == "Circall_pseudo") cmd = "pseudo";
    else if (me == "TxIndexer") cmd = "index";
    else { if (argc < 2) die("usage: circall_gpu {wt|bsj|fused|pseudo|index|parse} options..."); cmd = argv[1]; first = 2; }
    if (cmd == "fld") return run_fld(argc, argv, first);
    Args a = parse_args(argc, argv, first);
    if (cmd == "wt") return run_mapper(1, a);
    if (cmd == "bsj") return run_mapper(2, a);
    if (cmd == "fused") { if (a.bsj_index.empty() || a.out_bsj.empty()) die("fused needs -b BSJINDEX and -O OUTBSJ"); return run_mapper(0, a); }
    if (cmd == "pseudo") return run_pseudo(a);
    if (cmd == "index") return run_indexer(a);
    if (cmd == "parse") return run_parse(a);
    die("unknown command " + cmd);
}
