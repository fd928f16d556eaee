// Synthetic data for tests and bench.py (NOT part of the product path, NOT part of the oracle).
//
// Generates, deterministically from seeds:
//   * a "genome" of genes (exons/introns), a transcriptome built from exon subsets so 31-mers
//     multi-map across isoforms, with optional antisense genes (k-mers present on both strands ->
//     exercises the collector's strand vote), repeat-family inserts (wide SA intervals -> exercises
//     maxInterval) and internal poly-A runs (homopolymer k-mer skip);
//   * the BSJ reference with the geometry of R/buildBSJdb.R:230-234 — for every gene, every
//     exon-start x exon-end pair with start <= end: last (maxReadLen-1) genomic bases ending at the
//     exon end ++ first (maxReadLen-1) genomic bases from the exon start, named
//     chr__start__end__gene__exS__exE;
//   * paired-end reads: linear fragments from transcripts and back-spliced fragments from circular
//     exon ranges, substitution errors, optional N, unstranded (recipe of R/Circall_simulator.R).
//
// Every random draw is a pure function of (seed, object index), so output does not depend on
// the number of threads.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <algorithm>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    static uint64_t mix(uint64_t z) {
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    uint64_t next() { s += 0x9e3779b97f4a7c15ULL; return mix(s); }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double gauss() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};
static inline uint64_t key2(uint64_t seed, uint64_t a, uint64_t b) {
    return Rng::mix(seed * 0x9e3779b97f4a7c15ULL + Rng::mix(a + 0x632be59bd9b4e019ULL) + 3 * Rng::mix(b ^ 0xd1342543de82ef95ULL));
}
static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

struct Params {
    uint64_t seed;
    uint32_t n_genes;
    double mean_iso;        // mean isoforms per gene (1..8)
    double p_antisense;     // probability a gene is the reverse complement of the previous gene
    double p_repeat;        // probability an exon (>= 320 bp) receives a repeat-family insert
    double p_polya;         // probability an exon (>= 200 bp) receives an internal 40-A run
    uint32_t bsj_flank;     // maxReadLen - 1 (149 for the documented maxReadLen = 150)
};

struct Gene {
    uint64_t goff;                    // offset of this gene in the genome string
    uint32_t glen;
    std::vector<uint32_t> es, ee;     // exon start / end (inclusive), gene-local coordinates
};

struct Txome {
    Params p;
    std::string genome;
    std::vector<Gene> genes;
    std::string tx_text;              // transcripts joined by '$'
    std::vector<uint64_t> tx_off;     // n_tx + 1
    std::vector<uint32_t> tx_gene;
    std::string bsj_text;
    std::vector<uint32_t> bsj_gene, bsj_exs, bsj_exe;   // per record
    std::vector<std::string> repeats;
};

static void gen_gene(Txome& T, uint32_t g) {
    const Params& p = T.p;
    Rng r(key2(p.seed, 1, g));
    Gene G;
    G.goff = T.genome.size();
    const uint32_t flank = p.bsj_flank + 1;
    bool anti = g > 0 && r.uni() < p.p_antisense;
    if (anti) {
        const Gene& P = T.genes[g - 1];
        std::string s(P.glen, 'A');
        for (uint32_t i = 0; i < P.glen; ++i) s[i] = comp(T.genome[P.goff + P.glen - 1 - i]);
        G.glen = P.glen;
        for (size_t e = P.es.size(); e-- > 0;) { G.es.push_back(P.glen - 1 - P.ee[e]); G.ee.push_back(P.glen - 1 - P.es[e]); }
        T.genome += s;
        T.genes.push_back(G);
        return;
    }
    uint32_t E = 2 + (uint32_t)r.below(13);            // 2..14 exons
    std::string s;
    auto rnd = [&](uint32_t n) { for (uint32_t i = 0; i < n; ++i) s.push_back(BASES[r.next() & 3]); };
    rnd(flank);
    for (uint32_t e = 0; e < E; ++e) {
        double ln = std::exp(std::log(170.0) + 0.65 * r.gauss());
        if (e == 0 || e + 1 == E) ln *= 2.0;
        uint32_t len = (uint32_t)std::min(3000.0, std::max(45.0, ln));
        uint32_t st = (uint32_t)s.size();
        rnd(len);
        if (len >= 320 && !T.repeats.empty() && r.uni() < p.p_repeat) {
            size_t fam = r.uni() < 0.8 ? 0 : 1 + r.below(T.repeats.size() - 1 ? T.repeats.size() - 1 : 1);
            if (fam >= T.repeats.size()) fam = 0;
            const std::string& rep = T.repeats[fam];
            uint32_t at = st + 10 + (uint32_t)r.below(len - (uint32_t)rep.size() - 20);
            double div = r.uni() < 0.5 ? 0.0 : 0.02 + 0.06 * r.uni();
            for (size_t i = 0; i < rep.size(); ++i) {
                char c = rep[i];
                if (div > 0 && r.uni() < div) c = BASES[(std::strchr("ACGT", c) - "ACGT" + 1 + r.below(3)) & 3];
                s[at + i] = c;
            }
        } else if (len >= 200 && r.uni() < p.p_polya) {
            uint32_t at = st + 60 + (uint32_t)r.below(len - 130);
            for (uint32_t i = 0; i < 40; ++i) s[at + i] = 'A';
        }
        G.es.push_back(st); G.ee.push_back(st + len - 1);
        if (e + 1 < E) rnd(150 + (uint32_t)r.below(250));   // intron
    }
    rnd(flank);
    G.glen = (uint32_t)s.size();
    T.genome += s;
    T.genes.push_back(G);
}

static void gen_transcripts(Txome& T, uint32_t g) {
    const Params& p = T.p;
    const Gene& G = T.genes[g];
    Rng r(key2(p.seed, 2, g));
    uint32_t E = (uint32_t)G.es.size();
    // isoform count: 1 + geometric with the requested mean, capped at 8
    double q = 1.0 / std::max(1.0, p.mean_iso);
    uint32_t niso = 1;
    while (niso < 8 && r.uni() > q) ++niso;
    for (uint32_t it = 0; it < niso; ++it) {
        std::vector<uint32_t> ex;
        uint32_t a = (E > 2 && r.uni() < 0.2) ? 1 : 0;
        uint32_t b = (E > 3 && r.uni() < 0.2) ? E - 2 : E - 1;
        for (uint32_t e = a; e <= b; ++e)
            if (e == a || e == b || it == 0 || r.uni() < 0.8) ex.push_back(e);
        uint64_t len = 0;
        for (auto e : ex) len += G.ee[e] - G.es[e] + 1;
        if (len < 300) { ex.clear(); for (uint32_t e = 0; e < E; ++e) ex.push_back(e); }
        uint64_t st = T.tx_text.size();
        for (auto e : ex) T.tx_text.append(T.genome, G.goff + G.es[e], G.ee[e] - G.es[e] + 1);
        // keep the uncertain index-time normalisation out of play (no >=10-A tail)
        size_t n = T.tx_text.size();
        if (n - st >= 10) {
            bool tail = true;
            for (size_t i = n - 10; i < n; ++i) if (T.tx_text[i] != 'A') { tail = false; break; }
            if (tail) T.tx_text[n - 1] = 'C';
        }
        // pad very short genes so every transcript is >= 300
        while (T.tx_text.size() - st < 300) T.tx_text.push_back(BASES[r.next() & 3]);
        T.tx_off.push_back(st);
        T.tx_gene.push_back(g);
        T.tx_text.push_back('$');
    }
}

static void gen_bsj(Txome& T, uint32_t g) {
    const Gene& G = T.genes[g];
    const uint32_t F = T.p.bsj_flank;
    uint32_t E = (uint32_t)G.es.size();
    for (uint32_t i = 0; i < E; ++i)          // donor: end of exon i
        for (uint32_t j = 0; j <= i; ++j) {   // acceptor: start of exon j (j <= i: back-splice)
            T.bsj_text.append(T.genome, G.goff + G.ee[i] + 1 - F, F);
            T.bsj_text.append(T.genome, G.goff + G.es[j], F);
            T.bsj_text.push_back('$');
            T.bsj_gene.push_back(g); T.bsj_exs.push_back(j); T.bsj_exe.push_back(i);
        }
}

struct ReadParams {
    uint64_t seed;
    uint32_t L;
    double frag_mean, frag_sd;
    double err, circ_frac, n_rate, junk_frac;
};

static void gen_pair(const Txome& T, const ReadParams& rp, uint64_t idx, char* m1, char* m2, int32_t* truth) {
    Rng r(key2(rp.seed, 3, idx));
    const uint32_t L = rp.L;
    std::string frag;
    int32_t kind = 0, tid = -1, tpos = -1;
    double u = r.uni();
    if (u < rp.junk_frac) {
        kind = 2;
        frag.resize(2 * L);
        for (auto& c : frag) c = BASES[r.next() & 3];
    } else {
        int64_t F = (int64_t)std::llround(rp.frag_mean + rp.frag_sd * r.gauss());
        if (F < (int64_t)L) F = L;
        if (F > 1000) F = 1000;
        bool circ = u < rp.junk_frac + rp.circ_frac;
        if (circ) {
            // circular RNA from exons j..i of a random gene; fragments wrap around the circle
            for (int tries = 0; tries < 64; ++tries) {
                uint32_t g = (uint32_t)r.below(T.genes.size());
                const Gene& G = T.genes[g];
                uint32_t E = (uint32_t)G.es.size();
                uint32_t j = (uint32_t)r.below(E), i = j + (uint32_t)r.below(std::min<uint32_t>(E - j, 4));
                std::string C;
                for (uint32_t e = j; e <= i; ++e) C.append(T.genome, G.goff + G.es[e], G.ee[e] - G.es[e] + 1);
                if (C.size() < 150) continue;
                uint64_t st = r.below(C.size());
                frag.resize((size_t)F);
                for (int64_t x = 0; x < F; ++x) frag[(size_t)x] = C[(st + (uint64_t)x) % C.size()];
                kind = 1; tid = (int32_t)g; tpos = (int32_t)(j * 1000 + i);
                break;
            }
        }
        if (frag.empty()) {
            uint32_t ntx = (uint32_t)T.tx_gene.size();
            uint32_t t = (uint32_t)r.below(ntx);
            uint64_t st = T.tx_off[t], len = T.tx_off[t + 1] - 1 - st;
            if ((uint64_t)F > len) F = (int64_t)len;
            uint64_t p = r.below(len - (uint64_t)F + 1);
            frag.assign(T.tx_text, st + p, (size_t)F);
            kind = 0; tid = (int32_t)t; tpos = (int32_t)p;
        }
    }
    size_t F = frag.size();
    if (r.next() & 1) {   // unstranded library: fragment read from the other strand
        std::string f2(F, 'A');
        for (size_t i = 0; i < F; ++i) f2[i] = comp(frag[F - 1 - i]);
        frag.swap(f2);
        kind |= 16;
    }
    for (uint32_t i = 0; i < L; ++i) { m1[i] = frag[i]; m2[i] = comp(frag[F - 1 - i]); }
    for (int m = 0; m < 2; ++m) {
        char* s = m ? m2 : m1;
        for (uint32_t i = 0; i < L; ++i) {
            double e = r.uni();
            if (e < rp.err) s[i] = BASES[(std::strchr("ACGT", s[i]) - "ACGT" + 1 + r.below(3)) & 3];
            else if (e < rp.err + rp.n_rate) s[i] = 'N';
        }
    }
    if (truth) { truth[0] = kind; truth[1] = tid; truth[2] = tpos; }
}

} // namespace

extern "C" {

void* syn_txome_create(uint64_t seed, uint32_t n_genes, double mean_iso, double p_antisense, double p_repeat,
                       double p_polya, uint32_t bsj_flank) {
    auto* T = new Txome();
    T->p = Params{seed, n_genes, mean_iso, p_antisense, p_repeat, p_polya, bsj_flank};
    Rng rr(key2(seed, 0, 0));
    for (int f = 0; f < 4; ++f) {
        std::string rep(280, 'A');
        for (auto& c : rep) c = BASES[rr.next() & 3];
        T->repeats.push_back(rep);
    }
    for (uint32_t g = 0; g < n_genes; ++g) gen_gene(*T, g);
    for (uint32_t g = 0; g < n_genes; ++g) gen_transcripts(*T, g);
    T->tx_off.push_back(T->tx_text.size());
    for (uint32_t g = 0; g < n_genes; ++g) gen_bsj(*T, g);
    return T;
}
void syn_txome_free(void* h) { delete (Txome*)h; }
const char* syn_tx_text(void* h, uint64_t* n, uint32_t* ntx) { auto* T = (Txome*)h; *n = T->tx_text.size(); *ntx = (uint32_t)T->tx_gene.size(); return T->tx_text.data(); }
const char* syn_bsj_text(void* h, uint64_t* n, uint32_t* nrec) { auto* T = (Txome*)h; *n = T->bsj_text.size(); *nrec = (uint32_t)T->bsj_gene.size(); return T->bsj_text.data(); }
uint32_t syn_num_genes(void* h) { return (uint32_t)((Txome*)h)->genes.size(); }
int syn_tx_name(void* h, uint32_t t, char* buf, int cap) {
    auto* T = (Txome*)h;
    return snprintf(buf, cap, "SYNT%08u.G%06u", t, T->tx_gene[t]);
}
// chr__start__end__gene__exS__exE  (R/buildBSJdb.R:234; >= 5 "__" separators, Circall_functions.R:32-36)
int syn_bsj_name(void* h, uint32_t i, char* buf, int cap) {
    auto* T = (Txome*)h;
    uint32_t g = T->bsj_gene[i];
    const Gene& G = T->genes[g];
    unsigned long long s = G.goff + G.es[T->bsj_exs[i]] + 1, e = G.goff + G.ee[T->bsj_exe[i]] + 1;
    return snprintf(buf, cap, "chrS__%llu__%llu__G%06u__%u__%u", s, e, g, g * 100 + T->bsj_exs[i], g * 100 + T->bsj_exe[i]);
}

// exons of gene g in genome coordinates (1-based, inclusive, like the coordinates in the BSJ names): returns their number
int syn_gene_exons(void* h, uint32_t g, uint64_t* starts, uint64_t* ends, int cap) {
    auto* T = (Txome*)h;
    if (g >= T->genes.size()) return 0;
    const Gene& G = T->genes[g];
    int n = (int)G.es.size();
    for (int i = 0; i < n && i < cap; ++i) { starts[i] = G.goff + G.es[i] + 1; ends[i] = G.goff + G.ee[i] + 1; }
    return n;
}

// fixed-length reads: r1/r2 are n_pairs*L byte arrays; truth (optional) = 3 int32 per pair:
// kind (0 linear, 1 back-spliced, 2 junk; +16 if the fragment was read from the other strand), id, pos
void syn_reads(void* h, uint64_t seed, uint64_t first, uint64_t n_pairs, uint32_t L, double frag_mean, double frag_sd,
               double err, double circ_frac, double n_rate, double junk_frac, char* r1, char* r2, int32_t* truth,
               int nthreads) {
    auto* T = (Txome*)h;
    ReadParams rp{seed, L, frag_mean, frag_sd, err, circ_frac, n_rate, junk_frac};
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            uint64_t b = n_pairs * (uint64_t)t / (uint64_t)nthreads, e = n_pairs * (uint64_t)(t + 1) / (uint64_t)nthreads;
            for (uint64_t i = b; i < e; ++i)
                gen_pair(*T, rp, first + i, r1 + i * L, r2 + i * L, truth ? truth + 3 * i : nullptr);
        });
    for (auto& t : th) t.join();
}

} // extern "C"
