// circall_b200 — host side of the drop-in executables: read-pair streaming and the reference's output files.
//
// What it mirrors (paths relative to Circall_v1.0.1/ in datngu/Circall):
//   PairStream        pair_sequence_parser<char**> as Circall_wt.cpp:87,1197-1200 uses it: FASTA or FASTQ,
//                     detected per file from its first byte; header = the whole line after '>' / '@' (spaces kept,
//                     the R stage matches on it); any number of files per mate, read strictly sequentially so that
//                     FIFOs (`<(gunzip -c x.fa.gz)`, Circall_source.sh:273) work.  Files go through zlib, which passes plain
//                     text through and inflates gzip: `-1 reads_1.fq.gz` needs no gunzip child any more (row N3, second half;
//                     the inflate stays on the host).
//   write_* helpers   the record formats of SURVEY Appendix C (Circall_wt.cpp:352-363,526-538; Circall_bsj.cpp:339-341,
//                     366-368,392-399; ExportFeq.cpp:81-103).
// Header-only, no CUDA: the CPU tests drive it through `circall_gpu parse`.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <zlib.h>

namespace cgh {

// one mate file list, read record by record
class SeqStream {
public:
    explicit SeqStream(std::vector<std::string> files) : files_(std::move(files)) {}
    ~SeqStream() { if (f_) gzclose(f_); }
    // next record; false at the end of the last file.  Throws std::string on a malformed file.
    bool next(std::string& header, std::string& seq) {
        for (;;) {
            if (!f_) {
                if (fi_ >= files_.size()) return false;
                f_ = gzopen(files_[fi_].c_str(), "rb");
                if (!f_) throw std::string("cannot open ") + files_[fi_];
                gzbuffer(f_, 1 << 20);
                ++fi_;
                kind_ = 0; have_line_ = false; bp_ = bn_ = 0;
            }
            if (read_record(header, seq)) return true;
            gzclose(f_); f_ = nullptr;
        }
    }

private:
    int getc_buf() {
        if (bp_ == bn_) {
            const int n = gzread(f_, buf_, (unsigned)sizeof buf_);
            if (n < 0) { int e = 0; const char* m = gzerror(f_, &e); throw std::string("read error in ") + files_[fi_ - 1] + ": " + (m ? m : "?"); }
            if (n == 0) return EOF;
            bp_ = 0; bn_ = (size_t)n;
        }
        return (unsigned char)buf_[bp_++];
    }
    bool getline(std::string& s) {
        if (have_line_) { s.swap(line_); have_line_ = false; return true; }
        s.clear();
        int c;
        bool any = false;
        while ((c = getc_buf()) != EOF) {
            any = true;
            if (c == '\n') break;
            s.push_back((char)c);
        }
        if (!any) return false;
        if (!s.empty() && s.back() == '\r') s.pop_back();
        return true;
    }
    void unget(std::string& s) { line_.swap(s); have_line_ = true; }
    bool read_record(std::string& header, std::string& seq) {
        std::string l;
        do { if (!getline(l)) return false; } while (l.empty());
        if (kind_ == 0) {
            if (l[0] == '>') kind_ = 1; else if (l[0] == '@') kind_ = 2;
            else throw std::string("neither FASTA nor FASTQ: ") + files_[fi_ - 1];
        }
        if (l[0] != (kind_ == 1 ? '>' : '@')) throw std::string("malformed record in ") + files_[fi_ - 1];
        header.assign(l, 1, std::string::npos);
        seq.clear();
        if (kind_ == 1) {
            while (getline(l)) {
                if (!l.empty() && l[0] == '>') { unget(l); break; }
                seq += l;
            }
        } else {
            if (!getline(seq)) throw std::string("truncated FASTQ record in ") + files_[fi_ - 1];
            std::string plus, qual;
            if (!getline(plus) || plus.empty() || plus[0] != '+' || !getline(qual)) throw std::string("truncated FASTQ record in ") + files_[fi_ - 1];
        }
        return true;
    }
    std::vector<std::string> files_;
    size_t fi_ = 0;
    gzFile f_ = nullptr;
    char buf_[1 << 16];
    size_t bp_ = 0, bn_ = 0;
    int kind_ = 0;
    std::string line_;
    bool have_line_ = false;
};

// one batch of read pairs in the layout cg_session_submit takes (bases back to back + n + 1 offsets per mate)
struct Batch {
    char* b1 = nullptr; char* b2 = nullptr;            // caller-provided (pinned) buffers of `cap` bytes each
    uint64_t cap = 0;
    std::vector<uint64_t> off1, off2;
    std::vector<std::string> h1, h2;
    uint64_t n = 0;
    void clear() { off1.assign(1, 0); off2.assign(1, 0); h1.clear(); h2.clear(); n = 0; }
    const char* seq1(uint64_t i) const { return b1 + off1[i]; }
    const char* seq2(uint64_t i) const { return b2 + off2[i]; }
    size_t len1(uint64_t i) const { return (size_t)(off1[i + 1] - off1[i]); }
    size_t len2(uint64_t i) const { return (size_t)(off2[i + 1] - off2[i]); }
};

class PairStream {
public:
    PairStream(std::vector<std::string> f1, std::vector<std::string> f2) : s1_(std::move(f1)), s2_(std::move(f2)) {}
    // fill b with up to max_pairs pairs (fewer when a buffer is full); false when nothing is left
    bool fill(Batch& b, uint64_t max_pairs) {
        b.clear();
        while (b.n < max_pairs) {
            if (!pending_) {
                bool a = s1_.next(ph1_, ps1_), c = s2_.next(ph2_, ps2_);
                if (a != c) throw std::string("the two mate files hold different numbers of records");
                if (!a) break;
                pending_ = true;
            }
            if (ps1_.size() > b.cap || ps2_.size() > b.cap) throw std::string("a read is longer than the staging buffer");
            if (b.off1.back() + ps1_.size() > b.cap || b.off2.back() + ps2_.size() > b.cap) break;
            memcpy(b.b1 + b.off1.back(), ps1_.data(), ps1_.size());
            memcpy(b.b2 + b.off2.back(), ps2_.data(), ps2_.size());
            b.off1.push_back(b.off1.back() + ps1_.size());
            b.off2.push_back(b.off2.back() + ps2_.size());
            b.h1.push_back(ph1_); b.h2.push_back(ph2_);
            ++b.n;
            pending_ = false;
        }
        return b.n > 0;
    }

private:
    SeqStream s1_, s2_;
    std::string ph1_, ph2_, ps1_, ps2_;
    bool pending_ = false;
};

inline void put_record(FILE* f, const std::string& header_with_marker, const char* seq, size_t len) {
    fwrite(header_with_marker.data(), 1, header_with_marker.size(), f);
    fputc('\n', f);
    fwrite(seq, 1, len, f);
    fputc('\n', f);
}

// The fragment-length prior nobody updates in Circall_wt / Circall_bsj: getNormalFragLengthCounts
// (Circall_bsj.cpp:832-861: mean 200, sd 80, 10 000 samples over 1000 bins) and the conditional means behind
// computeSmoothedEffectiveLengths (:977-1006).  These columns are constants of a run and not read downstream.
struct FragPrior {
    // Nobody decrements remainingFLOps in Circall_wt / Circall_bsj, so both tools always take the normal prior (mean 200, sd 80, 10 000
    // samples over 1000 lengths; src/Circall_bsj.cpp:1171-1181 and its options :1350,1374-1381):
    //   counts   getNormalFragLengthCounts (:832-861)          -> ReadExperiment::setFragLengthDist (include/ReadExperiment.hpp:189-196)
    //            -> EmpiricalDistribution::buildDistribution (src/EmpiricalDistribution.cpp:36-113): median / mean / sd of fragmentInfo.txt,
    //               handed out as float (:128-146)
    //   factors  getNormalFragLengthDist (:805-830), from the kernel itself, not from the rounded counts -> computeSmoothedEffectiveLengths
    //            (:977-1006): the effLens / EffectiveLength columns of rawCount.txt
    std::vector<uint32_t> counts;
    std::vector<double> factors;
    float median = 0, mean = 0, sd = 0;
    FragPrior(int maxLen = 1000, int samples = 10000, size_t mu = 200, size_t sigma = 80) : counts(maxLen, 0), factors(maxLen, 0.0) {
        auto kernel = [&](double p) { double inv = 1.0 / sigma, x = inv * (p - mu); return std::exp(-0.5 * x * x) * inv; };
        double mass = 0;
        for (int i = 0; i < maxLen; ++i) mass += kernel((double)i);
        if (mass > 0) for (int i = 0; i < maxLen; ++i) counts[i] = (uint32_t)(int)std::round(kernel((double)i) * samples / mass);
        double cm = 0, cd = 0;
        for (int i = 0; i < maxLen; ++i) {                          // :819-828
            const double d = kernel((double)i);
            cm += i * d; cd += d;
            if (cd > 0) factors[i] = cm / cd;
        }
        // EmpiricalDistribution::buildDistribution with vals = 0 .. maxLen - 1, lens = counts
        const size_t n = counts.size();
        double valsum = 0;
        for (size_t i = 0; i < n; ++i) valsum += counts[i];
        double cumpr = 0;
        unsigned lastval = 0;
        for (; lastval < n; ++lastval) { cumpr += counts[lastval] / valsum; if (cumpr > 1.0 - 1e-6) break; }      // :52-59
        const unsigned upper = std::min<unsigned>(lastval + 1, (unsigned)n);
        valsum = 0;
        for (unsigned i = 0; i < upper; ++i) valsum += counts[i];
        std::vector<double> pdf(upper);
        for (unsigned v = 0; v < upper; ++v) pdf[v] = counts[v] / valsum;                                           // :68-78 (vals[i] == i)
        double m = 0;
        for (unsigned v = 1; v < upper; ++v) m += v * pdf[v];                                                       // :82-86
        double s2 = 0;
        for (unsigned i = 0; i < upper; ++i) s2 += (i - m) * (i - m) * pdf[i];                                      // :90-94
        size_t i = 0, j = n - 1;                                                                                    // :98-110: the median walks in from both ends
        unsigned u = counts[0], v = counts[n - 1];
        while (i < j) {
            if (u <= v) { v -= u; u = counts[++i]; }
            else { u -= v; v = counts[--j]; }
        }
        median = (float)i; mean = (float)m; sd = (float)std::sqrt(s2);
    }
    double effective_length(uint32_t refLen) const {               // :990-1003
        int maxLen = (int)factors.size();
        double cf = (int)refLen >= maxLen ? factors[maxLen - 1] : factors[refLen];
        double e = (double)refLen - cf + 1.0;
        return e < 1.0 ? (double)refLen : e;
    }
};

// fragmentInfo.txt, paired-end form (ExportFeq.cpp:81-82)
inline bool write_fragment_info(const std::string& path, uint32_t readlen, const FragPrior& fp, uint64_t observed, uint64_t mapped, uint64_t hits, int k) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    fprintf(f, "readlen\tfragLengthMedian\tfragLengthMean\tfragLengthSd\tnumObservedFragments\tnumMappedFragments\tnumHits\tkmer\n");
    fprintf(f, "%u\t%g\t%g\t%g\t%llu\t%llu\t%llu\t%d\n", readlen, fp.median, fp.mean, fp.sd, (unsigned long long)observed, (unsigned long long)mapped,
            (unsigned long long)hits, k);
    fclose(f);
    return true;
}

// equivalence classes accumulated over the batches of a run: key = ascending transcript ids (EquivalenceClassBuilder.hpp:112-130)
typedef std::map<std::vector<uint32_t>, uint64_t> EqClasses;

}  // namespace cgh
