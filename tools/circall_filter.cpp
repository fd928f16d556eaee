// circall_b200 — the R stage between the mappers and the FDR model, in C++:
//
//   circall_filter se  --bs OUTDIR_BS --exons EXONS.tsv --out DIR            R/doSingleEndFiltering.R -> SE_filtering  (R/Circall_functions.R:12-93)
//   circall_filter pe  --se DIR --pseudo OUTDIR_PSEUDO --wt OUTDIR_WT --out FINAL.txt
//                                                                            R/doPairEndFiltering.R -> PE_filter (:135-239) + R/export.R:10-24
//
// `se` joins the junction table (UN_read*, written by Circall_bsj) with rawCount.txt and the circRNA lengths of the gene model
// (get_circInfo :631-644, get_circLen :490-608) and writes what SE_filtering keeps in its RData file as three TSV files
// (SE_read_raw.tsv, SE_read_dat.tsv, SE_eqc_dat.tsv) plus SE_circRNA_names.txt.  `pe` joins that with Circall_pseudo's hit lists
// (merged_hitheader.txt + merged_mapInfo.txt, or the Mapped_read_*.rh / mapInfo_*.txt parts in name order) and the wt run's
// numObservedFragments and writes the final table with junction_fragment_count and junction_FPM.
//
// The gene model comes as a flat exon table with a header line and the columns GENEID, TXNAME, EXONSTART, EXONEND in any position
// (what process_Sqlite, :610-617, selects from the TxDb; `write.table(process_Sqlite(f), sep = "\t", quote = FALSE)` exports it):
// reading the GenomicFeatures sqlite file itself is not something this tool does.
//
// Host code only: these are joins over tables of a few thousand to a few million rows (the reference does them in R); the GPU
// path ends at the files this reads.  tests/test_filters.py compares every output with oracle/r_filters.py (a Python reading of
// the same R lines; no R in this image, so that oracle is unpinned).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

[[noreturn]] void die(const std::string& m) { fprintf(stderr, "[circall_filter] error: %s\n", m.c_str()); exit(1); }

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        size_t b = s.find(sep, a);
        if (b == std::string::npos) { out.push_back(s.substr(a)); return out; }
        out.push_back(s.substr(a, b - a));
        a = b + 1;
    }
}
std::vector<std::string> split_str(const std::string& s, const std::string& sep) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        size_t b = s.find(sep, a);
        if (b == std::string::npos) { out.push_back(s.substr(a)); return out; }
        out.push_back(s.substr(a, b - a));
        a = b + sep.size();
    }
}
std::vector<std::string> read_lines(const std::string& path) {
    std::ifstream in(path);
    if (!in) die("cannot open " + path);
    std::vector<std::string> out;
    std::string s;
    while (std::getline(in, s)) { if (!s.empty() && s.back() == '\r') s.pop_back(); out.push_back(s); }
    return out;
}
std::vector<std::string> list_dir(const std::string& d) {          // sorted, as list.files() and the shell's glob give them
    std::vector<std::string> out;
    DIR* h = opendir(d.c_str());
    if (!h) die("cannot open directory " + d);
    while (dirent* e = readdir(h)) if (e->d_name[0] != '.') out.push_back(e->d_name);
    closedir(h);
    std::sort(out.begin(), out.end());
    return out;
}
bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

// a value that may be NA
struct Num { double v = 0; bool na = true; Num() {} Num(double x) : v(x), na(false) {} };
std::string fmt(const Num& n) {
    if (n.na) return "NA";
    if (n.v == std::floor(n.v) && std::fabs(n.v) < 1e15) { char b[32]; snprintf(b, sizeof b, "%lld", (long long)n.v); return b; }
    char b[40]; snprintf(b, sizeof b, "%.15g", n.v); return b;
}

// text before the n-th "__" (Circall_functions.R:32-36 with n = 3, :179-184 with n = 6); false = fewer separators (R: NA)
bool cut_before(const std::string& x, int n, std::string& out) {
    size_t p = 0;
    for (int i = 0; i < n; ++i) {
        p = x.find("__", i == 0 ? 0 : p + 2);
        if (p == std::string::npos) return false;
    }
    out = x.substr(0, p);
    return true;
}

struct CircLen { std::string normID; Num min, mean, median, max, txNum, start_exon_len, end_exon_len, max_exon_num; };
struct Exon { std::string tx; long long s, e; };

// get_circInfo + get_circLen for one circRNA name (Circall_functions.R:631-644, 490-608)
CircLen circ_len(const std::string& name, const std::unordered_map<std::string, std::vector<Exon>>& by_gene) {
    std::vector<std::string> f = split_str(name, "__");
    if (f.size() < 6) die("BSJ reference name without six __-separated fields: " + name);
    const long long start = atoll(f[1].c_str()), end = atoll(f[2].c_str());
    CircLen r;
    r.normID = f[0] + "__" + std::to_string(start) + "__" + std::to_string(end);
    static const std::vector<Exon> none;
    auto it = by_gene.find(f[3]);
    const std::vector<Exon>& ex = it == by_gene.end() ? none : it->second;
    std::vector<std::string> tx1, tx2, mytx;
    for (const Exon& x : ex) {
        if (x.s == start && std::find(tx1.begin(), tx1.end(), x.tx) == tx1.end()) tx1.push_back(x.tx);
        if (x.e == end && std::find(tx2.begin(), tx2.end(), x.tx) == tx2.end()) tx2.push_back(x.tx);
    }
    for (const std::string& t : tx1) if (std::find(tx2.begin(), tx2.end(), t) != tx2.end()) mytx.push_back(t);
    std::vector<double> lens; std::vector<long long> exon_num;
    if (!mytx.empty()) {
        for (const std::string& atx : mytx) {                       // :514-527
            std::vector<std::pair<long long, long long>> te;
            for (const Exon& x : ex) if (x.tx == atx) te.push_back({x.s, x.e});
            std::stable_sort(te.begin(), te.end(), [](auto& a, auto& b) { return a.first < b.first; });
            long long sum = 0, n = 0, first = -1, last = -1;
            for (auto& p : te) if (p.first >= start && p.second <= end) { sum += p.second - p.first + 1; if (n == 0) first = p.second - p.first + 1; last = p.second - p.first + 1; ++n; }
            lens.push_back((double)sum); exon_num.push_back(n);
            r.start_exon_len = n ? Num((double)first) : Num(); r.end_exon_len = n ? Num((double)last) : Num();     // of the last transcript (overwritten every turn)
        }
    } else {                                                        // :529-591: exons of the gene inside the region, overlapping ones merged
        std::vector<std::pair<long long, long long>> in;
        for (const Exon& x : ex) if (x.s >= start && x.e <= end) in.push_back({x.s, x.e});
        std::stable_sort(in.begin(), in.end(), [](auto& a, auto& b) { return a.first < b.first; });
        if (in.empty()) lens.push_back(0);                          // (R's row is malformed here: NULLs dropped by c(); NA in this tool)
        else {
            const int n = (int)in.size();
            std::vector<int> cl(n, -1);
            for (int j = 0; j < n; ++j) {
                if (cl[j] != -1) continue;
                cl[j] = j;
                std::vector<int> pick; int x = n;
                for (int k = 0; k < n; ++k) {                       // exon k holds exon j's start or its end: R's (a - s_k) * (a - e_k) <= 0 with s_k <= e_k
                    const bool a = in[k].first <= in[j].first && in[j].first <= in[k].second;
                    const bool b = in[k].first <= in[j].second && in[j].second <= in[k].second;
                    if (a || b) { pick.push_back(k); if (cl[k] >= 0) x = std::min(x, cl[k]); }
                }
                for (int k : pick) cl[k] = x;
            }
            for (;;) { std::vector<int> nx(n); for (int k = 0; k < n; ++k) nx[k] = cl[cl[k]]; if (nx == cl) break; cl = nx; }
            std::map<int, std::pair<long long, long long>> reg;
            for (int k = 0; k < n; ++k) {
                auto f2 = reg.find(cl[k]);
                if (f2 == reg.end()) reg[cl[k]] = in[k];
                else { f2->second.first = std::min(f2->second.first, in[k].first); f2->second.second = std::max(f2->second.second, in[k].second); }
            }
            std::vector<std::pair<long long, long long>> regs;
            for (auto& kv : reg) regs.push_back(kv.second);
            std::stable_sort(regs.begin(), regs.end(), [](auto& a, auto& b) { return a.first < b.first; });
            long long sum = 0;
            for (auto& p : regs) sum += p.second - p.first + 1;
            lens.push_back((double)sum); exon_num.push_back((long long)regs.size());
            r.start_exon_len = Num((double)(regs.front().second - regs.front().first + 1));
            r.end_exon_len = Num((double)(regs.back().second - regs.back().first + 1));
        }
    }
    std::vector<double> srt = lens;
    std::sort(srt.begin(), srt.end());
    double sum = 0; for (double v : lens) sum += v;
    r.min = Num(srt.front()); r.max = Num(srt.back()); r.mean = Num(sum / (double)lens.size());
    r.median = Num(srt.size() % 2 ? srt[srt.size() / 2] : 0.5 * (srt[srt.size() / 2 - 1] + srt[srt.size() / 2]));
    r.txNum = Num((double)mytx.size());
    if (!exon_num.empty()) r.max_exon_num = Num((double)*std::max_element(exon_num.begin(), exon_num.end()));
    return r;
}

std::unordered_map<std::string, std::vector<Exon>> load_exons(const std::string& path) {
    std::vector<std::string> L = read_lines(path);
    if (L.empty()) die("empty exon table " + path);
    std::vector<std::string> h = split(L[0], '\t');
    int cg = -1, ct = -1, cs = -1, ce = -1;
    for (size_t i = 0; i < h.size(); ++i) { if (h[i] == "GENEID") cg = (int)i; if (h[i] == "TXNAME") ct = (int)i; if (h[i] == "EXONSTART") cs = (int)i; if (h[i] == "EXONEND") ce = (int)i; }
    if (cg < 0 || ct < 0 || cs < 0 || ce < 0) die("the exon table needs the columns GENEID, TXNAME, EXONSTART, EXONEND");
    std::unordered_map<std::string, std::vector<Exon>> out;
    for (size_t i = 1; i < L.size(); ++i) {
        if (L[i].empty()) continue;
        std::vector<std::string> f = split(L[i], '\t');
        if ((int)f.size() <= std::max(std::max(cg, ct), std::max(cs, ce))) die("short line in the exon table: " + L[i]);
        out[f[cg]].push_back(Exon{f[ct], atoll(f[cs].c_str()), atoll(f[ce].c_str())});
    }
    return out;
}

struct Tsv { std::vector<std::string> hdr; std::vector<std::vector<std::string>> rows; int col(const std::string& n) const { for (size_t i = 0; i < hdr.size(); ++i) if (hdr[i] == n) return (int)i; die("missing column " + n); } };
Tsv load_tsv(const std::string& path) {
    std::vector<std::string> L = read_lines(path);
    if (L.empty()) die("empty table " + path);
    Tsv t; t.hdr = split(L[0], '\t');
    for (size_t i = 1; i < L.size(); ++i) if (!L[i].empty()) t.rows.push_back(split(L[i], '\t'));
    return t;
}

struct Read {                                                       // one junction-table row with SE_filtering's derived columns
    std::string header, tx, normID; bool hasNorm = false;
    long long dir, qpos, hitlen, hitpos, reflen;
    double dis5, dis3, f_reflen, readlen;
    const CircLen* cl = nullptr;
    Num min_circlen, min_hitlength;
};
const char* CIRC_COLS = "normID.1\tmin.circlen\tmean.circlen\tmedian.circlen\tmax.circlen\ttxNum\tstart_exon_len\tend_exon_len\tmax_exon_num";
std::string circ_cells(const CircLen* c) {
    if (!c) return "NA\tNA\tNA\tNA\tNA\tNA\tNA\tNA\tNA";
    return c->normID + "\t" + fmt(c->min) + "\t" + fmt(c->mean) + "\t" + fmt(c->median) + "\t" + fmt(c->max) + "\t" + fmt(c->txNum) + "\t" + fmt(c->start_exon_len) + "\t" +
           fmt(c->end_exon_len) + "\t" + fmt(c->max_exon_num);
}
void write_reads(const std::string& path, const std::vector<const Read*>& R) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) die("cannot write " + path);
    fprintf(f, "Header\tDirection\tQueryPos\tHitLen\tHitPos\tTxname\tReflen\tdis5\tdis3\tnormID\t%s\tmin.circlen.adj\treadlen\tf_reflen\tmin_hitlength\n", CIRC_COLS);
    for (const Read* r : R)
        fprintf(f, "%s\t%lld\t%lld\t%lld\t%lld\t%s\t%lld\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n", r->header.c_str(), r->dir, r->qpos, r->hitlen, r->hitpos, r->tx.c_str(), r->reflen,
                fmt(Num(r->dis5)).c_str(), fmt(Num(r->dis3)).c_str(), r->hasNorm ? r->normID.c_str() : "NA", circ_cells(r->cl).c_str(), fmt(r->min_circlen).c_str(),
                fmt(Num(r->readlen)).c_str(), fmt(Num(r->f_reflen)).c_str(), fmt(r->min_hitlength).c_str());
    fclose(f);
}

struct Opt { std::map<std::string, std::string> kv; std::string get(const std::string& k, const std::string& d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; } };
Opt parse(int argc, char** argv, int first) {
    Opt o;
    for (int i = first; i < argc; ++i) {
        std::string a = argv[i];
        if (a.rfind("--", 0) != 0 || i + 1 >= argc) die("usage: circall_filter {se|pe} --option value ...  (bad argument " + a + ")");
        o.kv[a.substr(2)] = argv[++i];
    }
    return o;
}

// ---- SE_filtering (Circall_functions.R:12-93) ---------------------------------------------------------------------------
int run_se(const Opt& o) {
    const std::string bs = o.get("bs"), out = o.get("out"), exf = o.get("exons");
    if (bs.empty() || out.empty() || exf.empty()) die("se needs --bs OUTDIR_BS --exons EXONS.tsv --out DIR");
    const double mapProp = atof(o.get("read_mapProp", "0.99").c_str()), anchor = atof(o.get("read_anchorLen", "10").c_str());
    const double txMinProp = atof(o.get("eqc_txMinProp", "1").c_str()), minCount = atof(o.get("eqc_minCount", "1").c_str()), hard = atof(o.get("hard_hitlen", "50").c_str());
    mkdir(out.c_str(), 0777);
    Tsv frag = load_tsv(bs + "/fragmentInfo.txt");
    if (frag.rows.empty()) die("fragmentInfo.txt has no row");
    const double readlen = atof(frag.rows[0][frag.col("readlen")].c_str());
    auto by_gene = load_exons(exf);
    std::vector<Read> rd;
    for (const std::string& fn : list_dir(bs)) {                    // :14-22
        if (fn.find("UN_read") == std::string::npos) continue;
        for (const std::string& ln : read_lines(bs + "/" + fn)) {
            std::vector<std::string> f = split(ln, '\t');
            if (f.size() != 7) continue;
            Read r; r.header = f[0]; r.dir = atoll(f[1].c_str()); r.qpos = atoll(f[2].c_str()); r.hitlen = atoll(f[3].c_str()); r.hitpos = atoll(f[4].c_str());
            r.tx = f[5]; r.reflen = atoll(f[6].c_str());
            rd.push_back(r);
        }
    }
    // circRNA lengths per distinct reference name, first occurrence per normID wins (match(), :38)
    std::vector<std::string> uniq; std::unordered_set<std::string> seen;
    for (const Read& r : rd) if (seen.insert(r.tx).second) uniq.push_back(r.tx);
    std::vector<CircLen> cls; cls.reserve(uniq.size());
    for (const std::string& n : uniq) cls.push_back(circ_len(n, by_gene));
    std::unordered_map<std::string, const CircLen*> first;
    for (const CircLen& c : cls) first.emplace(c.normID, &c);
    for (Read& r : rd) {
        r.dis5 = (double)r.reflen / 2 - (double)r.hitpos;           // :24-25
        r.dis3 = (double)r.hitpos + (double)r.hitlen - (double)r.reflen / 2;
        r.hasNorm = cut_before(r.tx, 3, r.normID);
        auto it = r.hasNorm ? first.find(r.normID) : first.end();
        r.cl = it == first.end() ? nullptr : it->second;
        if (r.cl) { r.min_circlen = r.cl->min; if (!r.cl->max_exon_num.na && r.cl->max_exon_num.v == 1) r.min_circlen = Num(r.cl->min.v * 2); }      // :41
        r.readlen = readlen;
        r.f_reflen = (double)r.reflen / 2 + anchor;                 // :43
        if (!r.min_circlen.na) {
            double mh = mapProp * std::min(std::min(readlen, r.f_reflen), r.min_circlen.v);   // :44
            r.min_hitlength = Num(mh >= hard ? mh : hard);          // :45
        }
    }
    std::vector<const Read*> raw, kept;
    for (const Read& r : rd) raw.push_back(&r);
    for (const Read& r : rd)                                        // :55-57 (a row whose threshold is NA is dropped; R keeps it as an all-NA row)
        if (!r.min_hitlength.na && (double)r.hitlen >= r.min_hitlength.v && r.dis3 > anchor && r.dis5 > anchor) kept.push_back(&r);
    write_reads(out + "/SE_read_raw.tsv", raw);
    write_reads(out + "/SE_read_dat.tsv", kept);
    // equivalence classes (:49,:60-84)
    Tsv eq = load_tsv(bs + "/rawCount.txt");
    const int cT = eq.col("Transcript"), cW = eq.col("Weight"), cC = eq.col("Count");
    std::unordered_map<std::string, long long> cnt;
    for (const Read* r : kept) ++cnt[r->tx];
    std::vector<const std::vector<std::string>*> rows;
    for (const auto& r : eq.rows)
        if (atof(r[cW].c_str()) >= txMinProp && atof(r[cC].c_str()) >= minCount && cnt.count(r[cT])) rows.push_back(&r);
    std::stable_sort(rows.begin(), rows.end(), [&](auto a, auto b) { return atof((*a)[cC].c_str()) > atof((*b)[cC].c_str()); });      // :67
    FILE* f = fopen((out + "/SE_eqc_dat.tsv").c_str(), "w");
    if (!f) die("cannot write SE_eqc_dat.tsv");
    for (size_t i = 0; i < eq.hdr.size(); ++i) fprintf(f, "%s\t", eq.hdr[i].c_str());
    fprintf(f, "normID\tmedian.circlen\t%s\tfragment_count_SE\n", CIRC_COLS);
    std::vector<std::string> names; std::unordered_set<std::string> nseen;
    for (auto r : rows) {
        for (const std::string& c : *r) fprintf(f, "%s\t", c.c_str());
        std::string nid; const bool has = cut_before((*r)[cT], 3, nid);
        auto it = has ? first.find(nid) : first.end();
        const CircLen* c = it == first.end() ? nullptr : it->second;
        fprintf(f, "%s\t%s\t%s\t%lld\n", has ? nid.c_str() : "NA", c ? fmt(c->median).c_str() : "NA", circ_cells(c).c_str(), cnt[(*r)[cT]]);
        if (nseen.insert((*r)[cT]).second) names.push_back((*r)[cT]);
    }
    fclose(f);
    f = fopen((out + "/SE_circRNA_names.txt").c_str(), "w");
    if (!f) die("cannot write SE_circRNA_names.txt");
    for (const std::string& n : names) fprintf(f, "%s\n", n.c_str());
    fclose(f);
    fprintf(stderr, "[circall_filter] se: %zu junction rows, %zu kept, %zu candidate circRNAs\n", rd.size(), kept.size(), names.size());
    return 0;
}

std::string strip_suffix(std::string h) {                           // gsub("___11|___12|___21|___22", "", x)
    for (const char* s : {"___11", "___12", "___21", "___22"}) {
        size_t p;
        while ((p = h.find(s)) != std::string::npos) h.erase(p, 5);
    }
    return h;
}

// ---- PE_filter (Circall_functions.R:135-239) + export.R:10-24 -----------------------------------------------------------
int run_pe(const Opt& o) {
    const std::string se = o.get("se"), ps = o.get("pseudo"), wt = o.get("wt"), out = o.get("out");
    if (se.empty() || ps.empty() || wt.empty() || out.empty()) die("pe needs --se DIR --pseudo OUTDIR_PSEUDO --wt OUTDIR_WT --out FINAL.txt");
    const double anchor = atof(o.get("read_anchorLen", "10").c_str()), txMinProp = atof(o.get("eqc_txMinProp", "1").c_str());
    const double minFrag = atof(o.get("min_fragment_count", "1").c_str());
    Tsv fw = load_tsv(wt + "/fragmentInfo.txt");
    if (fw.rows.empty()) die("the wt fragmentInfo.txt has no row");
    const double nobs = atof(fw.rows[0][fw.col("numObservedFragments")].c_str());
    // hit lists of Circall_pseudo: the merged files when the shell made them, else the parts in name order (Circall_source.sh:343-346)
    std::vector<std::string> hdr, map;
    if (exists(ps + "/merged_hitheader.txt")) { hdr = read_lines(ps + "/merged_hitheader.txt"); map = read_lines(ps + "/merged_mapInfo.txt"); }
    else for (const std::string& fn : list_dir(ps)) {
        if (fn.size() > 3 && fn.compare(fn.size() - 3, 3, ".rh") == 0) { auto l = read_lines(ps + "/" + fn); hdr.insert(hdr.end(), l.begin(), l.end()); }
        if (fn.rfind("mapInfo_", 0) == 0) { auto l = read_lines(ps + "/" + fn); map.insert(map.end(), l.begin(), l.end()); }
    }
    if (hdr.size() != map.size()) die("hit header list and map info list differ in length");
    // single-end side: read_raw rows that clear the anchor on both sides (:148-154)
    Tsv raw = load_tsv(se + "/SE_read_raw.tsv");
    const int rH = raw.col("Header"), rT = raw.col("Txname"), r5 = raw.col("dis5"), r3 = raw.col("dis3");
    std::unordered_set<std::string> seHeaders, seKey;
    for (const auto& r : raw.rows) {
        if (!(atof(r[r3].c_str()) > anchor && atof(r[r5].c_str()) > anchor)) continue;
        const std::string h = strip_suffix(r[rH]);
        seHeaders.insert(h); seKey.insert(h + "-" + r[rT]);
    }
    struct PE { std::string header, cut, key; };
    std::vector<PE> rows;
    std::unordered_set<std::string> td;                             // reads whose hits hold no circRNA ("++") sequence: tandem-unique (:163-164)
    for (size_t i = 0; i < hdr.size(); ++i) {
        const std::string h = strip_suffix(hdr[i]);
        if (!seHeaders.count(h)) continue;                          // :160
        if (map[i].find("++") == std::string::npos) td.insert(h);
        for (const std::string& name : split(map[i], ' ')) {        // :168-186
            if (name.empty()) continue;
            std::string cut; const bool ok = cut_before(name, 6, cut);
            rows.push_back(PE{h, ok ? cut : "NA", h + "-" + (ok ? cut : "NA")});
        }
    }
    std::unordered_set<std::string> dedup;
    std::unordered_map<std::string, long long> frag, tand;
    for (const PE& r : rows) {
        if (!seKey.count(r.key)) continue;                          // :191-192
        if (!dedup.insert(r.key).second) continue;                  // :199
        if (td.count(r.header)) ++tand[r.cut]; else ++frag[r.cut];  // :202-208
    }
    Tsv eq = load_tsv(se + "/SE_eqc_dat.tsv");
    const int cT = eq.col("Transcript"), cW = eq.col("Weight"), cN = eq.col("normID"), cM = eq.col("median.circlen");
    struct Fin { std::string chr, start, end, gene, circ, median; long long n; double fpm; };
    std::vector<Fin> fin;
    for (const auto& r : eq.rows) {
        auto it = frag.find(r[cT]);
        if (it == frag.end()) continue;                             // :221-222
        const long long fc = it->second, tc = tand.count(r[cT]) ? tand[r[cT]] : 0;
        if (!(atof(r[cW].c_str()) >= txMinProp && (double)fc >= minFrag)) continue;      // :224-225
        if (!(fc >= 3 * tc)) continue;                              // :227-228
        std::vector<std::string> f = split_str(r[cT], "__");        // get_circInfo, doPairEndFiltering.R:28
        if (f.size() < 6) die("candidate name without six fields: " + r[cT]);
        fin.push_back(Fin{f[0], std::to_string(atoll(f[1].c_str())), std::to_string(atoll(f[2].c_str())), f[3], r[cN], r[cM], fc, (double)fc / nobs * 1e6});      // :235
    }
    std::stable_sort(fin.begin(), fin.end(), [](const Fin& a, const Fin& b) { return a.n > b.n; });           // export.R:14
    FILE* f = fopen(out.c_str(), "w");
    if (!f) die("cannot write " + out);
    fprintf(f, "chr\tstart\tend\tgeneID\tcircID\tjunction_fragment_count\tjunction_FPM\tmedian_circlen\n");   // export.R:15-16
    for (const Fin& x : fin) fprintf(f, "%s\t%s\t%s\t%s\t%s\t%lld\t%s\t%s\n", x.chr.c_str(), x.start.c_str(), x.end.c_str(), x.gene.c_str(), x.circ.c_str(), x.n, fmt(Num(x.fpm)).c_str(), x.median.c_str());
    fclose(f);
    fprintf(stderr, "[circall_filter] pe: %zu hit lists, %zu circRNAs in the final table\n", hdr.size(), fin.size());
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) die("usage: circall_filter {se|pe} --option value ...");
    const std::string cmd = argv[1];
    Opt o = parse(argc, argv, 2);
    if (cmd == "se") return run_se(o);
    if (cmd == "pe") return run_pe(o);
    die("unknown command " + cmd);
}
