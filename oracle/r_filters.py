"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no R in this image: this is a reading of the R code, not a run of it).

Python restatement of the R stage that turns the mappers' files into `junction_fragment_count` / `junction_FPM`
(SURVEY.md section 8, row N4).  Paths relative to /root/reference/Circall_v1.0.1/:

  get_circInfo      R/Circall_functions.R:631-644     circRNA name -> chr / start / end / gene / exon ids
  get_circLen       R/Circall_functions.R:490-608     circRNA lengths from the gene model's exon table
  SE_filtering      R/Circall_functions.R:12-93       junction table x rawCount.txt -> candidates of the single-end stage
  PE_filter         R/Circall_functions.R:135-239     + Circall_pseudo's hit lists -> fragment_count, tandem_unique_count
  export            R/export.R:10-24                  final table (junction_fragment_count, junction_FPM, median_circlen)

The gene model enters as a flat exon table (what process_Sqlite, :610-617, selects from the TxDb: GENEID, TXNAME, EXONSTART,
EXONEND) instead of the GenomicFeatures sqlite file, which nothing here can read.  Only tests/ may import this module.
"""
import os
import statistics

NA = None
SUFFIXES = ("___11", "___12", "___21", "___22")


def norm_id(name):
    """text before the third "__" (Circall_functions.R:32-36); R's substring(x, 1, NA - 1) on a name with fewer separators is NA"""
    p, pos = -2, []
    for _ in range(3):
        p = name.find("__", p + 2)
        if p < 0:
            return NA
        pos.append(p)
    return name[:pos[2]]


def cut_before_sixth(name):
    """text before the sixth "__" (Circall_functions.R:179-184)"""
    p = -2
    for _ in range(6):
        p = name.find("__", p + 2)
        if p < 0:
            return NA
    return name[:p]


def get_circ_info(names):
    """:631-644 — six "__"-separated fields"""
    out = []
    for n in names:
        f = n.split("__")
        out.append(dict(Chr=f[0], start=int(f[1]), end=int(f[2]), GENEID=f[3], start_EXONID=f[4], end_EXONID=f[5]))
    return out


def get_circ_len(exons, circ_info):
    """:490-608.  exons: list of (GENEID, TXNAME, EXONSTART, EXONEND) in table order.  Returns rows in circ_info order."""
    by_gene = {}
    for g, tx, s, e in exons:
        by_gene.setdefault(g, []).append((tx, int(s), int(e)))
    res = []
    for c in circ_info:
        nid = "%s__%d__%d" % (c["Chr"], c["start"], c["end"])
        ex = by_gene.get(c["GENEID"], [])
        start, end = c["start"], c["end"]
        tx1, tx2 = [], []
        for tx, s, e in ex:                                   # unique() keeps first occurrences, intersect() the order of its first argument
            if s == start and tx not in tx1:
                tx1.append(tx)
            if e == end and tx not in tx2:
                tx2.append(tx)
        mytx = [t for t in tx1 if t in tx2]
        lens, exon_num, s_len, e_len = [], [], NA, NA
        if mytx:
            for atx in mytx:
                te = sorted([(s, e) for tx, s, e in ex if tx == atx], key=lambda x: x[0])     # order(): stable
                pick = [(s, e) for s, e in te if s >= start and e <= end]
                lens.append(sum(e - s + 1 for s, e in pick))
                exon_num.append(len(pick))
                s_len = pick[0][1] - pick[0][0] + 1 if pick else NA      # of the last transcript looked at (overwritten every turn)
                e_len = pick[-1][1] - pick[-1][0] + 1 if pick else NA
        else:
            inside = sorted([(s, e) for tx, s, e in ex if s >= start and e <= end], key=lambda x: x[0])
            if not inside:
                lens = [0]            # (R leaves exon_num / *_exon_len NULL here and c() then drops them: a malformed row; NA in this restatement)
            else:
                n = len(inside)
                cl = [-1] * n
                for j in range(n):
                    if cl[j] == -1:
                        cl[j] = j
                        sj, ej = inside[j]
                        pick = [k for k in range(n) if (sj - inside[k][0]) * (sj - inside[k][1]) <= 0 or (ej - inside[k][0]) * (ej - inside[k][1]) <= 0]
                        x = min(cl[k] for k in pick if cl[k] >= 0)       # R indices are 1-based and tests x > 0: here 0-based, >= 0
                        for k in pick:
                            cl[k] = x
                while True:
                    nxt = [cl[c2] for c2 in cl]
                    if nxt == cl:
                        break
                    cl = nxt
                regs = sorted((min(inside[k][0] for k in range(n) if cl[k] == cid), max(inside[k][1] for k in range(n) if cl[k] == cid))
                              for cid in sorted(set(cl)))
                el = [e - s + 1 for s, e in regs]
                lens = [sum(el)]
                exon_num = [len(el)]
                s_len, e_len = el[0], el[-1]
        res.append(dict(normID=nid, min_circlen=float(min(lens)), mean_circlen=sum(lens) / len(lens), median_circlen=float(statistics.median(lens)),
                        max_circlen=float(max(lens)), txNum=len(mytx), start_exon_len=s_len, end_exon_len=e_len,
                        max_exon_num=max(exon_num) if exon_num else NA))
    return res


def read_junction_tables(bs_dir):
    """list.files(pattern = "UN_read") in name order, rbind (:14-22)"""
    rows = []
    for fn in sorted(os.listdir(bs_dir)):
        if "UN_read" not in fn:
            continue
        for line in open(os.path.join(bs_dir, fn)):
            f = line.rstrip("\n").split("\t")
            if len(f) != 7:
                continue
            rows.append(dict(Header=f[0], Direction=int(f[1]), QueryPos=int(f[2]), HitLen=int(f[3]), HitPos=int(f[4]), Txname=f[5], Reflen=int(f[6])))
    return rows


def read_tsv(path):
    lines = open(path).read().split("\n")
    hdr = lines[0].split("\t")
    return [dict(zip(hdr, ln.split("\t"))) for ln in lines[1:] if ln]


def se_filtering(bs_dir, exons, read_map_prop=0.99, read_anchor_len=10, eqc_tx_min_prop=1, eqc_min_count=1, hard_hitlen=50):
    """:12-93.  Returns dict(read_raw, read_dat, eqc_dat, circRNA_names)."""
    frag = read_tsv(os.path.join(bs_dir, "fragmentInfo.txt"))[0]
    readlen = float(frag["readlen"])
    rd = read_junction_tables(bs_dir)
    uniq = []
    seen = set()
    for r in rd:
        if r["Txname"] not in seen:
            seen.add(r["Txname"]); uniq.append(r["Txname"])
    circ_len = get_circ_len(exons, get_circ_info(uniq))
    first = {}
    for row in circ_len:                                              # match() returns the first row with that normID
        first.setdefault(row["normID"], row)
    for r in rd:
        r["dis5"] = r["Reflen"] / 2 - r["HitPos"]                     # :24-25
        r["dis3"] = r["HitPos"] + r["HitLen"] - r["Reflen"] / 2
        r["normID"] = norm_id(r["Txname"])
        cl = first.get(r["normID"])
        r["circ"] = cl
        mc = cl["min_circlen"] if cl else NA
        if cl and cl["max_exon_num"] == 1:                            # :41
            mc = mc * 2
        r["min_circlen"] = mc
        r["readlen"] = readlen
        r["f_reflen"] = r["Reflen"] / 2 + read_anchor_len             # :43
        if mc is NA:
            r["min_hitlength"] = NA
        else:
            mh = read_map_prop * min(readlen, r["f_reflen"], mc)      # :44
            r["min_hitlength"] = mh if mh >= hard_hitlen else float(hard_hitlen)   # :45
    read_raw = [dict(r) for r in rd]
    # :55-57 (rows whose comparison is NA are dropped here; R would keep them as all-NA rows)
    kept = [r for r in rd if r["min_hitlength"] is not NA and r["HitLen"] >= r["min_hitlength"]]
    kept = [r for r in kept if r["dis3"] > read_anchor_len and r["dis5"] > read_anchor_len]
    eqc = read_tsv(os.path.join(bs_dir, "rawCount.txt"))
    for e in eqc:
        e["Weight"] = float(e["Weight"]); e["Count"] = int(e["Count"])
    eqc = [e for e in eqc if e["Weight"] >= eqc_tx_min_prop and e["Count"] >= eqc_min_count]     # :60
    names = {r["Txname"] for r in kept}
    eqc = [e for e in eqc if e["Transcript"] in names]                                             # :63-64
    eqc.sort(key=lambda e: -e["Count"])                                                            # :67 (stable)
    cnt = {}
    for r in kept:
        cnt[r["Txname"]] = cnt.get(r["Txname"], 0) + 1
    for e in eqc:
        e["normID"] = norm_id(e["Transcript"])
        e["circ"] = first.get(e["normID"])
        e["median_circlen"] = e["circ"]["median_circlen"] if e["circ"] else NA
        e["fragment_count_SE"] = cnt.get(e["Transcript"], NA)
    circ_names = []
    for e in eqc:
        if e["Transcript"] not in circ_names:
            circ_names.append(e["Transcript"])
    return dict(read_raw=read_raw, read_dat=kept, eqc_dat=eqc, circRNA_names=circ_names, fragInfo=frag)


def strip_suffix(h):
    for s in SUFFIXES:                    # gsub("___11|___12|___21|___22", "", x): every occurrence
        h = h.replace(s, "")
    return h


def pe_filter(pe_headers, pe_txnames, res1, num_observed_fragments, read_anchor_len=10, eqc_tx_min_prop=1, min_fragment_count=1):
    """:135-239.  pe_headers / pe_txnames: the lines of merged_hitheader.txt / merged_mapInfo.txt (:118-125); res1: se_filtering's
    result.  Returns dict(eqc_dat, circRNAs_candidates, final) — final as R/export.R:10-24 writes it."""
    SE = []
    for r in res1["read_raw"]:
        if r["dis3"] > read_anchor_len and r["dis5"] > read_anchor_len:                            # :152-153
            SE.append(dict(Header=strip_suffix(r["Header"]), Txname=r["Txname"], HitLen=r["HitLen"]))
    se_headers = {x["Header"] for x in SE}
    se_key = {}
    for x in SE:
        se_key.setdefault(x["Header"] + "-" + x["Txname"], x["HitLen"])                            # match(): first
    PE = [(strip_suffix(h), t) for h, t in zip(pe_headers, pe_txnames)]
    PE = [(h, t) for h, t in PE if h in se_headers]                                                # :160
    td_headers = {h for h, t in PE if "++" not in t}                                               # :163-164
    rows = []
    for h, t in PE:                                                                                # :168-186
        for name in t.split(" "):
            if name == "":
                continue
            cut = cut_before_sixth(name)
            rows.append(dict(Header=h, TxnamePE0=name, TxnamePE=cut, key=h + "-" + (cut if cut is not NA else "NA")))
    rows = [r for r in rows if r["key"] in se_key]                                                 # :191-192
    seen, ded = set(), []
    for r in rows:                                                                                 # :199
        if r["key"] not in seen:
            seen.add(r["key"]); ded.append(r)
    pe_td = [r for r in ded if r["Header"] in td_headers]                                          # :202-205
    pe = [r for r in ded if r["Header"] not in td_headers]
    frag, tand = {}, {}
    for r in pe:
        frag[r["TxnamePE"]] = frag.get(r["TxnamePE"], 0) + 1
    for r in pe_td:
        tand[r["TxnamePE"]] = tand.get(r["TxnamePE"], 0) + 1
    eqc = [dict(e) for e in res1["eqc_dat"]]
    for e in eqc:
        e["fragment_count"] = frag.get(e["Transcript"], NA)
        e["tandem_unique_count"] = tand.get(e["Transcript"], 0)
    eqc = [e for e in eqc if e["Transcript"] in frag]                                              # :221-222
    eqc = [e for e in eqc if e["Weight"] >= eqc_tx_min_prop and e["fragment_count"] >= min_fragment_count]     # :224-225
    eqc = [e for e in eqc if e["fragment_count"] >= 3 * e["tandem_unique_count"]]                  # :227-228
    for e in eqc:
        e["fragment_count_norm"] = e["fragment_count"] / num_observed_fragments * 10 ** 6          # :235
    cands = [e["Transcript"] for e in eqc]
    info = get_circ_info(cands)
    final = []
    for c, e in zip(info, eqc):                                                                    # export.R:10-20
        final.append(dict(chr=c["Chr"], start=c["start"], end=c["end"], geneID=c["GENEID"], circID=e["normID"],
                          junction_fragment_count=e["fragment_count"], junction_FPM=e["fragment_count_norm"], median_circlen=e["median_circlen"]))
    final.sort(key=lambda r: -r["junction_fragment_count"])                                        # order(decreasing = TRUE): stable
    return dict(eqc_dat=eqc, circRNAs_candidates=cands, final=final)


def fmt(v):
    """how the C++ tool prints a number: integers as such, other values with 15 significant digits, NA as NA"""
    if v is NA:
        return "NA"
    if isinstance(v, int) or (isinstance(v, float) and v == int(v) and abs(v) < 1e15):
        return str(int(v))
    return "%.15g" % v


def final_lines(final):
    out = ["chr\tstart\tend\tgeneID\tcircID\tjunction_fragment_count\tjunction_FPM\tmedian_circlen"]
    for r in final:
        out.append("\t".join([r["chr"], fmt(r["start"]), fmt(r["end"]), r["geneID"], r["circID"], fmt(r["junction_fragment_count"]),
                              fmt(r["junction_FPM"]), fmt(r["median_circlen"])]))
    return out
