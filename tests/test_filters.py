"""Row N4 (SURVEY.md section 8 f): the R stage's joins in C++ (tools/circall_filter.cpp: SE_filtering, PE_filter, export) against the
Python restatement of the same R lines (oracle/r_filters.py; unpinned: there is no R in this image), on the files the mappers write.

The inputs are produced here without a GPU: the oracle runs Circall_wt + Circall_bsj on a synthetic sample and its results are
written in the reference's file formats; the gene model is an exon table over the synthetic genes with made-up
isoforms, so that both branches of get_circLen run (an isoform holds both ends of the circle / none does and the exons are merged)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def flt():
    from tools import build
    build.build()
    return os.path.join(ROOT, "tools", "bin", "circall_filter")


def _write_inputs(tmp, T, otx, obsj, r1, r2):
    import oracle
    R = oracle.Run(otx, obsj, 2, r1, r2)
    n = r1.shape[0]
    h1 = ["read%d/1" % i for i in range(n)]
    bsn = [T.bsj_name(t) for t in range(T.n_bsj)]
    reflen = obsj.lens
    ep, ec = R.wt_export()
    suf1 = {1: "___11", 2: "___21"}
    bs, wt, ps = tmp / "bs", tmp / "wt", tmp / "pseudo"
    for d in (bs, wt, ps):
        os.makedirs(d)
    rows = R.bsj_junctions().tolist()
    for part, sl in enumerate((rows[: len(rows) // 2], rows[len(rows) // 2:])):            # two parts, as several workers would leave them
        with open(bs / ("UN_read1_%d.fa" % (1804289383 + part)), "w") as f:
            for rec, d, q, ln, hp, t in sl:
                f.write(">%s%s\t%d\t%d\t%d\t%d\t%s\t%d\n" % (h1[ep[rec]], suf1[int(ec[rec])], d, q, ln, hp, bsn[t], reflen[t]))
    ob, ow = R.counters(1), R.counters(0)
    hdr = "readlen\tfragLengthMedian\tfragLengthMean\tfragLengthSd\tnumObservedFragments\tnumMappedFragments\tnumHits\tkmer\n"
    open(bs / "fragmentInfo.txt", "w").write(hdr + "%d\t200\t200\t80\t%d\t%d\t%d\t31\n" % (ob["readLen"], ob["numObserved"], ob["numMapped"], ob["numHits"]))
    open(wt / "fragmentInfo.txt", "w").write(hdr + "%d\t200\t200\t80\t%d\t%d\t%d\t31\n" % (ow["readLen"], ow["numObserved"], ow["numMapped"], ow["numHits"]))
    with open(bs / "rawCount.txt", "w") as f:
        f.write("Transcript\tWeight\tCount\teffLens\tRefLength\tEffectiveLength\teqClass\n")
        for cls, (tids, cnt) in enumerate(sorted(R.eq_classes(1).items()), 1):
            for t in tids:
                f.write("%s\t%g\t%d\t%g\t%d\t%g\t%d\n" % (bsn[t], 1.0 / len(tids), cnt, 100.0, reflen[t], 100.0, cls))
    # Circall_pseudo's outputs for the candidate pairs: the pseudo-sequence names are the BSJ name + an isoform tag, "++" marking a
    # circRNA sequence and "--" a tandem one (Circall_functions.R:164,308,464); some reads hit only tandem sequences
    nh, sp = R.bsj_records()
    rng = np.random.default_rng(5)
    per_rec = {}
    for rec, d, q, ln, hp, t in rows:
        per_rec.setdefault(rec, []).append(t)
    with open(ps / "Mapped_read_1.rh", "w") as fh, open(ps / "mapInfo_1.txt", "w") as fm:
        for rec in R.bsj_candidates().tolist():
            tids = sorted(set(per_rec.get(rec, [])))
            if not tids or rng.random() < 0.1:
                continue
            tandem_only = rng.random() < 0.15
            names = []
            for t in tids:
                for iso in range(int(rng.integers(1, 3))):
                    names.append("%s__SYNT%d%s" % (bsn[t], iso, "--" if tandem_only or rng.random() < 0.2 else "++"))
            fh.write(">%s%s\n" % (h1[ep[rec]], suf1[int(ec[rec])]))
            fm.write("".join(x + " " for x in names) + "\n")
    # exon table: every gene's exons with made-up isoforms (all exons / the even ones / a middle stretch)
    exons = []
    with open(tmp / "exons.tsv", "w") as f:
        f.write("GENEID\tTXNAME\tEXONID\tEXONSTART\tEXONEND\n")
        for g in range(T.n_genes):
            ex = T.gene_exons(g)
            iso = [ex, ex[::2], ex[1:-1] if len(ex) > 3 else ex[:1]]
            iso = iso[: 1 + g % 3]
            if g % 4 == 3:              # no isoform with every exon: circles from an even to an odd exon have no transcript of their own, and one
                iso = [ex[::2], ex[1::2]]           # isoform carries a retained intron (an exon that overlaps two others: the merge of get_circLen)
                if len(ex) > 4:
                    iso.append([(ex[1][0], ex[2][1]), ex[3]])
            for k, lst in enumerate(iso):
                for j, (s, e) in enumerate(lst):
                    exons.append(("G%06u" % g, "G%06u.T%d" % (g, k), s, e))
                    f.write("G%06u\tG%06u.T%d\t%d\t%d\t%d\n" % (g, g, k, g * 100 + j, s, e))
    return bs, wt, ps, exons


def test_se_and_pe_filters_against_the_r_restatement(flt, tmp_path):
    import oracle
    import synth
    from oracle import r_filters as rf
    synth.build(); oracle.build()
    T = synth.Txome(31, 150)
    otx, obsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)
    r1, r2 = T.reads(8, 40000, L=100, circ_frac=0.3)
    bs, wt, ps, exons = _write_inputs(tmp_path, T, otx, obsj, r1, r2)
    se_dir = tmp_path / "se"
    subprocess.run([flt, "se", "--bs", str(bs), "--exons", str(tmp_path / "exons.tsv"), "--out", str(se_dir)], check=True)
    want = rf.se_filtering(str(bs), exons)
    assert len(want["read_raw"]) > 500 and len(want["read_dat"]) > 100 and len(want["circRNA_names"]) > 20

    def table(path):
        return rf.read_tsv(str(path))

    def same_reads(got, exp):
        assert len(got) == len(exp)
        for g, e in zip(got, exp):
            assert (g["Header"], g["Txname"], int(g["HitLen"]), int(g["HitPos"]), int(g["Reflen"])) == (e["Header"], e["Txname"], e["HitLen"], e["HitPos"], e["Reflen"])
            assert g["dis5"] == rf.fmt(e["dis5"]) and g["dis3"] == rf.fmt(e["dis3"]) and g["normID"] == e["normID"]
            assert g["min.circlen.adj"] == rf.fmt(e["min_circlen"]) and g["min_hitlength"] == rf.fmt(e["min_hitlength"]) and g["f_reflen"] == rf.fmt(e["f_reflen"])
            c = e["circ"]
            assert [g[k] for k in ("min.circlen", "mean.circlen", "median.circlen", "max.circlen", "txNum", "start_exon_len", "end_exon_len", "max_exon_num")] == \
                   [rf.fmt(c[k]) for k in ("min_circlen", "mean_circlen", "median_circlen", "max_circlen", "txNum", "start_exon_len", "end_exon_len", "max_exon_num")]
    same_reads(table(se_dir / "SE_read_raw.tsv"), want["read_raw"])
    same_reads(table(se_dir / "SE_read_dat.tsv"), want["read_dat"])
    got_eq = table(se_dir / "SE_eqc_dat.tsv")
    assert [(g["Transcript"], int(g["Count"]), g["normID"], g["median.circlen"], g["fragment_count_SE"]) for g in got_eq] == \
           [(e["Transcript"], e["Count"], e["normID"], rf.fmt(e["median_circlen"]), rf.fmt(e["fragment_count_SE"])) for e in want["eqc_dat"]]
    assert open(se_dir / "SE_circRNA_names.txt").read().split("\n")[:-1] == want["circRNA_names"]
    # both branches of get_circLen ran
    tx_nums = {c["circ"]["txNum"] for c in want["read_raw"]}
    assert 0 in tx_nums and max(tx_nums) >= 1

    final = tmp_path / "final.txt"
    subprocess.run([flt, "pe", "--se", str(se_dir), "--pseudo", str(ps), "--wt", str(wt), "--out", str(final)], check=True)
    hd = open(ps / "Mapped_read_1.rh").read().split("\n")[:-1]
    mp = open(ps / "mapInfo_1.txt").read().split("\n")[:-1]
    nobs = int(table(wt / "fragmentInfo.txt")[0]["numObservedFragments"])
    pe = rf.pe_filter(hd, mp, want, nobs)
    assert len(pe["final"]) > 10
    assert open(final).read().split("\n")[:-1] == rf.final_lines(pe["final"])
    assert all(r["junction_FPM"] == r["junction_fragment_count"] / nobs * 1e6 for r in pe["final"])
    # the merged files of Circall_source.sh:343-346 are read when they exist
    os.rename(ps / "Mapped_read_1.rh", ps / "merged_hitheader.txt"); os.rename(ps / "mapInfo_1.txt", ps / "merged_mapInfo.txt")
    subprocess.run([flt, "pe", "--se", str(se_dir), "--pseudo", str(ps), "--wt", str(wt), "--out", str(tmp_path / "final2.txt")], check=True)
    assert open(final).read() == open(tmp_path / "final2.txt").read()
