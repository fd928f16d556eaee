"""The drop-in executables (tools/): input parsing on the CPU, and on the GPU the files the two-process pipeline of
Circall_source.sh:273-305 produces — and the fused single pass — against the oracle's records written in the
reference's formats."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import make_reads  # noqa: E402


@pytest.fixture(scope="module")
def exe():
    from tools import build as tb
    return os.path.dirname(tb.build())


def run(cmd, ok=True):
    p = subprocess.run(cmd, capture_output=True, text=True)
    if ok:
        assert p.returncode == 0, p.stderr
    return p


def test_parse_fasta_fastq_mix(exe, tmp_path):
    a, b, c = tmp_path / "a.fa", tmp_path / "b.fq", tmp_path / "c.fa"
    a.write_text(">r1 some text\tkept\nACGT\nACGTAC\n\n>r2\nGGGG\r\n")          # multi-line FASTA, blank line, CRLF
    c.write_text(">r3\nTT\n")
    b.write_text("@r1/2 x\nACGTAA\n+\nIIIIII\n@r2/2\nGG\n+r2/2\nII\n@r3/2\nA\n+\n@\n")   # '@' as a quality character
    p = run([exe + "/circall_gpu", "parse", "-1", str(a), str(c), "-2", str(b)])
    f = p.stdout.split()
    assert f[1] == "3" and f[3] == str(10 + 4 + 2 + 6 + 2 + 1)
    # gzip files are read directly (zlib passes plain text through): no `gunzip -c` FIFO needed (Circall_source.sh:273)
    import gzip
    ag, bg = tmp_path / "a.fa.gz", tmp_path / "b.fq.gz"
    ag.write_bytes(gzip.compress(a.read_bytes()))
    bg.write_bytes(gzip.compress(b.read_bytes()[:len("@r1/2 x\nACGTAA\n+\nIIIIII\n@r2/2\nGG\n+r2/2\nII\n")]))
    p = run([exe + "/circall_gpu", "parse", "-1", str(ag), "-2", str(bg)])
    f = p.stdout.split()
    assert f[1] == "2" and f[3] == str(10 + 4 + 6 + 2)
    # unequal record counts are an error, exit status 1 (the reference's error behaviour)
    p = run([exe + "/circall_gpu", "parse", "-1", str(a), "-2", str(b)], ok=False)
    assert p.returncode == 1 and "different numbers" in p.stderr
    p = run([exe + "/Circall_wt", "-i", "nowhere", "-1", str(a), "-o", str(tmp_path / "o")], ok=False)
    assert p.returncode == 1


def test_fragment_length_prior_columns(exe):
    """Row A11: the floating-point columns the executables write (fragLengthMedian / Mean / Sd of fragmentInfo.txt, effLens /
    EffectiveLength of rawCount.txt) follow the reference's never-updated normal prior; pinned against oracle/r_fld.py, a restatement of
    getNormalFragLengthCounts + EmpiricalDistribution::buildDistribution + computeSmoothedEffectiveLengths (unpinned: no reference run)."""
    from oracle import r_fld
    lens = [31, 100, 150, 298, 999, 1000, 5000]
    out = run([exe + "/circall_gpu", "fld"] + [str(x) for x in lens]).stdout.split("\n")
    med, mean, sd = (r_fld._f32(float(out[0].split()[i])) for i in (1, 3, 5))      # printed with 9 digits: enough to name the float
    assert (med, mean, sd) == r_fld.empirical(r_fld.normal_counts())           # float values, bit for bit
    assert 190 < mean < 210 and 70 < sd < 85
    f = r_fld.correction_factors()
    for ln, line in zip(lens, out[1:]):
        assert abs(float(line.split()[2]) - r_fld.effective_length(ln, f)) <= 1e-9 * ln


def _fasta(path, names, text):
    s = bytes(text).decode()
    with open(path, "w") as f:
        for n, seq in zip(names, s.split("$")):
            f.write(">%s some description\n" % n)
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + "\n")


def _records(path):
    """FASTA file -> list of (header line, sequence)"""
    out = []
    with open(path) as f:
        lines = f.read().split("\n")
    for i in range(0, len(lines) - 1, 2):
        out.append((lines[i], lines[i + 1]))
    return out


@pytest.mark.gpu
def test_two_process_pipeline_and_fused_files(exe, tmp_path):
    import oracle
    import synth
    synth.build(); oracle.build()
    T = synth.Txome(21, 150, p_polya=0.0)      # no poly-A tails: the indexer clips them (RapMap's normalisation), the oracle text would differ
    otx, obsj = oracle.Index.from_text(T.tx_text), oracle.Index.from_text(T.bsj_text)
    txn = [T.tx_name(t) for t in range(T.n_tx)]
    bsn = [T.bsj_name(t) for t in range(T.n_bsj)]
    _fasta(tmp_path / "tx.fa", txn, T.tx_text)
    _fasta(tmp_path / "bsj.fa", bsn, T.bsj_text)
    r1, r2 = make_reads(T, 9, 6000, n_rate=0.001)
    n = r1.shape[0]
    h1 = ["read%d/1 len=100" % i for i in range(n)]
    h2 = ["read%d/2" % i for i in range(n)]
    with open(tmp_path / "r1.fq", "w") as f:                                  # mate 1 as FASTQ, mate 2 as FASTA in two files
        for i in range(n):
            s = bytes(r1[i]).decode()
            f.write("@%s\n%s\n+\n%s\n" % (h1[i], s, "I" * len(s)))
    for part, (lo, hi) in enumerate(((0, n // 3), (n // 3, n))):
        with open(tmp_path / ("r2_%d.fa" % part), "w") as f:
            for i in range(lo, hi):
                f.write(">%s\n%s\n" % (h2[i], bytes(r2[i]).decode()))
    run([exe + "/TxIndexer", "-t", str(tmp_path / "tx.fa"), "-o", str(tmp_path / "txIdx")])
    run([exe + "/TxIndexer", "-t", str(tmp_path / "bsj.fa"), "-o", str(tmp_path / "bsjIdx"), "-k", "31"])
    assert "exists" in run([exe + "/TxIndexer", "-t", str(tmp_path / "tx.fa"), "-o", str(tmp_path / "txIdx")]).stderr

    # ---- expected records from the oracle
    R = oracle.Run(otx, obsj, 2, r1, r2)
    ep, ec = R.wt_export()
    suf1, suf2 = {1: "___11", 2: "___21"}, {1: "___22", 2: "___12"}
    exp_un1 = [(">" + h1[p] + suf1[c], bytes(r1[p] if c == 1 else r2[p]).decode()) for p, c in zip(ep.tolist(), ec.tolist())]
    exp_un2 = [(">" + h1[p] + suf2[c], bytes(r2[p] if c == 1 else r1[p]).decode()) for p, c in zip(ep.tolist(), ec.tolist())]
    reflen = np.diff(np.flatnonzero(np.concatenate([[True], T.bsj_text == ord("$")]))) - 1
    reflen[0] += 1
    exp_jt = sorted("%s\t%d\t%d\t%d\t%d\t%s\t%d" % (exp_un1[rec][0], d, q, ln, hp, bsn[t], reflen[t]) for rec, d, q, ln, hp, t in R.bsj_junctions().tolist())
    cand = R.bsj_candidates().tolist()
    ow, ob = R.counters(0), R.counters(1)
    assert len(exp_un1) > 100 and len(exp_jt) > 10 and len(cand) > 5

    def check_wt(d):
        un1 = [f for f in os.listdir(d) if f.startswith("UN_read1_")]
        un2 = [f for f in os.listdir(d) if f.startswith("UN_read2_")]
        assert len(un1) == 1 and len(un2) == 1
        assert _records(os.path.join(d, un1[0])) == exp_un1            # input order (one parser, batches in order)
        assert _records(os.path.join(d, un2[0])) == exp_un2
        hdr, row = open(os.path.join(d, "fragmentInfo.txt")).read().split("\n")[:2]
        assert hdr.split("\t") == ["readlen", "fragLengthMedian", "fragLengthMean", "fragLengthSd", "numObservedFragments", "numMappedFragments", "numHits", "kmer"]
        v = row.split("\t")
        assert (int(v[0]), int(v[4]), int(v[5]), int(v[6]), int(v[7])) == (ow["readLen"], ow["numObserved"], ow["numMapped"], ow["numHits"], 31)

    def check_bsj(d):
        jt = [f for f in os.listdir(d) if f.startswith("UN_read1_")]
        assert len(jt) == 1
        assert sorted(open(os.path.join(d, jt[0])).read().split("\n")[:-1]) == exp_jt
        c1 = _records(os.path.join(d, [f for f in os.listdir(d) if f.startswith("BSJ_read1_")][0]))
        c2 = _records(os.path.join(d, [f for f in os.listdir(d) if f.startswith("BSJ_read2_")][0]))
        assert c1 == [exp_un1[r] for r in cand] and c2 == [exp_un2[r] for r in cand]
        v = open(os.path.join(d, "fragmentInfo.txt")).read().split("\n")[1].split("\t")
        assert (int(v[0]), int(v[4]), int(v[5]), int(v[6])) == (ob["readLen"], ob["numObserved"], ob["numMapped"], ob["numHits"])
        rows = [ln.split("\t") for ln in open(os.path.join(d, "rawCount.txt")).read().split("\n")[1:-1]]
        got = {}
        for name, w, cnt, _e, rl, _e2, cls in rows:
            got.setdefault(int(cls), []).append((name, float(w), int(cnt), int(rl)))
        eq = {}
        for members in got.values():
            assert all(abs(m[1] - 1.0 / len(members)) < 1e-5 for m in members) and len({m[2] for m in members}) == 1
            eq[tuple(sorted(bsn.index(m[0]) for m in members))] = members[0][2]
            assert all(m[3] == reflen[bsn.index(m[0])] for m in members)
        assert eq == R.eq_classes(1)

    # ---- the two-process pipeline of Circall_source.sh:273-305
    wt, bs = str(tmp_path / "wt"), str(tmp_path / "bs")
    run([exe + "/Circall_wt", "-i", str(tmp_path / "txIdx"), "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2_0.fa"), str(tmp_path / "r2_1.fa"),
         "-o", wt, "-p", "4", "--batch", "2500", "--wt-classes"])
    check_wt(wt)
    # Circall_wt's own rawCount.txt (opt-in: nothing downstream reads it): mate 1's transcripts, then mate 1's + mate 2's — the reference
    # does not clear the list between the mates (Circall_wt.cpp:247-252,369-465,544-640) — keyed in hit order, weights 1 / class size
    rows = [ln.split("\t") for ln in open(os.path.join(wt, "rawCount.txt")).read().split("\n")[1:-1]]
    got = {}
    for name, w, cnt, _e, rl, _e2, cls in rows:
        got.setdefault(int(cls), []).append((name, float(w), int(cnt)))
    weq = {}
    for members in got.values():
        assert all(abs(m[1] - 1.0 / len(members)) < 1e-5 for m in members) and len({m[2] for m in members}) == 1
        weq[tuple(txn.index(m[0]) for m in members)] = members[0][2]
    assert weq == R.eq_classes(0) and len(weq) > 50
    m1, m2 = str(tmp_path / "Unmapped_readM_1.fa"), str(tmp_path / "Unmapped_readM_2.fa")
    for m, pat in ((m1, "UN_read1_"), (m2, "UN_read2_")):
        with open(m, "w") as f:
            for fn in sorted(os.listdir(wt)):
                if fn.startswith(pat):
                    f.write(open(os.path.join(wt, fn)).read())
    run([exe + "/Circall_bsj", "-i", str(tmp_path / "bsjIdx"), "-1", m1, "-2", m2, "-o", bs, "-p", "4", "--batch", "700"])
    check_bsj(bs)
    # ---- the fused pass writes the same two directories
    fw, fb = str(tmp_path / "fwt"), str(tmp_path / "fbs")
    run([exe + "/circall_gpu", "fused", "-i", str(tmp_path / "txIdx"), "-b", str(tmp_path / "bsjIdx"), "-1", str(tmp_path / "r1.fq"),
         "-2", str(tmp_path / "r2_0.fa"), str(tmp_path / "r2_1.fa"), "-o", fw, "-O", fb, "--batch", "1700"])
    check_wt(fw)
    check_bsj(fb)
    # ---- several GPUs in the product (--gpus N: indices replicated, batches dealt round-robin, one all-reduce of the counters at the
    # end): the same files, byte for byte, as one GPU — with as many devices as this box has, at most 4; on one device the run still
    # goes through cg_sessions_allreduce (two sessions)
    import circall_b200 as cg
    ng = min(cg.device_count(), 4)
    mw, mb = str(tmp_path / "mwt"), str(tmp_path / "mbs")
    run([exe + "/circall_gpu", "fused", "-i", str(tmp_path / "txIdx"), "-b", str(tmp_path / "bsjIdx"), "-1", str(tmp_path / "r1.fq"),
         "-2", str(tmp_path / "r2_0.fa"), str(tmp_path / "r2_1.fa"), "-o", mw, "-O", mb, "--batch", "1700", "--gpus", str(ng)])
    check_wt(mw)
    check_bsj(mb)
    for d0, d1 in ((fw, mw), (fb, mb)):
        for fn in sorted(os.listdir(d0)):
            a, b = open(os.path.join(d0, fn), "rb").read(), open(os.path.join(d1, fn), "rb").read()
            if d0 == fb and fn.startswith("UN_read1_"):          # junction rows of a batch come out in the order the warps finished
                a, b = sorted(a.split(b"\n")), sorted(b.split(b"\n"))
            assert a == b, fn
    p = run([exe + "/circall_gpu", "fused", "-i", str(tmp_path / "txIdx"), "-b", str(tmp_path / "bsjIdx"), "-1", str(tmp_path / "r1.fq"),
             "-2", str(tmp_path / "r2_0.fa"), str(tmp_path / "r2_1.fa"), "-o", mw, "-O", mb, "--gpus", "99"], ok=False)
    assert p.returncode == 1 and "devices" in p.stderr


@pytest.mark.gpu
def test_pseudo_files(exe, tmp_path):
    """Circall_pseudo (step 5): header list + transcript-name lists of the pairs whose mates' hits intersect (or, failing
    that, orphan hits), against the oracle's collector + mergeLeftRightHitsFuzzy."""
    import oracle
    import synth
    synth.build(); oracle.build()
    T = synth.Txome(23, 120, p_polya=0.0)
    obsj = oracle.Index.from_text(T.bsj_text)
    bsn = [T.bsj_name(t) for t in range(T.n_bsj)]
    _fasta(tmp_path / "pseudo.fa", bsn, T.bsj_text)
    r1, r2 = make_reads(T, 12, 3000, circ_frac=0.6, n_rate=0.001)
    n = r1.shape[0]
    for fn, r, tagc in ((tmp_path / "b1.fa", r1, 1), (tmp_path / "b2.fa", r2, 2)):
        with open(fn, "w") as f:
            for i in range(n):
                f.write(">cand%d/%d___11\n%s\n" % (i, tagc, bytes(r[i]).decode()))
    run([exe + "/TxIndexer", "-t", str(tmp_path / "pseudo.fa"), "-o", str(tmp_path / "pIdx"), "-f"])
    out = str(tmp_path / "pseudo")
    run([exe + "/Circall_pseudo", "-i", str(tmp_path / "pIdx"), "-1", str(tmp_path / "b1.fa"), "-2", str(tmp_path / "b2.fa"), "-o", out, "-p", "2", "--batch", "1100"])
    exp_h, exp_m = [], []
    for i in range(n):
        f1, _, _, h1 = obsj.collect(r1[i].tobytes(), strict=True)
        f2, _, _, h2 = obsj.collect(r2[i].tobytes(), strict=True)
        joint, _tm = oracle.merge_fuzzy(h1, h2, f1, f2, 100, 200)
        if joint.shape[0] == 0 or joint.shape[0] > 200:
            continue
        tids = joint[:, 0].tolist()
        if joint[0, 6] != 3:
            tids = sorted(tids)
        exp_h.append(">cand%d/1___11" % i)
        exp_m.append("".join(bsn[t] + " " for t in tids))
    assert len(exp_h) > 200
    got_h = open(os.path.join(out, [f for f in os.listdir(out) if f.endswith(".rh")][0])).read().split("\n")[:-1]
    got_m = open(os.path.join(out, [f for f in os.listdir(out) if f.startswith("mapInfo_")][0])).read().split("\n")[:-1]
    assert got_h == exp_h and got_m == exp_m
