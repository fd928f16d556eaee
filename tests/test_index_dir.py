"""Index directories (row b: "index formats stay unchanged"): the file set RapMap's indexer writes behind TxIndexer, as SURVEY.md B7
recollects it (UNPINNED: the reference ships no sample index), written here by an independent Python writer, against the library's
reader (host side without a GPU; suffix array / table rebuilt on the device with one), and the library's writer against an independent
Python reader."""
import json
import os
import struct

import numpy as np
import pytest

import circall_b200 as cg


def write_rapmap_dir(d, text, names, k=31, big=False, header=True):
    """cereal binary: u64 element count then the raw little-endian payload; a string = u64 length + bytes."""
    os.makedirs(d, exist_ok=True)
    text = bytes(text)
    ends = [i for i, c in enumerate(text) if c == ord("$")]
    starts = [0] + [e + 1 for e in ends[:-1]]
    assert len(starts) == len(names)
    with open(os.path.join(d, "txpInfo.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(names)))
        for n in names:
            f.write(struct.pack("<Q", len(n)) + n.encode())
        f.write(struct.pack("<Q", len(starts)))
        f.write(np.asarray(starts, np.int64 if big else np.int32).tobytes())
        f.write(struct.pack("<Q", len(text)) + text)
        f.write(struct.pack("<Q", len(starts)))
        f.write(np.asarray([e - s for s, e in zip(starts, ends)], np.uint32).tobytes())
    if header:
        json.dump({"value0": {"IndexType": 1, "IndexVersion": "q3", "UsesKmers": True, "KmerLen": k, "BigSA": big, "PerfectHash": False}},
                  open(os.path.join(d, "header.json"), "w"), indent=4)
    json.dump({"indexVersion": 2, "indexType": 1, "kmerLength": k}, open(os.path.join(d, "versionInfo.json"), "w"), indent=4)
    open(os.path.join(d, "hash.bin"), "wb").write(b"\x00" * 64)          # a sparsehash dump: never read


def read_rapmap_dir(d):
    """-> names, offsets, text, lens, sa, '$' bit positions, k  (independent of the library)"""
    b = open(os.path.join(d, "txpInfo.bin"), "rb").read()
    p = 0
    (n,) = struct.unpack_from("<Q", b, p); p += 8
    names = []
    for _ in range(n):
        (ln,) = struct.unpack_from("<Q", b, p); p += 8
        names.append(b[p:p + ln].decode()); p += ln
    (m,) = struct.unpack_from("<Q", b, p); p += 8
    off = np.frombuffer(b, np.int32, m, p); p += 4 * m
    (tl,) = struct.unpack_from("<Q", b, p); p += 8
    text = b[p:p + tl]; p += tl
    (m2,) = struct.unpack_from("<Q", b, p); p += 8
    lens = np.frombuffer(b, np.uint32, m2, p); p += 4 * m2
    assert p == len(b) and m == n == m2
    s = open(os.path.join(d, "sa.bin"), "rb").read()
    (ns,) = struct.unpack_from("<Q", s, 0)
    sa = np.frombuffer(s, np.int32, ns, 8)
    assert len(s) == 8 + 4 * ns
    r = open(os.path.join(d, "rsd.bin"), "rb").read()
    (nb,) = struct.unpack_from("<Q", r, 0)
    bits = np.unpackbits(np.frombuffer(r, np.uint8, (nb + 7) // 8, 8), bitorder="little")[:nb]
    hj = json.load(open(os.path.join(d, "header.json")))["value0"]
    vj = json.load(open(os.path.join(d, "versionInfo.json")))
    assert hj["KmerLen"] == vj["kmerLength"] and hj["BigSA"] is False
    return names, off, text, lens, sa, np.flatnonzero(bits), hj["KmerLen"]


def tiny_text(seed=5, n_tx=7):
    rng = np.random.default_rng(seed)
    parts, names = [], []
    for t in range(n_tx):
        parts.append(bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(60, 400))).astype(np.uint8)) + b"$")
        names.append("ENST%05d.%d" % (t, t % 3))
    return b"".join(parts), names


@pytest.mark.parametrize("big", [False, True])
def test_reference_layout_is_read_on_the_host(tmp_path, big):
    text, names = tiny_text()
    d = str(tmp_path / "idx")
    write_rapmap_dir(d, text, names, k=25, big=big)
    info = cg.Index.dir_info(d)
    assert info == dict(text_len=len(text), n_tx=len(names), k=25, n_kmers=0, hash_capacity=0, layout=1)
    os.remove(os.path.join(d, "header.json"))                       # k from versionInfo.json, the offset width from the file itself
    assert cg.Index.dir_info(d)["k"] == 25 and cg.Index.dir_info(d)["text_len"] == len(text)


def test_malformed_directories_are_refused(tmp_path):
    text, names = tiny_text()
    d = str(tmp_path / "idx")
    with pytest.raises(cg.CircallError) as e:
        cg.Index.dir_info(d)
    assert e.value.code == -3
    write_rapmap_dir(d, text, names)
    p = os.path.join(d, "txpInfo.bin")
    good = open(p, "rb").read()
    for cut in (4, 40, len(good) // 2, len(good) - 3):
        open(p, "wb").write(good[:cut])
        with pytest.raises(cg.CircallError) as e:
            cg.Index.dir_info(d)
        assert e.value.code == -3
    open(p, "wb").write(struct.pack("<Q", 1 << 60) + good[8:])      # absurd transcript count
    with pytest.raises(cg.CircallError):
        cg.Index.dir_info(d)
    bad = bytearray(good)
    at = 8 + sum(8 + len(n) for n in names) + 8 + 4                 # second offset
    bad[at:at + 4] = struct.pack("<i", 0)                           # offsets must increase
    open(p, "wb").write(bytes(bad))
    with pytest.raises(cg.CircallError):
        cg.Index.dir_info(d)


@pytest.mark.gpu
def test_open_a_directory_written_by_the_reference_layout_writer(tmp_path, small_world):
    T, otx, _ = small_world
    names = ["tx%d|gene%d some description" % (i, i // 3) for i in range(T.n_tx)]
    d = str(tmp_path / "ref_idx")
    write_rapmap_dir(d, T.tx_text.tobytes(), names)
    assert cg.Index.dir_info(d)["layout"] == 1
    g = cg.Index.open(d)                                             # no cache: text from txpInfo.bin, SA and table rebuilt on the device
    assert g.n_tx == T.n_tx and g.text_len == T.tx_text.size and g.k == 31 and g.n_kmers == otx.n_kmers
    assert np.array_equal(g.suffix_array(), otx.sa)
    assert [g.tx_name(t) for t in (0, 1, T.n_tx - 1)] == [names[0], names[1], names[-1]]
    # saving adds this library's cache and rewrites the reference's files; an independent reader finds what went in
    d2 = str(tmp_path / "saved")
    g.save(d2)
    n2, off, text, lens, sa, dollars, k = read_rapmap_dir(d2)
    assert n2 == names and text == T.tx_text.tobytes() and k == 31
    assert np.array_equal(sa, otx.sa) and np.array_equal(dollars, np.flatnonzero(T.tx_text == ord("$")))
    assert np.array_equal(off, [g.tx_offset(t) for t in range(T.n_tx)]) and np.array_equal(lens, [g.tx_len(t) for t in range(T.n_tx)])
    assert cg.Index.dir_info(d2)["layout"] == 2
    h = cg.Index.open(d2)                                            # with the cache: nothing rebuilt but the derived arrays
    assert h.n_kmers == g.n_kmers and h.hash_capacity == g.hash_capacity and np.array_equal(h.suffix_array(), otx.sa)
    # both handles map reads alike
    r1, r2 = T.reads(9, 3000, circ_frac=0.05)
    outs = []
    for ix in (g, h):
        S = cg.Session(ix, None, stage=1, keep_mate_detail=True)
        B = S.run(r1, r2)
        outs.append((B.exp_pair.tobytes(), B.exp_code.tobytes(), B.mate_flags.tobytes(), B.mate_nhits.tobytes()))
        S.close()
    assert outs[0] == outs[1]
    # a stale cache (another text's) is ignored, not trusted
    write_rapmap_dir(d2, T.bsj_text.tobytes(), ["b%d" % i for i in range(T.n_bsj)])
    assert cg.Index.dir_info(d2)["layout"] == 1
