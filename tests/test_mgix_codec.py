"""MGIX index stream codec (Index::SaveToStream / LoadFromStream, index_serialization.cpp:111-224,
279-613; PostingList::Serialize / Deserialize, posting_list.cpp:973-1102). The codec is host-only, so everything
here runs without a GPU:
  * streams written by mgx_mgix_encode load into the REFERENCE's own Index::LoadFromStream (oracle/_ref) and answer
    like the index they came from;
  * streams written by the reference's Index::SaveToStream decode to the same CSR;
  * the byte layouts the reference's tests pin (tests/index/posting_list_serialization_test.cpp:30-50, 111-130) and
    the published Roaring interchange layouts (array / bitset / run containers);
  * every rejection path of LoadFromData / Deserialize;
  * a committed stream recorded from the reference (tests/golden/ref_mgix.json, oracle/gen_golden.py)."""
import base64
import json
import os
import random
import struct
import zlib

import numpy as np
import pytest

from test_oracle_bulk import _docs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_mgix.json")


def csr_of(oi):
    terms, offs, posts = oi.export()
    return [bytes(t) for t in terms], np.asarray(offs, dtype=np.uint64), np.asarray(posts, dtype=np.uint32)


def make_index(lib, cfg, docs, ids):
    idx = lib.index(*cfg)
    idx.add_texts(ids, docs)
    return idx


def big_corpus(seed):
    """Lists below and above 4096 entries, bitset containers (> 4096 ids inside one 64Ki chunk), several chunks."""
    rnd = random.Random(seed)
    docs = _docs(rnd, 12000)
    for i in range(0, len(docs), 2):
        docs[i] += b" zq"          # 6000 postings for "zq" / " z"
    ids = np.concatenate([np.arange(1, 9001, dtype=np.uint32),                     # dense run inside chunk 0
                          np.arange(70000, 70000 + 3 * 3000, 3, dtype=np.uint32)])  # sparse ids in chunks 1..
    return docs, ids


@pytest.mark.parametrize("cfg", [(2, 0, True), (2, 1, True), (1, 1, True), (3, 2, False)])
def test_encoded_stream_loads_into_the_reference(mgx, oracle, reflib, cfg):
    docs, ids = big_corpus(hash(cfg) & 0xFFFF)
    oi = make_index(oracle, cfg, docs, ids)
    terms, offs, posts = csr_of(oi)
    sizes = np.diff(offs.astype(np.int64))
    assert sizes.max() > 4096 and sizes.min() >= 1
    eff_kanji = cfg[1] if cfg[1] > 0 else cfg[0]
    for roaring_min_len in (0.0, 0.18 * 500):  # as built / as after Index::Optimize(500): more lists as Roaring
        stream = mgx.mgix_encode(terms, offs, posts, cfg[0], eff_kanji, cfg[2], roaring_min_len=roaring_min_len)
        ri = reflib.index(*cfg)
        assert ri.load_stream(stream) == 0
        assert ri.term_count() == len(terms) and ri.total_postings() == posts.size
        rnd = random.Random(1)
        for t in rnd.sample(range(len(terms)), min(300, len(terms))) + [int(np.argmax(sizes))]:
            assert np.array_equal(ri.postings(terms[t]), posts[int(offs[t]):int(offs[t + 1])]), terms[t]
        for _ in range(30):
            pick = [terms[rnd.randrange(len(terms))] for _ in range(2)]
            assert np.array_equal(ri.search_and(pick), oi.search_and(pick))
            assert np.array_equal(ri.search_or(pick), oi.search_or(pick))
        # and back: what the reference saves from that loaded index decodes to the same CSR
        meta, t2, o2, p2 = mgx.mgix_decode(ri.save_stream())
        assert t2 == terms and np.array_equal(o2, offs) and np.array_equal(p2, posts)
        assert (meta["ngram_size"], meta["kanji_ngram_size"], bool(meta["cross_boundary"])) == (cfg[0], eff_kanji, cfg[2])
    # a configuration mismatch is the reference's to refuse (kStorageVersionMismatch), the stream itself is valid
    other = reflib.index(cfg[0] + 1, cfg[1], cfg[2])
    assert other.load_stream(stream) != 0


def test_reference_stream_decodes_to_the_same_index(mgx, oracle, reflib):
    cfg = (2, 1, True)
    docs, ids = big_corpus(7)
    ri = make_index(reflib, cfg, docs, ids)
    oi = make_index(oracle, cfg, docs, ids)
    meta, terms, offs, posts = mgx.mgix_decode(ri.save_stream())
    t0, o0, p0 = csr_of(oi)
    assert terms == t0 and np.array_equal(offs, o0) and np.array_equal(posts, p0)
    assert meta["version"] == 4 and meta["normalize_width"] == "keep" and meta["n_postings"] == p0.size
    # re-encoding what was decoded is loadable again and byte-stable under a second round trip
    again = mgx.mgix_encode(terms, offs, posts, 2, 1, True)
    assert mgx.mgix_decode(again)[1:] [0] == terms
    assert mgx.mgix_encode(*mgx.mgix_decode(again)[1:], 2, 1, True) == again
    empty = reflib.index(*cfg)
    meta, terms, offs, posts = mgx.mgix_decode(empty.save_stream())
    assert meta["n_terms"] == 0 and terms == [] and posts.size == 0
    assert empty.load_stream(mgx.mgix_encode([], [0], [], 2, 1, True)) == 0


def stream_of(records, version=4, ngram=2, kanji=2, cross=1, crc=True):
    s = b"MGIX" + struct.pack("<II", version, ngram)
    if version >= 3:
        s += struct.pack("<IB", kanji, cross)
    if version >= 4:
        s += struct.pack("<BI", 1, 4) + b"keep" + struct.pack("<B", 1)
    s += struct.pack("<Q", len(records))
    for term, body in records:
        s += struct.pack("<I", len(term)) + term + struct.pack("<Q", len(body)) + body
    if version >= 2 and crc:
        s += struct.pack("<I", zlib.crc32(s))
    return s


def delta_body(ids):
    gaps = [ids[0]] + [b - a for a, b in zip(ids, ids[1:])] if ids else []
    return struct.pack("<BI", 0, len(gaps)) + b"".join(struct.pack("<I", g) for g in gaps)


def test_byte_layouts_pinned_by_the_reference_tests_and_the_roaring_format(mgx):
    # PostingListSerializationTest.LittleEndianByteOrder (posting_list_serialization_test.cpp:30-50): {1, 2}
    s = mgx.mgix_encode([b"ab"], [0, 2], [1, 2], 2, 2, True)
    body = struct.pack("<BI", 0, 2) + struct.pack("<II", 1, 1)
    assert s == stream_of([(b"ab", body)])
    assert s[-4:] == struct.pack("<I", zlib.crc32(s[:-4]))  # utils/crc32.h is zlib's CRC-32
    # RejectsInternallyInvalidRoaringBitmap (:111-130): 200 values i*3 as Roaring; the first array value sits at byte
    # 5 + 16 of the body, and making it 3 (a duplicate of the second value) must be refused
    ids = [i * 3 for i in range(200)]
    s = mgx.mgix_encode([b"ab"], [0, 200], ids, 2, 2, True, roaring_min_len=1.0)
    roaring = struct.pack("<II", 12346, 1) + struct.pack("<HH", 0, 199) + struct.pack("<I", 16) + \
        b"".join(struct.pack("<H", v) for v in ids)
    body = struct.pack("<BI", 1, len(roaring)) + roaring
    assert s == stream_of([(b"ab", body)])
    assert np.array_equal(mgx.mgix_decode(s)[3], ids)
    bad = bytearray(body)
    bad[5 + 16] = 3
    with pytest.raises(mgx.MgxError, match="kIndexDeserializationFailed"):
        mgx.mgix_decode(stream_of([(b"ab", bytes(bad))]))
    # bitset container (cardinality > 4096) + array container in a second chunk, offsets 24 and 24 + 8192
    ids = list(range(10, 5010)) + [65536 + 7, 65536 + 9]
    s = mgx.mgix_encode([b"xy"], [0, len(ids)], ids, 2, 2, True)
    words = [0] * 1024
    for v in range(10, 5010):
        words[v >> 6] |= 1 << (v & 63)
    roaring = struct.pack("<II", 12346, 2) + struct.pack("<HHHH", 0, 4999, 1, 1) + struct.pack("<II", 24, 24 + 8192) + \
        b"".join(struct.pack("<Q", w) for w in words) + struct.pack("<HH", 7, 9)
    assert s == stream_of([(b"xy", struct.pack("<BI", 1, len(roaring)) + roaring)])
    assert np.array_equal(mgx.mgix_decode(s)[3], ids)
    # run containers, as CRoaring writes them after run_optimize (index Optimize, posting_list.cpp:811):
    # cookie 12347 | (n-1) << 16, run flags, no offset header below 4 containers
    runs = struct.pack("<I", 12347 | (1 << 16)) + b"\x01" + struct.pack("<HHHH", 0, 6, 2, 1) + \
        struct.pack("<HHHHH", 2, 10, 4, 100, 1) + struct.pack("<HH", 5, 6)
    s = stream_of([(b"r", struct.pack("<BI", 1, len(runs)) + runs)])
    assert mgx.mgix_decode(s)[3].tolist() == [10, 11, 12, 13, 14, 100, 101, 2 * 65536 + 5, 2 * 65536 + 6]
    # ... and with the offset header from 4 containers on
    n = 5
    hdr = struct.pack("<I", 12347 | ((n - 1) << 16)) + b"\x10" + b"".join(struct.pack("<HH", k, 0) for k in range(n))
    hdr += b"".join(struct.pack("<I", 0) for _ in range(n))
    data = b"".join(struct.pack("<H", 1) for _ in range(4)) + struct.pack("<HHH", 1, 3, 0)
    s = stream_of([(b"r", struct.pack("<BI", 1, len(hdr + data)) + hdr + data)])
    assert mgx.mgix_decode(s)[3].tolist() == [1, 65537, 131073, 196609, 4 * 65536 + 3]


def test_older_versions_and_every_rejection_path(mgx):
    recs = [(b"bc", delta_body([5, 9, 10])), (b"ab", delta_body([1]))]
    for version in (1, 2, 3, 4):
        meta, terms, offs, posts = mgx.mgix_decode(stream_of(recs, version=version, kanji=1, cross=0))
        assert meta["version"] == version and terms == [b"ab", b"bc"]  # ascending term order
        assert offs.tolist() == [0, 1, 4] and posts.tolist() == [1, 5, 9, 10]
        assert meta["kanji_ngram_size"] == (1 if version >= 3 else 2)
    good = stream_of(recs)

    def rejected(data, code):
        with pytest.raises(mgx.MgxError, match=code) as e:
            mgx.mgix_decode(data)
        assert e.value.code == -6

    rejected(b"MGIX" + b"\0" * 8, "kStorageInvalidFormat")                       # too short
    rejected(b"XGIX" + good[4:], "kStorageInvalidFormat")                        # magic
    rejected(stream_of(recs, version=9), "kStorageVersionMismatch")
    rejected(good[:-1] + bytes([good[-1] ^ 1]), "kStorageCRCMismatch")
    flipped = bytearray(good)
    flipped[40] ^= 0x40
    rejected(bytes(flipped), "kStorageCRCMismatch")
    rejected(stream_of(recs, version=1)[:-3], "kStorageCorrupted")               # truncated body (v1 has no CRC)
    rejected(stream_of([(b"x" * 10001, delta_body([1]))]), "kStorageCorrupted")  # term length guard
    rejected(stream_of([(b"ab", delta_body([1]))])[:35] + b"", "kStorage")       # cut inside the header
    rejected(stream_of([(b"ab", struct.pack("<BI", 0, 3) + struct.pack("<III", 5, 0, 1))]),
             "kIndexDeserializationFailed")                                       # zero gap (:132-148)
    rejected(stream_of([(b"ab", struct.pack("<BI", 0, 2) + struct.pack("<II", 0xFFFFFFFF, 1))]),
             "kIndexDeserializationFailed")                                       # cumulative overflow
    rejected(stream_of([(b"ab", struct.pack("<BI", 2, 0))]), "kIndexDeserializationFailed")  # unknown strategy
    rejected(stream_of([(b"ab", struct.pack("<BI", 0, 7) + b"\0" * 8)]), "kIndexDeserializationFailed")
    rejected(stream_of([(b"ab", struct.pack("<BI", 1, 4) + struct.pack("<I", 999))]), "kIndexDeserializationFailed")
    # encode refuses what could not be a posting list
    with pytest.raises(mgx.MgxError):
        mgx.mgix_encode([b"ab"], [0, 2], [2, 2], 2, 2, True)


def test_stream_recorded_from_the_reference(mgx):
    """tests/golden/ref_mgix.json: Index::SaveToStream output of the reference's own sources (oracle/gen_golden.py),
    so the decode side stays pinned where /root/reference does not exist."""
    g = json.load(open(GOLDEN))
    for case in g["cases"]:
        meta, terms, offs, posts = mgx.mgix_decode(base64.b64decode(case["stream_b64"]))
        assert [t.hex() for t in terms] == case["terms_hex"]
        assert offs.tolist() == case["posting_offsets"]
        assert zlib.crc32(posts.tobytes()) == case["postings_crc32"] and posts.size == case["n_postings"]
        assert [meta["ngram_size"], meta["kanji_ngram_size"], meta["cross_boundary"]] == case["config"]


@pytest.mark.parametrize("threads", ["1", "6"])
def test_serial_and_threaded_codec_write_the_same_bytes(mgx, monkeypatch, threads):
    """The codec is two-pass and multi-threaded for whole shards (ranges balanced by postings, CRC-32 computed in
    pieces and combined); MGX_MGIX_THREADS pins the worker count. Both paths must produce the same stream, a CRC that
    zlib agrees with, and decode back to the input; a corrupted list is reported for the FIRST bad record."""
    rng = np.random.default_rng(5)
    n_terms = 4000
    sizes = np.minimum((rng.pareto(1.0, n_terms) + 1).astype(np.int64), 40000)
    sizes[:3] = [30000, 5000, 4097]
    offs = np.zeros(n_terms + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(sizes)
    posts = np.concatenate([np.cumsum(rng.integers(1, 9, int(n), dtype=np.uint32), dtype=np.uint32) for n in sizes])
    terms = [b"%05d" % i for i in range(n_terms)]
    monkeypatch.setenv("MGX_MGIX_THREADS", "1")
    serial = mgx.mgix_encode(terms, offs, posts, 2, 2, True)
    monkeypatch.setenv("MGX_MGIX_THREADS", threads)
    stream = mgx.mgix_encode(terms, offs, posts, 2, 2, True)
    assert stream == serial
    assert stream[-4:] == struct.pack("<I", zlib.crc32(stream[:-4]))
    meta, t2, o2, p2 = mgx.mgix_decode(stream)
    assert t2 == terms and np.array_equal(o2, offs) and np.array_equal(p2, posts)
    # two corrupted delta lists (a zero gap each): the error names the first one in stream order
    bad = bytearray(stream)
    pos = {}
    cursor = 27 + 8  # v4 header with "keep" (27 bytes) + term count
    for i in range(n_terms):
        tl = struct.unpack_from("<I", bad, cursor)[0]
        body_len = struct.unpack_from("<Q", bad, cursor + 4 + tl)[0]
        pos[i] = cursor + 4 + tl + 8
        cursor = pos[i] + body_len
    victims = [i for i in range(10, n_terms) if 3 <= sizes[i] <= 4096][:400:399]
    for i in victims:
        struct.pack_into("<I", bad, pos[i] + 5 + 4, 0)  # second word of the delta body = a zero gap
    struct.pack_into("<I", bad, len(bad) - 4, zlib.crc32(bytes(bad[:-4])))
    with pytest.raises(mgx.MgxError, match="term %05d" % victims[0]):
        mgx.mgix_decode(bytes(bad))
