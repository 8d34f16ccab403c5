"""The envelope layer (csrc/envelope.cpp): zn_envelope_parse against the oracle's restatement of
ZNB1, bare frames, malformed envelopes, and the foreign-parser hook where an OpenZL frame-header reader plugs in.
CPU tests cover the host parser; the GPU tests decode enveloped rows through the C ABI and compare with the oracle."""
import ctypes as C
import random

import numpy as np
import pytest


@pytest.fixture(scope="module")
def codec():
    from znippy_b200 import codec
    return codec


def test_znb1_header_matches_the_oracle(codec, oracle):
    O = oracle
    for c in range(5):
        for n in (0, 1, 127, 128, 16383, 16384, 8 << 20, (1 << 32) - 1):
            payload = b"\x01\x02\x03" if c else bytes(min(n, 3))
            if c == O.PAYLOAD_RAW and n > 3:
                continue
            blob = codec.envelope_wrap(c, n, payload)
            assert blob == O.znb1_wrap(c, n, payload)
            kind, pc, off, ln, out_len = codec.envelope_parse(blob)
            assert (kind, pc, off, ln, out_len) == (codec.ENV_ZNB1,) + O.znb1_parse(blob)
    with pytest.raises(ValueError):
        codec.envelope_wrap(5, 10, b"")
    with pytest.raises(ValueError):
        codec.envelope_wrap(1, 1 << 32, b"")


def test_bare_frames_and_malformed_envelopes(codec, oracle):
    O = oracle
    data = O.real_text(50_000).tobytes()
    zf = O.libzstd().compress(data, 3)
    lf = O.liblz4().compress_frame(data)
    assert codec.envelope_parse(zf) == (codec.ENV_BARE, codec.PAYLOAD_ZSTD, 0, len(zf), len(data))
    assert codec.envelope_parse(lf) == (codec.ENV_BARE, codec.PAYLOAD_LZ4_FRAME, 0, len(lf), len(data))
    for bad in (b"", b"ZNB", b"ZNB1", b"ZNB1\x01", b"ZNB1\x07\x05abcde", b"ZNB1\x01\x80\x80\x80\x80\x80\x01x",
                b"ZNB1\x01\x80", b"ZNB1\x00\x05abc", b"OZL!\x00\x00\x00\x00\x00\x00", b"\x00" * 64):
        assert codec.envelope_parse(bad) is None, bad
        assert O.znb1_parse(bad) is None
    rnd = random.Random(3)
    good = O.znb1_wrap(O.PAYLOAD_ZSTD, len(data), zf)
    for _ in range(300):  # header fuzz: the parser and the oracle agree on accept / reject and on every field
        b = bytearray(good[:16])
        b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        b = bytes(b) + good[16:]
        got, want = codec.envelope_parse(b), O.znb1_parse(b)
        if b[:4] == b"ZNB1":
            assert (got is None) == (want is None)
            if got:
                assert got[1:] == want
        else:
            assert got is None or got[0] == codec.ENV_BARE


def test_foreign_parser_hook(codec, oracle):
    """Where an OpenZL frame-header reader plugs in: a registered parser sees blobs that are neither bare nor ZNB1; what it
    returns is bounds-checked before any kernel sees it."""
    from znippy_b200 import _native as N
    from znippy_b200.codec import _Envelope
    L = N.lib()
    PROTO = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.POINTER(_Envelope))
    calls = []

    def parse(blob, n, out):
        raw = C.string_at(blob, min(n, 12))
        calls.append(raw[:4])
        if raw[:4] != b"OZL!":
            return 1
        out.contents.codec = raw[4]
        out.contents.payload_off = 12 if raw[5] == 0 else 10_000
        out.contents.payload_len = n - 12
        out.contents.out_len = int.from_bytes(raw[8:12], "little")
        return 0

    cb = PROTO(parse)
    L.zn_envelope_register(C.cast(cb, C.c_void_p))
    try:
        blob = b"OZL!" + bytes([codec.PAYLOAD_ZSTD_MAGICLESS, 0, 0, 0]) + (77).to_bytes(4, "little") + b"payload"
        assert codec.envelope_parse(blob) == (codec.ENV_FOREIGN, codec.PAYLOAD_ZSTD_MAGICLESS, 12, 7, 77)
        assert codec.envelope_parse(b"ABCD" + blob[4:]) is None and calls[-1] == b"ABCD"
        bad = bytearray(blob); bad[5] = 1      # payload range outside the blob
        assert codec.envelope_parse(bytes(bad)) is None
        bad = bytearray(blob); bad[4] = 9      # unknown payload codec
        assert codec.envelope_parse(bytes(bad)) is None
    finally:
        L.zn_envelope_register(None)
    assert codec.envelope_parse(blob) is None


@pytest.mark.gpu
def test_enveloped_rows_decode_like_the_oracle(codec, oracle):
    O = oracle
    z, l4 = O.libzstd(), O.liblz4()
    rng = np.random.default_rng(11)
    contents = [O.real_text(200_000).tobytes(), O.gen_text(1 << 20).tobytes(), bytes(rng.integers(0, 256, 70_001, dtype=np.uint8)),
                O.real_text(9_000).tobytes(), O.gen_binary(300_000).tobytes(), b"", O.real_text(40_000).tobytes(),
                O.real_text(3 << 20).tobytes()]
    zf = [z.compress(c, 3) for c in contents]
    blobs = [O.znb1_wrap(O.PAYLOAD_ZSTD, len(contents[0]), zf[0]),
             O.znb1_wrap(O.PAYLOAD_ZSTD_MAGICLESS, len(contents[1]), z.compress(contents[1], 19)[4:]),
             O.znb1_wrap(O.PAYLOAD_RAW, len(contents[2]), contents[2]),
             O.znb1_wrap(O.PAYLOAD_LZ4_FRAME, len(contents[3]), l4.compress_frame(contents[3])),
             O.znb1_wrap(O.PAYLOAD_LZ4_BLOCK, len(contents[4]), l4.compress_block(contents[4])),
             O.znb1_wrap(O.PAYLOAD_ZSTD, 0, z.compress(b"", 3)),
             zf[6],                                                     # a bare frame behind flag 4
             O.znb1_wrap(O.PAYLOAD_ZSTD_MAGICLESS, len(contents[7]), zf[7][4:])]   # entropy-coded, device-wide pipeline
    for b, c in zip(blobs, contents):
        if b[:4] == b"ZNB1":
            assert O.envelope_decode(b) == c  # the oracle's own decode of the envelope
    # + rows that must be refused without disturbing their neighbours
    bad = [b"OZL!" + zf[0], O.znb1_wrap(O.PAYLOAD_ZSTD, len(contents[0]) + 1, zf[0]), b"ZNB1\x09\x00", O.znb1_wrap(O.PAYLOAD_ZSTD, 5, b"")]
    all_blobs = blobs + bad
    all_contents = contents + [contents[0], contents[0], b"", b"12345"]
    offs, cur = [], 3  # odd placement: payload offsets are unaligned anyway
    for b in all_blobs:
        offs.append(cur)
        cur += len(b) + 5
    buf = np.zeros(cur, np.uint8)
    for o, b in zip(offs, all_blobs):
        buf[o:o + len(b)] = np.frombuffer(b, np.uint8)
    keep = buf.copy()
    out_len = [len(c) for c in all_contents]
    out_off = np.concatenate([[0], np.cumsum([(n + 15) & ~15 for n in out_len])])[:-1]
    out = np.zeros(int(out_off[-1]) + out_len[-1] + 16, np.uint8)
    expect = b"".join(O.blake3(c) for c in all_contents)
    for with_out in (True, False):
        st, dg = codec.decode_verify_batch(buf, offs, [len(b) for b in all_blobs], [4] * len(all_blobs), out_len, expect,
                                           out if with_out else None, out_off if with_out else None)
        assert st.tolist() == [0] * 8 + [codec.S_UNSUPPORTED, codec.S_SIZE_MISMATCH, codec.S_UNSUPPORTED, codec.S_DECODE_ERROR], st
        for i, c in enumerate(contents):
            assert dg[i].tobytes() == O.blake3(c)
            if with_out:
                assert out[out_off[i]:out_off[i] + len(c)].tobytes() == c, i
    assert (buf == keep).all()  # the magic of a magicless payload is rebuilt on the DEVICE copy only


@pytest.mark.gpu
def test_envelope_archive_roundtrip_and_store_if_incompressible(codec, oracle, tmp_path):
    from znippy_b200 import archive as A
    O = oracle
    rng = np.random.default_rng(2)
    entries = [("a/text.txt", O.real_text(3_000_000).tobytes()), ("a/noise.bin", bytes(rng.integers(0, 256, 1_000_000, dtype=np.uint8))),
               ("b/pattern.txt", O.gen_text(9 << 20).tobytes()), ("b/empty", b""), ("b/app.jar", bytes(rng.integers(0, 256, 50_000, dtype=np.uint8)))]
    sc = A.compress_stream(str(tmp_path / "env.znippy"), False, envelope=True, level=3)
    for p, d in entries:
        sc.sender().send(A.ArchiveEntry(p, d))
    rep = sc.finish()
    t = A.read_znippy_index(sc.output)
    assert t.schema.metadata[b"znippy_envelope"] == b"ZNB1"
    rows = {p: (bs, c) for p, bs, c in zip(t.column("relative_path").to_pylist(), t.column("blob_size").to_pylist(),
                                         t.column("compressed").to_pylist())}
    # incompressible slice: kept `compressed` (as the reference does) but stored RAW inside the envelope: size + header
    assert rows["a/noise.bin"] == (1_000_000 + 4 + 1 + 3, True)
    assert rows["b/app.jar"] == (50_000, False)  # skip list: no envelope at all
    raw = open(sc.output, "rb").read()
    bo = dict(zip(t.column("relative_path").to_pylist(), t.column("blob_offset").to_pylist()))
    assert raw[bo["a/text.txt"]:bo["a/text.txt"] + 5] == b"ZNB1\x01" and raw[bo["a/noise.bin"]:bo["a/noise.bin"] + 5] == b"ZNB1\x00"
    for native in (True, False):
        out = tmp_path / f"x{int(native)}"
        vr = A.decompress_archive(sc.output, True, str(out), native=native)
        assert vr.corrupt_files == 0 and vr.total_files == 5 and vr.chunks == rep.chunks == 6
        for p, d in entries:
            assert (out / p).read_bytes() == d
    ar = A.ZnippyArchive.open(sc.output)
    assert ar.extract_file("a/noise.bin") == entries[1][1] and ar.extract_file("b/pattern.txt") == entries[2][1]
    # every compressed row's blob is decodable by the oracle's reading of the envelope
    for p, d in entries[:3]:
        rws = [i for i, q in enumerate(t.column("relative_path").to_pylist()) if q == p]
        got = b"".join(O.envelope_decode(raw[t.column("blob_offset")[i].as_py():t.column("blob_offset")[i].as_py() + t.column("blob_size")[i].as_py()])
                       for i in rws)
        assert got == d


@pytest.mark.gpu
def test_compress_dir_matches_compress_stream(codec, oracle, tmp_path):
    """slot_packer path: the native directory pipeline (readers filling pinned slots in parallel) writes the same rows —
    paths, chunking, digests, sizes — as compress_stream fed the same files in the same (sorted) order."""
    from znippy_b200 import archive as A
    O = oracle
    src = tmp_path / "src"
    files = {"top.txt": O.real_text(70_000).tobytes(), "d1/big.bin": O.gen_binary(20 << 20).tobytes(), "d1/empty.dat": b"",
             "d1/z.jar": O.gen_random(300_000).tobytes()}
    for i in range(400):
        files[f"d2/s{i % 7}/f{i:04d}.txt"] = O.real_text(500 + 37 * i).tobytes()
    for p, d in files.items():
        (src / p).parent.mkdir(parents=True, exist_ok=True)
        (src / p).write_bytes(d)
    rep = A.compress_dir(str(src), str(tmp_path / "dir.znippy"), level=3, slot_bytes=16 << 20, io_threads=4)
    assert rep.total_files == len(files) and rep.uncompressed_files == 1
    order = []
    def walk(d, rel):  # the native walk: sorted names, directories descended in place
        for n in sorted(p.name for p in d.iterdir()):
            q = d / n
            if q.is_dir():
                walk(q, rel + n + "/")
            else:
                order.append(rel + n)
    walk(src, "")
    sc = A.compress_stream(str(tmp_path / "stream.znippy"), False, level=3)
    for p in order:
        sc.sender().send(A.ArchiveEntry(p, files[p]))
    rep2 = sc.finish()
    assert (rep.chunks, rep.total_bytes_in, rep.total_bytes_out) == (rep2.chunks, rep2.total_bytes_in, rep2.total_bytes_out)
    t1, t2 = A.read_znippy_index(str(tmp_path / "dir.znippy")), A.read_znippy_index(sc.output)
    assert t1.equals(t2)
    vr = A.decompress_archive(str(tmp_path / "dir.znippy"), True, str(tmp_path / "out"))
    assert vr.corrupt_files == 0 and vr.total_files == len(files)
    for p, d in files.items():
        assert (tmp_path / "out" / p).read_bytes() == d
