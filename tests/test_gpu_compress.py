"""Write-side parity: GPU-compressed blobs must be decodable by the reference decoders
(stock libzstd 1.5.5 / liblz4 1.9.4 and the oracle restatements) to identical bytes, the digest must be the blake3 of
the ORIGINAL bytes (stream_packer.rs:219), and the GPU decoder must round-trip its own output.  Also the reference's
own codec unit tests (codec.rs:84-123) through the mirrored API."""
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    from znippy_b200 import codec as c
    c.default_ctx()
    return c


def _cases(O):
    return {
        "empty": np.zeros(0, np.uint8), "one": np.array([7], np.uint8), "t11": O.gen_text(11), "t13": O.gen_text(13),
        "text10k": O.gen_text(10240), "text3m": O.gen_text(3 << 20), "bin1m+17": O.gen_binary((1 << 20) + 17),
        "real": O.real_text(1_500_000), "rand": O.gen_random(300_000), "zeros": np.zeros(500_000, np.uint8),
        "small": O.gen_small_alphabet(200_000), "rle": O.gen_rle_literals(), "inc": O.gen_incompressible(100_000, 3),
        # more than 128 literal symbols: the Huffman tree goes out as FSE-compressed weights
        "skew256": _skew256(400_000), "exe": np.fromfile(sys.executable, np.uint8)[:900_000]}


def _skew256(n, seed=5):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, 257) ** 1.1
    return rng.choice(256, n, p=p / p.sum()).astype(np.uint8)


@pytest.mark.parametrize("codec_name", ["zstd", "lz4"])
def test_frames_decode_with_reference_decoders(codec, oracle, codec_name):
    O = oracle
    cid = codec.CODEC_ZSTD if codec_name == "zstd" else codec.CODEC_LZ4
    cases = _cases(O)
    names = list(cases)
    datas = [cases[k].tobytes() for k in names]
    offs, cur = [], 5  # misaligned source offsets
    for d in datas:
        offs.append(cur)
        cur += len(d) + 3
    src = np.zeros(cur + 1, np.uint8)
    for o, d in zip(offs, datas):
        src[o:o + len(d)] = np.frombuffer(d, np.uint8)
    blobs, dg, st = codec.compress_batch(src, offs, [len(d) for d in datas], level=3, codec=cid)
    assert not st.any()
    for name, d, b, g in zip(names, datas, blobs, dg):
        assert g.tobytes() == O.blake3(d), name
        assert len(b) <= codec.compress_bound(len(d), cid), name
        if codec_name == "zstd":
            assert O.libzstd().decompress(b, len(d)) == d, name
            rc, o = O.zstd_decompress(b, len(d))
            assert rc == 0 and o == d, name
        else:
            assert O.liblz4().decompress_frame(b, len(d)) == d, name
            rc, o = O.lz4_frame_decompress(b, len(d))
            assert rc == 0 and o == d, name
        assert codec.frame_content_size(b) == len(d)
    # and back through the GPU decoder, as one batch
    buf = np.frombuffer(b"".join(blobs), np.uint8)
    boffs = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])[:-1]
    olen = [len(d) for d in datas]
    ooff = np.concatenate([[0], np.cumsum(olen)])[:-1]
    out = np.zeros(sum(olen) + 1, np.uint8)
    st, dg2 = codec.decode_verify_batch(buf, boffs, [len(b) for b in blobs], [1] * len(blobs), olen, dg.tobytes(), out, ooff)
    assert not st.any()
    for i, d in enumerate(datas):
        assert out[ooff[i]:ooff[i] + len(d)].tobytes() == d


def test_ratio_against_reference_libraries(codec, oracle):
    """North-star criterion: "ratio within a stated percentage of the reference AT THE SAME LEVEL".  The reference's codec
    is OpenZL over zstd at `compression_level` (codec.rs:16-28); libzstd 1.5.5 at the same level stands in for it.  The
    stated gaps (size of our frame over libzstd's, same level):
        pattern corpora (README shapes)   <= 1 % at levels 1 and 3, <= 10 % at level 19
        real text (python sources, 2 MB)  <= 12 % at level 1, <= 22 % at level 3, <= 55 % at level 19
    The real-text gap GROWS with the level and that is the honest state of the compressor: its match finder is greedy /
    warp-lazy inside a 20-62 KiB window per 128 KiB block, libzstd's level 3 searches a 2 MiB window
    with hash chains and level 19 adds an optimal parser over 8 MiB.  Measured: 9.8 / 19.4 / 50.9 %."""
    O = oracle
    z, l = O.libzstd(), O.liblz4()
    gap = {"text": {1: 1.01, 3: 1.01, 19: 1.10}, "binary": {1: 1.01, 3: 1.01, 19: 1.10}, "real": {1: 1.12, 3: 1.22, 19: 1.55}}
    for name, d in [("text", O.gen_text(8 << 20)), ("binary", O.gen_binary(8 << 20)), ("real", O.real_text(2_000_000))]:
        sizes = []
        for level in (1, 3, 19):  # the three efforts of the zstd match finder
            ref = len(z.compress(d, level))  # the SAME level
            b = codec.CompressCtx(level).compress(d)
            assert z.decompress(b, len(d)) == d.tobytes()
            assert len(b) <= ref * gap[name][level] + 16, (name, level, len(b), ref, round(len(b) / ref, 3))
            sizes.append(len(b))
        if name == "real":
            assert sizes[2] <= sizes[0], sizes  # more window, no worse
        ours4 = len(codec.CompressCtx(3, codec.CODEC_LZ4).compress(d))
        ref4 = len(l.compress_frame(d))
        assert ours4 <= ref4 * 1.10 + 64, (name, ours4, ref4)


def test_codec_rs_roundtrip_tests(codec):
    """codec.rs:84-123 through the mirrored API: test_roundtrip, test_multi_compress_same_ctx, test_parallel_contexts."""
    ctx = codec.CompressCtx.new(3)
    inp = (b"Hello world! This is a test of compression roundtrip. Repeated data helps compression. "
           b"Repeated data helps compression. Repeated data helps compression.")
    comp = ctx.compress(inp)
    assert codec.decompress_frame(comp) == inp
    for i in range(10):
        d = bytes((x + i) % 251 for x in range(4096))
        out = bytearray()
        n = ctx.compress_into(d, out)
        assert n == len(out)
        assert codec.decompress_frame(bytes(out)) == d, i
    errs = []

    def worker(t):
        try:
            from znippy_b200 import Ctx
            c = codec.CompressCtx(3, ctx=Ctx(0))  # one context per thread, as codec.rs:13 requires
            for i in range(5):
                d = bytes((x + i + t * 100) % 251 for x in range(8192))
                b = c.compress(d)
                out = np.zeros(8192, np.uint8)
                st, _ = codec.decode_verify_batch(np.frombuffer(b, np.uint8), [0], [len(b)], [1], [8192], None, out, [0], c.ctx)
                assert st[0] == 0 and out.tobytes() == d
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_full_size_sweep_500mib_binary_property(codec, oracle):
    """BASELINE configs[3] at full size: 500 MiB binary pattern in 8 MiB slices, zstd and LZ4; every blob is
    round-tripped through the stock library decoder and digests obey the slice-periodicity property."""
    O = oracle
    total = 500 << 20
    sl = 8 << 20
    src = O.gen_binary(total)
    offs = list(range(0, total, sl))
    lens = [min(sl, total - o) for o in offs]
    for cid, dec in [(codec.CODEC_ZSTD, lambda b, n: O.libzstd().decompress(b, n)),
                     (codec.CODEC_LZ4, lambda b, n: O.liblz4().decompress_frame(b, n))]:
        blobs, dg, st = codec.compress_batch(src, offs, lens, level=3, codec=cid)
        assert not st.any()
        for i in (0, 1, 31, len(offs) - 1):
            assert dec(blobs[i], lens[i]) == src[offs[i]:offs[i] + lens[i]].tobytes()
        # slices starting at the same phase of the 251-byte period hold identical bytes -> identical digests / blobs
        phase = {}
        for i, o in enumerate(offs):
            key = (o % 251, lens[i])
            if key in phase:
                assert dg[i].tobytes() == dg[phase[key]].tobytes() and blobs[i] == blobs[phase[key]]
            phase[key] = i
        assert dg[0].tobytes().hex() == "1adedad9735f565ac6e22dab203db63b960c27098f2c0f0fda9adf9238d4c0c9"
        ratio = total / sum(len(b) for b in blobs)
        assert ratio > (1000 if cid == codec.CODEC_ZSTD else 100), ratio
