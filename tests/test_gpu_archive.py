"""The reference's integration tests (tests/tests/integration_test.rs, repro_crate.rs) restated against the GPU-backed
host loops of znippy_b200.archive: compress_stream -> .znippy v0.7 -> decompress_archive / verify / extract_file."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from znippy_b200 import archive, codec
    codec.default_ctx()
    return archive


def _pack(A, tmp_path, entries, no_skip=False, **kw):
    sc = A.compress_stream(str(tmp_path / "out.znippy"), no_skip, **kw)
    for path, data in entries:
        sc.sender().send(A.ArchiveEntry(path, data))
    rep = sc.finish()
    return sc.output, rep


def test_single_small_file_roundtrip(A, tmp_path):  # integration_test.rs:38-67
    content = b"Hello, znippy! This is a test of streaming compression.\n" * 4
    path, rep = _pack(A, tmp_path, [("hello.txt", content)])
    assert rep.total_files == 1 and rep.chunks == 1
    out = tmp_path / "x"
    vr = A.decompress_archive(path, True, str(out))
    assert vr.total_files == 1 and vr.verified_files == 1 and vr.corrupt_files == 0 and vr.chunks == 1
    assert (out / "hello.txt").read_bytes() == content


def test_multiple_files_and_empty_file(A, tmp_path):  # :69-131
    entries = [(f"dir/file_{i}.txt", (f"content of file {i}\n" * (i + 1)).encode()) for i in range(10)] + [("empty.txt", b"")]
    path, rep = _pack(A, tmp_path, entries)
    t = A.read_znippy_index(path)
    assert t.num_rows == 11  # empty file -> exactly one row (stream_packer.rs:169-183)
    vr = A.decompress_archive(path, True, str(tmp_path / "x"))
    assert vr.total_files == 11 and vr.corrupt_files == 0
    for p, d in entries:
        assert (tmp_path / "x" / p).read_bytes() == d


def test_large_file_is_cut_at_8mib_and_reassembles(A, tmp_path):  # :133-158 and :616-642
    data = (np.arange(12 * 1024 * 1024, dtype=np.uint32) % 251).astype(np.uint8).tobytes()
    path, rep = _pack(A, tmp_path, [("big.bin", data)])
    t = A.read_znippy_index(path)
    assert t.num_rows >= 2 and t.column("chunk_seq").to_pylist() == [0, 1]
    assert t.column("fdata_offset").to_pylist() == [0, 8 << 20]
    ar = A.ZnippyArchive.open(path)
    assert ar.contains("big.bin") and ar.file_size("big.bin") == len(data)
    assert ar.extract_file("big.bin") == data
    with pytest.raises(KeyError):
        ar.extract_file("nope")
    vr = A.decompress_archive(path, True, str(tmp_path / "x"))
    assert vr.chunks == 2 and (tmp_path / "x" / "big.bin").read_bytes() == data


def test_skip_list_and_no_skip(A, tmp_path):  # :160-210
    png = bytes(np.random.default_rng(0).integers(0, 256, 5000, dtype=np.uint8))
    path, rep = _pack(A, tmp_path, [("image.png", png), ("a.txt", b"abc" * 100)])
    t = A.read_znippy_index(path)
    rows = dict(zip(t.column("relative_path").to_pylist(), t.column("compressed").to_pylist()))
    assert rows == {"image.png": False, "a.txt": True}
    sizes = dict(zip(t.column("relative_path").to_pylist(), t.column("blob_size").to_pylist()))
    assert sizes["image.png"] == 5000
    (tmp_path / "ns").mkdir()
    path2, _ = _pack(A, tmp_path / "ns", [("image.png", png)], no_skip=True)
    assert A.read_znippy_index(path2).column("compressed").to_pylist() == [True]
    assert A.ZnippyArchive.open(path2).extract_file("image.png") == png


def test_empty_archive(A, tmp_path):  # :212-223
    path, rep = _pack(A, tmp_path, [])
    assert rep.total_files == 0
    vr = A.verify_archive_integrity(path)
    assert vr.total_files == 0 and vr.chunks == 0


def test_verify_detects_corruption(A, tmp_path, oracle):  # :414-443 + fault injection
    text = oracle.real_text(300_000).tobytes()
    jar = bytes(np.random.default_rng(1).integers(0, 256, 40_000, dtype=np.uint8))
    path, _ = _pack(A, tmp_path, [("t.txt", text), ("lib.jar", jar), ("u.txt", text[:1000])])
    assert A.verify_archive_integrity(path).verified_files == 3
    t = A.read_znippy_index(path)
    rows = {p: (o, s) for p, o, s in zip(t.column("relative_path").to_pylist(), t.column("blob_offset").to_pylist(),
                                         t.column("blob_size").to_pylist())}
    raw = bytearray(open(path, "rb").read())
    raw[rows["lib.jar"][0] + 100] ^= 0x40        # store-as-is row: digest mismatch
    open(path, "wb").write(raw)
    vr = A.verify_archive_integrity(path)
    assert vr.corrupt_files == 1 and vr.verified_files == 2 and vr.corrupt_bytes == 40_000 and vr.chunks == 3
    raw[rows["t.txt"][0] + rows["t.txt"][1] // 2] ^= 0x01  # compressed row: decode error (row skipped) or mismatch
    open(path, "wb").write(raw)
    vr = A.verify_archive_integrity(path)
    assert vr.chunks == 3 and vr.verified_bytes == 1000 and vr.total_bytes in (41_000, 341_000)


def test_repro_incompressible_blobs(A, tmp_path, oracle):  # repro_crate.rs:19-66, scaled to 600 blobs
    entries = []
    for i in range(600):
        n = 1000 + (i * 7919) % 99_000
        entries.append((f"crates/c{i}.crate", oracle.gen_incompressible(n, i).tobytes()))
    path, rep = _pack(A, tmp_path, entries, no_skip=True)
    vr = A.decompress_archive(path, False, "/dev/null")
    assert vr.corrupt_files == 0 and vr.chunks == 600
    ar = A.ZnippyArchive.open(path)
    some = [entries[i] for i in range(0, 600, 137)]
    got = ar.extract_files([p for p, _ in some])
    for (p, d), g in zip(some, got):
        assert g == d


def test_lz4_archive(A, tmp_path, oracle):
    from znippy_b200 import codec
    entries = [("a.txt", oracle.real_text(200_000).tobytes()), ("b.bin", oracle.gen_binary(9 << 20).tobytes())]
    path, _ = _pack(A, tmp_path, entries, codec_id=codec.CODEC_LZ4)
    vr = A.decompress_archive(path, True, str(tmp_path / "x"))
    assert vr.corrupt_files == 0 and vr.chunks == 3
    for p, d in entries:
        assert (tmp_path / "x" / p).read_bytes() == d


def test_native_and_pyarrow_container_paths_agree(A, tmp_path, oracle):
    """decompress_archive runs fully natively by default (container.cpp + zn_decompress_rows); the pyarrow-fronted path
    and the pyarrow-written index must give the same reports and bytes."""
    entries = [("t/a.txt", oracle.real_text(150_000).tobytes()), ("t/b.jar", oracle.gen_random(30_000).tobytes()),
               ("big.bin", oracle.gen_binary(9 << 20).tobytes()), ("e.txt", b"")]
    (tmp_path / "n").mkdir(); (tmp_path / "p").mkdir()
    pn, _ = _pack(A, tmp_path / "n", entries, native_index=True)
    pp, _ = _pack(A, tmp_path / "p", entries, native_index=False)
    tn, tp = A.read_znippy_index(pn), A.read_znippy_index(pp)
    assert tn.schema.names == tp.schema.names and tn.num_rows == tp.num_rows == 5
    for col in tn.schema.names:
        if col not in ("blob_offset",):
            assert tn.column(col).to_pylist() == tp.column(col).to_pylist(), col
    for path in (pn, pp):
        r1 = A.decompress_archive(path, True, str(tmp_path / "o1"), native=True)
        r2 = A.decompress_archive(path, True, str(tmp_path / "o2"), native=False)
        assert r1 == r2 and r1.corrupt_files == 0 and r1.total_files == 4 and r1.chunks == 5
        for p, d in entries:
            assert (tmp_path / "o1" / p).read_bytes() == d and (tmp_path / "o2" / p).read_bytes() == d
    # one shard of the rows (multi-GPU: one call per GPU)
    r = A.decompress_archive(pn, False, "/dev/null", row_range=(1, 3))
    assert r.chunks == 2


def test_native_and_python_write_pipelines_agree(A, tmp_path, oracle):
    """compress_stream natively (zn_archive_writer_*) vs the Python-driven pipeline: same index columns (blob offsets
    may differ only by batching), same decoded bytes, same CompressionReport; multi-group archives (index.rs:533-581)."""
    ents = [A.ArchiveEntry("m/a.pom", oracle.real_text(90_000).tobytes(), 1, "central"),
            A.ArchiveEntry("m/b.jar", oracle.gen_random(70_000).tobytes(), 1, "central"),
            A.ArchiveEntry("p/c.whl", oracle.gen_binary(10 << 20).tobytes(), 2, "pypi"),
            A.ArchiveEntry("plain.txt", b"hello" * 1000), A.ArchiveEntry("empty.bin", b"")]
    outs = {}
    for name, native in (("n", True), ("p", False)):
        (tmp_path / name).mkdir()
        sc = A.compress_stream(str(tmp_path / name / "x.znippy"), False, native=native)
        for e in ents:
            sc.sender().send(e)
        outs[name] = (sc.output, sc.finish())
    (pn, rn), (pp, rp) = outs["n"], outs["p"]
    assert rn == rp and rn.total_files == 5 and rn.uncompressed_files == 1 and rn.chunks == 6
    tn, tp = A.read_znippy_index(pn), A.read_znippy_index(pp)
    for col in ("relative_path", "chunk_seq", "fdata_offset", "compressed", "uncompressed_size", "blob_size", "checksum"):
        assert tn.column(col).to_pylist() == tp.column(col).to_pylist(), col
    assert [(e[0], e[1], e[5]) for e in A.read_znippy_manifest(pn)] == [(0, "", 2), (1, "central", 2), (2, "pypi", 2)]
    ar = A.ZnippyArchive.open(pn)
    for e in ents:
        assert ar.extract_file(e.relative_path) == e.data
    vr = A.verify_archive_integrity(pn)
    assert vr.corrupt_files == 0 and vr.total_files == 5 and vr.chunks == 6


def test_native_random_access(A, tmp_path, oracle):
    """ZnippyArchive over zn_archive_*: list / contains / file_size / extract_files incl. a missing path, a multi-chunk
    file, an empty file and a corrupt chunk (archive.rs:20-168)."""
    entries = [(f"f/{i:03d}.txt", oracle.real_text(1000 + 997 * i).tobytes()) for i in range(40)]
    entries += [("big.bin", oracle.gen_binary(20 << 20).tobytes()), ("lib.jar", oracle.gen_random(12345).tobytes()), ("nil", b"")]
    path, _ = _pack(A, tmp_path, entries)
    ar = A.ZnippyArchive.open(path)
    assert ar.list_files() == [p for p, _ in entries]
    assert ar.contains("big.bin") and not ar.contains("nope") and ar.file_size("big.bin") == 20 << 20 and ar.file_size("nope") is None
    want = dict(entries)
    names = ["big.bin", "f/007.txt", "missing.txt", "nil", "lib.jar", "f/039.txt", "f/007.txt"]
    got = ar.extract_files(names)
    for nme, g in zip(names, got):
        if nme == "missing.txt":
            assert isinstance(g, KeyError)
        else:
            assert g == want[nme], nme
    assert ar.extract_file("f/000.txt") == want["f/000.txt"]
    ar.close()
    # corrupt the second chunk of big.bin: that file fails, the others still extract
    t = A.read_znippy_index(path)
    rows = [(p, o, s) for p, o, s, q in zip(t.column("relative_path").to_pylist(), t.column("blob_offset").to_pylist(),
                                          t.column("blob_size").to_pylist(), t.column("chunk_seq").to_pylist()) if p == "big.bin" and q == 1]
    raw = bytearray(open(path, "rb").read())
    raw[rows[0][1]] ^= 0xFF   # frame magic
    open(path, "wb").write(raw)
    ar = A.ZnippyArchive.open(path)
    got = ar.extract_files(["big.bin", "f/001.txt"])
    assert isinstance(got[0], Exception) and got[1] == want["f/001.txt"]


def test_read_worker_pipeline_and_file_windows(A, tmp_path, oracle, monkeypatch):
    """The native read worker beyond one staging batch and beyond one window of open output files: a 288 MiB pattern file
    (extract: 4-5 batches, pread / GPU / pwrite overlapped over two buffer sets), 272 MiB of incompressible data (verify:
    batches cut on the blob bytes), and 60 small files extracted with at most 7 descriptors open at a time."""
    O = oracle
    big = O.gen_text(288 << 20)
    rnd = O.gen_random(272 << 20)
    small = [(f"s/{i % 5}/f{i}.txt", (f"file {i} " * (50 + 37 * i)).encode()) for i in range(60)]
    path, rep = _pack(A, tmp_path, [("big.txt", big), ("rnd.bin", rnd)] + small)
    assert rep.total_files == 62
    vr = A.decompress_archive(path, False, str(tmp_path))
    total = big.size + rnd.size + sum(len(d) for _, d in small)
    assert vr.corrupt_files == 0 and vr.verified_bytes == total and vr.chunks == 36 + 34 + 60
    monkeypatch.setenv("ZN_MAX_OPEN_FILES", "7")
    out = tmp_path / "x"
    vr = A.decompress_archive(path, True, str(out))
    assert vr.corrupt_files == 0 and vr.total_files == 62 and vr.total_bytes == total
    assert O.blake3((out / "big.txt").read_bytes()) == O.blake3(big)
    assert O.blake3((out / "rnd.bin").read_bytes()) == O.blake3(rnd)
    for p, d in small:
        assert (out / p).read_bytes() == d
    # a flipped byte inside a late batch is still attributed to its row
    with open(path, "r+b") as f:
        f.seek(200 << 20)
        b = f.read(1)
        f.seek(200 << 20)
        f.write(bytes([b[0] ^ 0x40]))
    vr = A.decompress_archive(path, False, str(tmp_path))
    assert vr.corrupt_files == 1 and vr.verified_bytes == total - (8 << 20)


@pytest.mark.parametrize("native", [True, False])
def test_row_shards_never_truncate_each_others_files(A, tmp_path, oracle, native):
    """shard_rows cuts on ROW boundaries, so a multi-chunk file can straddle two GPU shards.  Whatever the
    order in which the shards run, no shard may truncate bytes another shard has written (the later shard used to
    open the shared file with O_TRUNC).  A stale, longer file left from an earlier run is still cut to size."""
    O = oracle
    big = O.real_text(20 << 20).tobytes()  # 3 chunks
    entries = [("a.txt", b"first file\n" * 100), ("d/big.txt", big), ("z.txt", b"last file\n" * 50)]
    path, rep = _pack(A, tmp_path, entries)
    t = A.read_znippy_index(path)
    assert t.num_rows == 5
    for cut in (2, 3):  # both cuts fall inside d/big.txt
        out = tmp_path / f"x{cut}{int(native)}"
        os.makedirs(out / "d")
        (out / "d" / "big.txt").write_bytes(b"\xee" * (len(big) + 12345))  # stale leftover, longer than the file
        # the LATER shard first: with O_TRUNC in the other one its chunks would be lost
        v2 = A.decompress_archive(path, True, str(out), row_range=(cut, 5), native=native)
        v1 = A.decompress_archive(path, True, str(out), row_range=(0, cut), native=native)
        assert v1.corrupt_files == 0 and v2.corrupt_files == 0 and v1.chunks + v2.chunks == 5
        for p, d in entries:
            assert (out / p).read_bytes() == d, (cut, p)


def test_extract_cache_is_lru_and_transparent(A, tmp_path, oracle):
    """the LRU of decoded slices.  Same bytes with and without it; a repeated request is served from the
    cache (no new kernel launches); the byte budget is respected and the least recently used slices go first."""
    from znippy_b200 import codec
    O = oracle
    files = {f"pkg/f{i}.txt": O.real_text(300_000 + 1000 * i).tobytes() for i in range(6)}
    files["pkg/big.bin"] = O.gen_binary(20 << 20).tobytes()  # 3 slices
    path, _ = _pack(A, tmp_path, list(files.items()), level=3)
    plain = A.ZnippyArchive.open(path)
    ar = A.ZnippyArchive.open(path, cache_bytes=1_000_000)  # room for three of the ~300 KB files, not for an 8 MiB slice
    names = [f"pkg/f{i}.txt" for i in range(6)]
    assert ar.extract_files(names) == plain.extract_files(names) == [files[n] for n in names]
    st = ar.cache_stats()
    assert st["hits"] == 0 and st["misses"] == 6 and st["slices"] == 3 and st["evictions"] == 3 and st["bytes"] <= 1_000_000
    ctx = codec.default_ctx()
    before = ctx.launches()
    assert ar.extract_files(names[3:]) == [files[n] for n in names[3:]]      # the three most recent: all hits
    assert ctx.launches() == before and ar.cache_stats()["hits"] == 3
    assert ar.extract_file(names[0]) == files[names[0]]                       # miss: evicts f3 (least recently used)
    assert ar.extract_files([names[3], names[5]]) == [files[names[3]], files[names[5]]]
    st = ar.cache_stats()
    assert st["hits"] == 4 and st["misses"] == 8
    assert ar.extract_file("pkg/big.bin") == files["pkg/big.bin"]            # slices larger than the budget are never cached
    assert ar.cache_stats()["slices"] <= 3
    ar.set_cache(64 << 20)
    assert ar.extract_file("pkg/big.bin") == files["pkg/big.bin"] and ar.extract_file("pkg/big.bin") == files["pkg/big.bin"]
    st2 = ar.cache_stats()
    assert st2["hits"] == st["hits"] + 3 and st2["bytes"] >= 20 << 20
    ar.set_cache(0)
    assert ar.cache_stats()["slices"] == 0 and ar.extract_file(names[1]) == files[names[1]]


def test_row_shards_across_groups_parse_only_their_own_subindexes(A, tmp_path, oracle):
    """configs[4] in small: an archive of several (pkg_type, repo) groups decoded as three row shards whose cuts fall inside
    and between groups.  Each shard call opens only the sub-indexes it needs (archive rows -> local rows), yet files,
    bytes and counters are those of the one-shot call."""
    O = oracle
    entries = []
    for g in range(5):
        for i in range(7):
            entries.append(A.ArchiveEntry(f"g{g}/f{i}.txt", O.real_text(20_000 + 997 * (g * 7 + i)).tobytes(), pkg_type=1 + g % 3, repo=f"repo{g}"))
        entries.append(A.ArchiveEntry(f"g{g}/big.bin", O.gen_binary((9 << 20) + g).tobytes(), pkg_type=1 + g % 3, repo=f"repo{g}"))
    sc = A.compress_stream(str(tmp_path / "groups.znippy"), False, level=3)
    for e in entries:
        sc.sender().send(e)
    rep = sc.finish()
    assert len(A.read_znippy_manifest(sc.output)) == 5
    n = A.read_znippy_index(sc.output).num_rows
    assert n == rep.chunks == 5 * 9
    whole = A.decompress_archive(sc.output, False, "/dev/null")
    paths = A.read_znippy_index(sc.output).column("relative_path").to_pylist()
    want = {e.relative_path: e.data for e in entries}
    for cuts in ([0, 13, 31, n], [0, 9, 18, n], [0, 1, n - 1, n]):
        out = tmp_path / ("x" + "_".join(map(str, cuts)))
        reps = [A.decompress_archive(sc.output, True, str(out), row_range=(a, b)) for a, b in reversed(list(zip(cuts, cuts[1:])))]
        assert sum(r.chunks for r in reps) == n and sum(r.verified_bytes for r in reps) == whole.verified_bytes
        assert all(r.corrupt_files == 0 for r in reps)
        assert sum(r.total_files for r in reps) >= len(set(paths))  # a straddling file is counted by both of its shards
        for p, d in want.items():
            assert (out / p).read_bytes() == d, (cuts, p)
