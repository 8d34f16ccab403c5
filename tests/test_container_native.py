"""The native `.znippy` v0.7 container (znippy_b200/csrc/container.cpp, no Arrow library) against pyarrow, both ways:
pyarrow reads what the C++ writer emits, the C++ reader reads what pyarrow (the Python path and, by format, the Rust
reference: index.rs:43-54, 279-288, meta_sink.rs:71-118) emits.  CPU only."""
import ctypes as C
import os

import numpy as np
import pyarrow as pa
import pytest


@pytest.fixture(scope="module")
def L():
    from znippy_b200 import _native as N
    return N.lib()


def _rows(n, seed=0, compressed_every=3, payload=77):
    """Random index rows that are VALID for an archive with `payload` bytes of blobs (the reader rejects rows whose blob
    lies outside the payload region or whose sizes reach 4 GiB)."""
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        bo = int(rng.integers(0, payload + 1))
        rows.append((f"dir{i % 7}/file_{i:05d}.txt" if i % 11 else "", i % 4, int(rng.integers(0, 1 << 40)), i % compressed_every != 0,
                     int(rng.integers(0, 1 << 32)), bo, int(rng.integers(0, payload - bo + 1)),
                     bytes(rng.integers(0, 256, 32, dtype=np.uint8))))
    return rows


def _open(L, path):
    err = C.create_string_buffer(256)
    h = L.zn_index_open(str(path).encode(), err, 256)
    assert h, err.value
    return h


def _dump(L, h):
    n = L.zn_index_rows(h)
    cols = [np.ctypeslib.as_array(L.zn_index_u64(h, k), (n,)).copy() if n else np.zeros(0, np.uint64) for k in range(4)]
    seq = np.ctypeslib.as_array(L.zn_index_chunk_seq(h), (n,)).copy() if n else np.zeros(0, np.uint32)
    comp = np.ctypeslib.as_array(L.zn_index_compressed(h), (n,)).copy() if n else np.zeros(0, np.uint8)
    sums = np.ctypeslib.as_array(L.zn_index_checksums(h), (n * 32,)).copy().reshape(n, 32) if n else np.zeros((0, 32), np.uint8)
    paths = []
    for r in range(n):
        ln = C.c_uint32(0)
        p = L.zn_index_path(h, r, C.byref(ln))
        paths.append(C.string_at(p, ln.value).decode())
    return n, cols, seq, comp, sums, paths


def test_native_reader_reads_pyarrow_archives(L, tmp_path):
    from znippy_b200 import archive as A
    p = tmp_path / "a.znippy"
    g0, g1, g2 = _rows(1000, 1), _rows(1, 2), _rows(0, 3)
    schema = A.INDEX_SCHEMA.with_metadata(A.config_metadata())
    # a plugin-extended sub-index (index.rs:63-70): extra nullable columns after the base eight must be skipped
    ext = pa.schema(list(A.INDEX_SCHEMA) + [pa.field("pkg_type", pa.int8(), True), pa.field("group_id", pa.utf8(), True)])
    with open(p, "wb") as f:
        f.write(b"B" * 77)
        sink = A.ArrowIpcSink(f, 77)
        sink.push_subindex((0, ""), schema, [A.build_metadata_batch(g0[:600], schema), A.build_metadata_batch(g0[600:], schema)])
        b1 = A.build_metadata_batch(g1, A.INDEX_SCHEMA)
        b1x = pa.record_batch(list(b1.columns) + [pa.array([3], pa.int8()), pa.array([None], pa.utf8())], schema=ext)
        sink.push_subindex((3, "central"), ext, [b1x])
        sink.push_subindex((4, "empty"), schema, [A.build_metadata_batch(g2, schema)])
        sink.finish()
    h = _open(L, p)
    try:
        n, cols, seq, comp, sums, paths = _dump(L, h)
        want = g0 + g1 + g2
        assert n == len(want) == 1001
        assert paths == [r[0] for r in want]
        assert seq.tolist() == [r[1] for r in want]
        assert cols[2].tolist() == [r[2] for r in want] and cols[3].tolist() == [r[4] for r in want]
        assert cols[0].tolist() == [r[5] for r in want] and cols[1].tolist() == [r[6] for r in want]
        assert comp.tolist() == [int(r[3]) for r in want]
        assert sums.tobytes() == b"".join(r[7] for r in want)
        assert L.zn_index_groups(h) == 3
        pk, repo, io, il, rc = C.c_int8(), C.c_char_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        assert L.zn_index_group(h, 1, C.byref(pk), C.byref(repo), C.byref(io), C.byref(il), C.byref(rc)) == 0
        assert (pk.value, repo.value, rc.value) == (3, b"central", 1)
        assert L.zn_index_metadata(h, b"znippy_format_version") == b"3"
        assert L.zn_index_metadata(h, b"compression_level") == b"19" and L.zn_index_metadata(h, b"nope") is None
        assert [L.zn_index_field_name(h, i).decode() for i in range(L.zn_index_field_count(h))] == A.INDEX_SCHEMA.names
    finally:
        L.zn_index_close(h)


def test_pyarrow_reads_native_writer(L, tmp_path):
    from znippy_b200 import archive as A
    p = tmp_path / "w.znippy"
    groups = [((0, ""), _rows(500, 5, payload=1234)), ((2, "pypi"), _rows(3, 6, payload=1234)), ((7, "x"), _rows(0, 7))]
    fd = os.open(p, os.O_CREAT | os.O_RDWR, 0o644)
    os.pwrite(fd, b"Z" * 1234, 0)
    w = L.zn_index_writer_create(fd, 1234)
    for k, v in A.config_metadata().items():
        assert L.zn_index_writer_metadata(w, k.encode(), v.encode()) == 0
    for (pt, repo), rows in groups:
        n = len(rows)
        paths = (C.c_char_p * max(n, 1))(*[r[0].encode() for r in rows])
        seq = np.array([r[1] for r in rows], np.uint32)
        fo = np.array([r[2] for r in rows], np.uint64)
        cp = np.array([int(r[3]) for r in rows], np.uint8)
        us = np.array([r[4] for r in rows], np.uint64)
        bo = np.array([r[5] for r in rows], np.uint64)
        bs = np.array([r[6] for r in rows], np.uint64)
        ck = np.frombuffer(b"".join(r[7] for r in rows), np.uint8) if n else np.zeros(1, np.uint8)
        ptr = lambda a: C.c_void_p(a.ctypes.data)
        assert L.zn_index_writer_push_group(w, pt, repo.encode(), n, paths, ptr(seq), ptr(fo), ptr(cp), ptr(us), ptr(bo), ptr(bs),
                                            ptr(ck)) == 0
    assert L.zn_index_writer_finish(w) == 0
    os.close(fd)
    # pyarrow (the same Arrow IPC the Rust reference reads with arrow-rs) accepts it, field for field
    t = A.read_znippy_index(str(p))
    want = [r for _, rows in groups for r in rows]
    assert t.schema.names == A.INDEX_SCHEMA.names
    assert [f.nullable for f in t.schema] == [False] * 8 and t.schema.field("checksum").type == pa.binary(32)
    assert t.schema.field("chunk_seq").type == pa.uint32() and t.schema.field("compressed").type == pa.bool_()
    assert t.column("relative_path").to_pylist() == [r[0] for r in want]
    assert t.column("chunk_seq").to_pylist() == [r[1] for r in want]
    assert t.column("fdata_offset").to_pylist() == [r[2] for r in want]
    assert t.column("compressed").to_pylist() == [r[3] for r in want]
    assert t.column("uncompressed_size").to_pylist() == [r[4] for r in want]
    assert t.column("blob_offset").to_pylist() == [r[5] for r in want]
    assert t.column("blob_size").to_pylist() == [r[6] for r in want]
    assert t.column("checksum").to_pylist() == [r[7] for r in want]
    md = {k.decode(): v.decode() for k, v in t.schema.metadata.items()}
    assert md == A.config_metadata()
    man = A.read_znippy_manifest(str(p))
    assert [(e[0], e[1], e[5]) for e in man] == [(0, "", 500), (2, "pypi", 3), (7, "x", 0)]
    raw = open(p, "rb").read()
    assert raw[:1234] == b"Z" * 1234 and raw[-16:-8] == b"ZNPYMIDX"
    # and the native reader reads the native writer
    h = _open(L, p)
    try:
        n, cols, seq, comp, sums, paths = _dump(L, h)
        assert n == 503 and paths == [r[0] for r in want] and sums.tobytes() == b"".join(r[7] for r in want)
    finally:
        L.zn_index_close(h)


def test_native_reader_rejects_bad_archives(L, tmp_path):
    err = C.create_string_buffer(256)
    p = tmp_path / "bad.znippy"
    p.write_bytes(b"short")
    assert not L.zn_index_open(str(p).encode(), err, 256) and b"too small" in err.value
    p.write_bytes(b"x" * 100 + (8).to_bytes(8, "little") * 2)
    assert not L.zn_index_open(str(p).encode(), err, 256) and b"v0.6" in err.value  # index.rs:387-389
    p.write_bytes(b"x" * 100 + b"ZNPYMIDX" + (4000).to_bytes(8, "little"))
    assert not L.zn_index_open(str(p).encode(), err, 256)
    p.write_bytes(b"x" * 100 + b"ZNPYMIDX" + (10).to_bytes(8, "little"))   # garbage where the manifest should be
    assert not L.zn_index_open(str(p).encode(), err, 256)
    assert not L.zn_index_open(str(tmp_path / "missing").encode(), err, 256)


def _write_archive(A, path, rows, payload=b"B" * 77, manifest_patch=None):
    schema = A.INDEX_SCHEMA.with_metadata(A.config_metadata())
    with open(path, "wb") as f:
        f.write(payload)
        sink = A.ArrowIpcSink(f, len(payload))
        sink.push_subindex((0, ""), schema, [A.build_metadata_batch(rows, schema)])
        sink.finish()


def test_native_reader_rejects_hostile_index_rows(L, tmp_path):
    """index columns are untrusted.  Rows whose blob lies outside the payload region, or whose sizes
    reach 4 GiB (and could wrap the u64 sums of a batch), are refused when the index is opened."""
    from znippy_b200 import archive as A
    err = C.create_string_buffer(256)
    ck = bytes(32)
    p = tmp_path / "h.znippy"
    good = ("a", 0, 0, True, 100, 10, 60, ck)
    _write_archive(A, p, [good])
    L.zn_index_close(_open(L, p))
    for bad, what in [(("a", 0, 0, True, 100, 10, 68, ck), b"outside the payload"),          # 10 + 68 > 77
                      (("a", 0, 0, True, 100, 78, 0, ck), b"outside the payload"),
                      (("a", 0, 0, True, 100, (1 << 64) - 16, 32, ck), b"outside the payload"),  # offset + size wraps
                      (("a", 0, 0, True, 1 << 63, 0, 10, ck), b"4 GiB"),
                      (("a", 0, 0, True, 100, 0, (1 << 63) - 16, ck), b"4 GiB"),
                      (("a", 0, 1 << 63, True, 100, 0, 10, ck), b"file offset")]:
        _write_archive(A, p, [good, bad])
        assert not L.zn_index_open(str(p).encode(), err, 256), bad
        assert what in err.value, (bad, err.value)


def test_native_reader_requires_exact_manifest_types(L, tmp_path):
    """the hand-written Arrow IPC reader must not trust the schema — a manifest whose integer
    columns have another width than index.rs declares would make get_uint read past the buffer."""
    from znippy_b200 import archive as A
    err = C.create_string_buffer(256)
    p = tmp_path / "m.znippy"
    _write_archive(A, p, [("a", 0, 0, True, 100, 10, 60, bytes(32))])
    raw = bytearray(p.read_bytes())
    moff = int.from_bytes(raw[-8:], "little")
    man = pa.ipc.open_stream(pa.BufferReader(bytes(raw[moff:-16]))).read_all()
    assert man.schema.field("index_offset").type == pa.uint64()
    for col, typ in [("index_offset", pa.uint8()), ("row_count", pa.uint16()), ("pkg_type", pa.int64()), ("repo", pa.binary())]:
        i = man.schema.get_field_index(col)
        vals = man.column(col).to_pylist()
        if typ == pa.uint8():
            vals = [v & 0xFF for v in vals]
        if typ == pa.binary():
            vals = [v.encode() for v in vals]
        t2 = man.set_column(i, pa.field(col, typ, False), pa.array(vals, typ))
        sink = pa.BufferOutputStream()
        with pa.ipc.new_stream(sink, t2.schema) as w:
            w.write_table(t2)
        blob = sink.getvalue().to_pybytes()
        p.write_bytes(bytes(raw[:moff]) + blob + b"ZNPYMIDX" + moff.to_bytes(8, "little"))
        assert not L.zn_index_open(str(p).encode(), err, 256), col


def test_native_reader_rejects_unknown_envelope_declaration(L, tmp_path):
    """An archive may declare a blob envelope in its index metadata (znippy_envelope); one this reader does not know is
    refused at open instead of being decoded as bare frames."""
    from znippy_b200 import archive as A
    err = C.create_string_buffer(256)
    p = tmp_path / "e.znippy"
    for value, ok in (("ZNB1", True), ("OZL9", False)):
        md = dict(A.config_metadata(), znippy_envelope=value)
        schema = A.INDEX_SCHEMA.with_metadata(md)
        with open(p, "wb") as f:
            f.write(b"B" * 77)
            sink = A.ArrowIpcSink(f, 77)
            sink.push_subindex((0, ""), schema, [A.build_metadata_batch([("a", 0, 0, True, 100, 10, 60, bytes(32))], schema)])
            sink.finish()
        h = L.zn_index_open(str(p).encode(), err, 256)
        assert bool(h) == ok, (value, err.value)
        if h:
            assert L.zn_index_metadata(h, b"znippy_envelope") == b"ZNB1"
            L.zn_index_close(h)
        else:
            assert b"envelope" in err.value
