#!/usr/bin/env python
"""Pipeline-level report in the shape of the reference's perf table (tests/tests/perf_bench.rs:94-234: compress MB/s,
decompress+verify MB/s, ratio per corpus): the native archive writer (zn_archive_writer_* = compress_stream) and reader
(zn_archive_decompress = decompress_archive) end to end, archive on /dev/shm, one GPU; small-file corpora also through
zn_archive_compress_dir (files on /dev/shm).  One JSON line per corpus.
MB = 2^20 bytes as in perf_bench.rs:27-33."""
import json, os, shutil, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from znippy_b200 import archive as A


def corpora(scale):
    mib = 1 << 20
    yield "text 500 MiB", [("text.txt", bench.text_slice(0, int(500 * mib * scale)))]
    yield "binary 500 MiB", [("data.bin", (np.arange(int(500 * mib * scale), dtype=np.uint32) % 251).astype(np.uint8))]
    yield "random 500 MiB", [("random.bin", bench.lcg_random(int(500 * mib * scale)))]
    small = bench.text_slice(0, 10240)
    yield "100k small files", [(f"d{i // 1000}/f{i}.txt", small) for i in range(int(100_000 * scale))]
    import sysconfig
    root = sysconfig.get_paths()["stdlib"]
    parts = [open(os.path.join(root, f), "rb").read() for f in sorted(os.listdir(root)) if f.endswith(".py")]
    real = np.resize(np.frombuffer(b"".join(parts), np.uint8), int(256 * mib * scale))
    yield "real text 256 MiB (python sources)", [("src.txt", real)]


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 19  # the reference's compression_level (common_config.rs:37)
    tmp = tempfile.mkdtemp(dir="/dev/shm")
    try:
        for name, entries in corpora(scale):
            total = sum(len(d) for _, d in entries)
            path = os.path.join(tmp, "a.znippy")
            best_c = best_d = best_x = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                sc = A.compress_stream(path, no_skip=False, level=level)
                for p, d in entries:
                    sc.send(A.ArchiveEntry(p, d))
                rep_c = sc.finish()
                best_c = min(best_c, time.perf_counter() - t0)
                t0 = time.perf_counter()
                rep_v = A.decompress_archive(path, False, tmp)
                best_d = min(best_d, time.perf_counter() - t0)
            assert rep_v.corrupt_files == 0 and rep_v.verified_bytes == total, (rep_v, total)
            out = os.path.join(tmp, "out")
            t0 = time.perf_counter()
            rep_x = A.decompress_archive(path, True, out)
            best_x = time.perf_counter() - t0
            assert rep_x.corrupt_files == 0
            shutil.rmtree(out, ignore_errors=True)
            dir_c = None
            if len(entries) > 1000:  # the CLI path of the reference: compress_dir over files on disk (compress_dir_bench.rs:34-42)
                src = os.path.join(tmp, "src")
                for p, d in entries:
                    os.makedirs(os.path.dirname(os.path.join(src, p)), exist_ok=True)
                    with open(os.path.join(src, p), "wb") as f:
                        f.write(bytes(d))
                dir_c = 1e9
                for rep in range(2):
                    t0 = time.perf_counter()
                    rep_d = A.compress_dir(src, os.path.join(tmp, "d.znippy"), level=level)
                    dir_c = min(dir_c, time.perf_counter() - t0)
                assert rep_d.total_files == len(entries)
                os.remove(os.path.join(tmp, "d.znippy"))
                shutil.rmtree(src, ignore_errors=True)
            size = os.path.getsize(path)
            mbs = lambda s: round(total / (1 << 20) / s, 1)
            print(json.dumps({"corpus": name, "files": len(entries), "in_MiB": round(total / (1 << 20), 1),
                              "archive_MiB": round(size / (1 << 20), 3), "ratio": round(total / size, 2), "level": level,
                              "compress_MBps": mbs(best_c), "compress_dir_MBps": mbs(dir_c) if dir_c else None,
                              "verify_MBps": mbs(best_d), "extract_to_shm_MBps": mbs(best_x)}), flush=True)
            os.remove(path)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
