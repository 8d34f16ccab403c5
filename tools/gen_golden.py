#!/usr/bin/env python
"""Generates tests/golden/*.json — known-answer vectors for the hot path.

Sources of truth (all present in this image, none of them this repo's code):
  * official `blake3` Python bindings 1.0.8 (same upstream implementation as the reference's blake3 crate 1.8.5)
  * libzstd 1.5.5 (ZSTD_compress) and liblz4 1.9.4 (LZ4_compress_default / LZ4F_compressFrame) for compressed frames
The reference holds no golden vectors of its own for this path (round-trip tests only), so these
are the committed fixtures both the oracle and the CUDA path are checked against.  Inputs are regenerated from
the corpus generators of perf_bench.rs:74-92 / repro_crate.rs:8-16, so only (generator, size) -> hex is stored.

Run:  python tools/gen_golden.py       (rewrites tests/golden/)
"""
import base64
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (only for the corpus generators + library bindings)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen(kind, n, seed=0):
    if kind == "text":
        return O.gen_text(n)
    if kind == "binary":
        return O.gen_binary(n)
    if kind == "random":
        return O.gen_random(n)
    if kind == "incompressible":
        return O.gen_incompressible(n, seed)
    if kind == "realtext":
        return O.real_text(n)
    raise ValueError(kind)


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- blake3 KATs
    sizes = [0, 1, 63, 64, 65, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 4096, 10240, 31744, 32768, 32769,
             65536, 100000, 1 << 20, (1 << 20) + 1, 8 << 20]
    kats = []
    for kind in ("text", "binary", "random"):
        for n in sizes:
            kats.append({"gen": kind, "n": n, "blake3": O.blake3_official(gen(kind, n)).hex()})
    for seed, n in [(0, 1000), (1, 5000), (2, 10000), (3, 50000), (4, 100000)]:
        kats.append({"gen": "incompressible", "n": n, "seed": seed,
                     "blake3": O.blake3_official(gen("incompressible", n, seed)).hex()})
    json.dump(kats, open(os.path.join(OUT, "blake3_kat.json"), "w"), indent=0)

    # ---- compressed-frame fixtures (small: committed as base64)
    z, l = O.libzstd(), O.liblz4()
    frames = []

    def add(name, codec, blob, data):
        frames.append({"name": name, "codec": codec, "blob_b64": base64.b64encode(blob).decode(),
                       "out_len": int(len(data)), "out_blake3": O.blake3_official(data).hex()})

    add("zstd_empty", "zstd", z.compress(b"", 19), b"")
    add("zstd_text_10k_l19", "zstd", z.compress(gen("text", 10240), 19), gen("text", 10240))
    add("zstd_text_8m_l19", "zstd", z.compress(gen("text", 8 << 20), 19), gen("text", 8 << 20))
    add("zstd_binary_8m_l19", "zstd", z.compress(gen("binary", 8 << 20), 19), gen("binary", 8 << 20))
    add("zstd_binary_4096_l3", "zstd", z.compress(gen("binary", 4096), 3), gen("binary", 4096))
    add("zstd_incompressible_1000", "zstd", z.compress(gen("incompressible", 1000, 7), 19), gen("incompressible", 1000, 7))
    codec_rs = (b"Hello world! This is a test of compression roundtrip. Repeated data helps compression. "
                b"Repeated data helps compression. Repeated data helps compression.")
    add("zstd_codec_rs_roundtrip_l3", "zstd", z.compress(codec_rs, 3), codec_rs)
    rt = gen("realtext", 40000)
    for lvl in (1, 3, 19):
        add(f"zstd_realtext_40k_l{lvl}", "zstd", z.compress(rt, lvl), rt)
    add("zstd_realtext_40k_l3_checksum", "zstd", z.compress(rt, 3, checksum=True), rt)
    rle = bytes([7]) * 70000 + gen("realtext", 3000).tobytes() + bytes([9]) * 200000
    add("zstd_rle_mix_l3", "zstd", z.compress(rle, 3), rle)
    add("lz4f_text_10k", "lz4f", l.compress_frame(gen("text", 10240)), gen("text", 10240))
    add("lz4f_realtext_40k", "lz4f", l.compress_frame(rt), rt)
    add("lz4f_empty", "lz4f", l.compress_frame(b""), b"")
    add("lz4f_incompressible_100k", "lz4f", l.compress_frame(gen("incompressible", 100000, 3)), gen("incompressible", 100000, 3))
    add("lz4block_text_10k", "lz4block", l.compress_block(gen("text", 10240)), gen("text", 10240))
    add("lz4block_realtext_40k_hc9", "lz4block", l.compress_block(rt, 9), rt)
    # frame-size facts recorded in SURVEY.md §8(c)
    facts = {"zstd_text_10k_l19_len": len(z.compress(gen("text", 10240), 19)),
             "zstd_text_8m_l19_len": len(z.compress(gen("text", 8 << 20), 19)),
             "zstd_binary_8m_l19_len": len(z.compress(gen("binary", 8 << 20), 19)),
             "zstd_random_1m_l19_len": len(z.compress(gen("random", 1 << 20), 19)),
             "zstd_empty_hex": z.compress(b"", 19).hex(),
             "lz4block_text_10k_len": len(l.compress_block(gen("text", 10240))),
             "lz4block_text_8m_len": len(l.compress_block(gen("text", 8 << 20))),
             "libzstd": z.version(), "liblz4": int(l.l.LZ4_versionNumber())}
    json.dump({"frames": frames, "facts": facts}, open(os.path.join(OUT, "frames.json"), "w"), indent=0)
    print("wrote", len(kats), "KATs and", len(frames), "frames; facts:", facts)


if __name__ == "__main__":
    main()
