/*
 * znippy-b200 CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the arithmetic on znippy's per-chunk codec + integrity
 * hot path.  Nothing in the product path (znippy_b200/, the
 * C-ABI library) may link, import or call this; only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it, as the
 * checker.
 *
 * Parity status:
 *   - BLAKE3 (reference call sites decompress.rs:172, stream_packer.rs:219,
 *     slot_packer.rs:553; crate blake3 1.8.5, not vendored): PINNED against the
 *     known-answer table of SURVEY.md §8(c) and against the official `blake3`
 *     Python bindings (same upstream implementation) in tests/test_oracle_*.py.
 *   - zstd frame decode / LZ4 block+frame decode (reference call sites
 *     codec.rs:67-78 -> openzl-sys-rs 0.2.0 -> facebook/openzl -> zstd/lz4, not
 *     vendored): payload formats PINNED against libzstd 1.5.5 / liblz4 1.9.4
 *     present in this image.  The OpenZL *envelope* around those payloads is
 *     PARITY UNPINNED (no OpenZL source, library or golden frames exist in
 *     /root/reference or in this image).
 */
#ifndef ZN_ORACLE_H
#define ZN_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- BLAKE3 (default hash mode, 32-byte output) ---- */
void zn_ref_blake3(const uint8_t* data, size_t len, uint8_t out[32]);
/* SIMD-across-chunks variant used only to make the CPU baseline honest
 * (the reference's blake3 crate uses AVX2/AVX-512); same digest. */
void zn_ref_blake3_fast(const uint8_t* data, size_t len, uint8_t out[32]);

/* ---- xxHash (zstd content checksum = XXH64 low 32 bits; LZ4 frame HC = XXH32) ---- */
uint64_t zn_ref_xxh64(const uint8_t* data, size_t len, uint64_t seed);
uint32_t zn_ref_xxh32(const uint8_t* data, size_t len, uint32_t seed);

/* ---- zstd (RFC 8878) ---- */
enum {
  ZN_REF_OK = 0,
  ZN_REF_ERR_SRC_TRUNCATED = -1,
  ZN_REF_ERR_BAD_MAGIC = -2,
  ZN_REF_ERR_DST_TOO_SMALL = -3,
  ZN_REF_ERR_CORRUPT = -4,
  ZN_REF_ERR_UNSUPPORTED = -5, /* dictionary id, reserved bits */
  ZN_REF_ERR_CHECKSUM = -6,
  ZN_REF_ERR_SIZE_MISMATCH = -7
};

/* counters filled by the decoder so tests can assert which format paths a corpus exercised */
typedef struct {
  uint32_t frames, skippable_frames;
  uint32_t blocks_raw, blocks_rle, blocks_compressed;
  uint32_t lit_raw, lit_rle, lit_huf_1stream, lit_huf_4stream, lit_treeless;
  uint32_t huf_weights_direct, huf_weights_fse;
  uint32_t mode_predefined, mode_rle, mode_fse, mode_repeat; /* summed over LL/OF/ML */
  uint32_t repcode_uses, overlap_matches;
  uint64_t sequences, literal_bytes, match_bytes;
  uint32_t checksums_verified;
} zn_ref_zstd_stats;

/* Decodes all concatenated frames in src. Returns ZN_REF_OK or an error; *out_len = bytes written. */
int zn_ref_zstd_decompress(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap,
                           size_t* out_len, zn_ref_zstd_stats* stats /* nullable */);
/* Frame_Content_Size of the first frame. returns 0 ok, 1 = unknown (field absent), <0 error */
int zn_ref_zstd_frame_content_size(const uint8_t* src, size_t src_len, uint64_t* fcs);

/* ---- LZ4 ---- */
/* raw block; returns bytes written or a negative ZN_REF_ERR_* */
long zn_ref_lz4_block_decompress(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap);
/* LZ4 frame format (magic 0x184D2204); returns ZN_REF_OK.. */
int zn_ref_lz4_frame_decompress(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap,
                                size_t* out_len);
int zn_ref_lz4_frame_content_size(const uint8_t* src, size_t src_len, uint64_t* fcs);

/* ---- corpus generators (perf_bench.rs:74-92, repro_crate.rs:8-16) ---- */
void zn_ref_gen_text(uint8_t* dst, size_t len, size_t phase);       /* 45-byte phrase cycled, starting at phase */
void zn_ref_gen_binary(uint8_t* dst, size_t len, size_t start);     /* (start+i) % 251 */
void zn_ref_gen_random(uint8_t* dst, size_t len);                   /* LCG from 12345 */
void zn_ref_gen_incompressible(uint8_t* dst, size_t len, uint64_t seed);

/* ---- CPU restatement of the reference read/write worker loops (decompress.rs:105-192,
 *      stream_packer.rs:209-249) over an in-memory archive image; used as cpu_baseline. ---- */
typedef struct {
  uint64_t total_chunks, total_written_bytes, verified_bytes, corrupt_bytes, corrupt_rows, decode_errors;
} zn_ref_verify_stats;

/* codec libraries are passed in as function pointers (resolved by the caller with dlopen/ctypes);
 * NULL -> use the oracle's own restatement. */
typedef size_t (*zn_ref_zstd_decompress_fn)(void* dst, size_t cap, const void* src, size_t len);
typedef unsigned (*zn_ref_zstd_iserror_fn)(size_t code);
typedef size_t (*zn_ref_zstd_compress_fn)(void* dst, size_t cap, const void* src, size_t len, int level);

int zn_ref_decompress_rows(const uint8_t* archive, const uint64_t* blob_offset, const uint64_t* blob_size,
                           const uint64_t* fdata_offset, const uint8_t* compressed,
                           const uint64_t* uncompressed_size, const uint8_t* checksums /* n*32 */,
                           uint64_t n_rows, int n_threads, uint8_t* out_base /* nullable */,
                           const uint64_t* out_off /* nullable */, zn_ref_zstd_decompress_fn dfn,
                           zn_ref_zstd_iserror_fn efn, int fast_hash, zn_ref_verify_stats* stats);

int zn_ref_compress_slices(const uint8_t* src, const uint64_t* src_off, const uint64_t* src_len,
                           uint64_t n, int level, int n_threads, uint8_t* dst, const uint64_t* dst_off,
                           uint64_t* dst_len, uint8_t* digests, zn_ref_zstd_compress_fn cfn,
                           zn_ref_zstd_iserror_fn efn, int fast_hash);

#ifdef __cplusplus
}
#endif
#endif
