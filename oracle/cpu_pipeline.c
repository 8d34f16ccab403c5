/*
 * ORACLE (test infrastructure only) — CPU restatement of the reference's worker loops over an
 * in-memory archive image, used (a) as the semantic checker for per-row status / VerifyReport
 * counters and (b) as the timed `cpu_baseline` / `--impl reference` arm of bench.py.
 *
 * Read side follows znippy-common/src/decompress.rs:105-192: N threads share one atomic row cursor
 * (:104,136); per row: fetch blob (:148-153, pread -> here a pointer into the image), decode or pass
 * through (:156-166; a codec error logs and `continue`s, counting the row in total_chunks only),
 * blake3 of the UNCOMPRESSED bytes compared with the checksum column (:172-184), optional write at the
 * caller-provided output offset (:186-189).  Write side follows stream_packer.rs:215-247: blake3 of the
 * source slice (:219) then compress_into (:230).
 *
 * The codec is injected as function pointers so the baseline can run on the real libzstd 1.5.5 (what the
 * reference's OpenZL back-end ultimately calls); NULL selects the oracle's own restatement.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const uint8_t* archive;
  const uint64_t *blob_offset, *blob_size, *fdata_offset, *uncompressed_size, *out_off;
  const uint8_t *compressed, *checksums;
  uint64_t n_rows;
  uint8_t* out_base;
  zn_ref_zstd_decompress_fn dfn;
  zn_ref_zstd_iserror_fn efn;
  int fast_hash;
  atomic_ullong cursor;
} read_job;

typedef struct { read_job* job; zn_ref_verify_stats st; } read_worker;

static void* read_main(void* arg) {
  read_worker* w = (read_worker*)arg;
  read_job* j = w->job;
  uint8_t* out_buf = NULL; /* reused across rows, decompress.rs:132-133 */
  size_t out_cap = 0;
  for (;;) {
    uint64_t row = atomic_fetch_add_explicit(&j->cursor, 1, memory_order_relaxed);
    if (row >= j->n_rows) break;
    w->st.total_chunks++;
    const uint8_t* blob = j->archive + j->blob_offset[row];
    size_t blob_size = (size_t)j->blob_size[row];
    const uint8_t* out;
    size_t len;
    if (j->compressed[row]) {
      size_t want = (size_t)j->uncompressed_size[row];
      if (out_cap < want + 64) {
        free(out_buf);
        out_cap = want + 64;
        out_buf = (uint8_t*)malloc(out_cap);
      }
      if (j->dfn) {
        size_t r = j->dfn(out_buf, out_cap, blob, blob_size);
        if (j->efn(r)) { w->st.decode_errors++; continue; }
        len = r;
      } else {
        size_t r = 0;
        if (zn_ref_zstd_decompress(blob, blob_size, out_buf, out_cap, &r, NULL) != ZN_REF_OK) {
          w->st.decode_errors++;
          continue;
        }
        len = r;
      }
      out = out_buf;
    } else {
      out = blob;
      len = blob_size;
    }
    w->st.total_written_bytes += len;
    uint8_t dg[32];
    if (j->fast_hash) zn_ref_blake3_fast(out, len, dg); else zn_ref_blake3(out, len, dg);
    if (memcmp(dg, j->checksums + 32 * row, 32) == 0) w->st.verified_bytes += len;
    else { w->st.corrupt_bytes += len; w->st.corrupt_rows++; }
    if (j->out_base && j->out_off) memcpy(j->out_base + j->out_off[row], out, len);
  }
  free(out_buf);
  return NULL;
}

int zn_ref_decompress_rows(const uint8_t* archive, const uint64_t* blob_offset, const uint64_t* blob_size,
                           const uint64_t* fdata_offset, const uint8_t* compressed,
                           const uint64_t* uncompressed_size, const uint8_t* checksums, uint64_t n_rows,
                           int n_threads, uint8_t* out_base, const uint64_t* out_off,
                           zn_ref_zstd_decompress_fn dfn, zn_ref_zstd_iserror_fn efn, int fast_hash,
                           zn_ref_verify_stats* stats) {
  if (n_threads < 1) n_threads = 1;
  read_job job = {archive, blob_offset, blob_size, fdata_offset, uncompressed_size, out_off,
                  compressed, checksums, n_rows, out_base, dfn, efn, fast_hash};
  atomic_init(&job.cursor, 0);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  read_worker* ws = (read_worker*)calloc(n_threads, sizeof(read_worker));
  for (int i = 0; i < n_threads; i++) { ws[i].job = &job; pthread_create(&th[i], NULL, read_main, &ws[i]); }
  memset(stats, 0, sizeof *stats);
  for (int i = 0; i < n_threads; i++) {
    pthread_join(th[i], NULL);
    stats->total_chunks += ws[i].st.total_chunks;
    stats->total_written_bytes += ws[i].st.total_written_bytes;
    stats->verified_bytes += ws[i].st.verified_bytes;
    stats->corrupt_bytes += ws[i].st.corrupt_bytes;
    stats->corrupt_rows += ws[i].st.corrupt_rows;
    stats->decode_errors += ws[i].st.decode_errors;
  }
  free(th);
  free(ws);
  return 0;
}

typedef struct {
  const uint8_t* src;
  const uint64_t *src_off, *src_len, *dst_off;
  uint64_t n;
  int level;
  uint8_t* dst;
  uint64_t* dst_len;
  uint8_t* digests;
  zn_ref_zstd_compress_fn cfn;
  zn_ref_zstd_iserror_fn efn;
  int fast_hash;
  atomic_ullong cursor;
  atomic_int failed;
} write_job;

static void* write_main(void* arg) {
  write_job* j = (write_job*)arg;
  for (;;) {
    uint64_t i = atomic_fetch_add_explicit(&j->cursor, 1, memory_order_relaxed);
    if (i >= j->n) break;
    const uint8_t* s = j->src + j->src_off[i];
    size_t len = (size_t)j->src_len[i];
    if (j->fast_hash) zn_ref_blake3_fast(s, len, j->digests + 32 * i); else zn_ref_blake3(s, len, j->digests + 32 * i);
    size_t cap = (size_t)(j->dst_off[i + 1] - j->dst_off[i]);
    size_t r = j->cfn(j->dst + j->dst_off[i], cap, s, len, j->level);
    if (j->efn(r)) { atomic_store(&j->failed, 1); j->dst_len[i] = 0; }
    else j->dst_len[i] = r;
  }
  return NULL;
}

/* dst_off has n+1 entries (capacity of slice i = dst_off[i+1]-dst_off[i]) */
int zn_ref_compress_slices(const uint8_t* src, const uint64_t* src_off, const uint64_t* src_len, uint64_t n,
                           int level, int n_threads, uint8_t* dst, const uint64_t* dst_off, uint64_t* dst_len,
                           uint8_t* digests, zn_ref_zstd_compress_fn cfn, zn_ref_zstd_iserror_fn efn,
                           int fast_hash) {
  if (!cfn || !efn) return -1;
  if (n_threads < 1) n_threads = 1;
  write_job job = {src, src_off, src_len, dst_off, n, level, dst, dst_len, digests, cfn, efn, fast_hash};
  atomic_init(&job.cursor, 0);
  atomic_init(&job.failed, 0);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, write_main, &job);
  for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
  free(th);
  return atomic_load(&job.failed) ? -1 : 0;
}
