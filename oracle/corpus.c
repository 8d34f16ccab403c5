/*
 * ORACLE (test infrastructure only) — synthetic corpus generators, byte-identical to the reference's
 * tests/tests/perf_bench.rs:74-92 (text / binary / random) and tests/tests/repro_crate.rs:8-16
 * (incompressible).  `phase`/`start` let a caller generate a slice from the middle of a long file
 * without materialising the file.
 */
#include "oracle.h"

static const char PHRASE[] = "The quick brown fox jumps over the lazy dog. "; /* 45 bytes */

void zn_ref_gen_text(uint8_t* dst, size_t len, size_t phase) {
  size_t k = phase % 45;
  for (size_t i = 0; i < len; i++) {
    dst[i] = (uint8_t)PHRASE[k];
    if (++k == 45) k = 0;
  }
}

void zn_ref_gen_binary(uint8_t* dst, size_t len, size_t start) {
  size_t k = start % 251;
  for (size_t i = 0; i < len; i++) {
    dst[i] = (uint8_t)k;
    if (++k == 251) k = 0;
  }
}

void zn_ref_gen_random(uint8_t* dst, size_t len) {
  uint64_t val = 12345;
  for (size_t i = 0; i < len; i++) {
    val = val * 6364136223846793005ULL + 1ULL;
    dst[i] = (uint8_t)(val >> 33);
  }
}

void zn_ref_gen_incompressible(uint8_t* dst, size_t len, uint64_t seed) {
  uint64_t v = seed * 0x9E3779B97F4A7C15ULL + 1ULL;
  for (size_t i = 0; i < len; i++) {
    v = v * 6364136223846793005ULL + 1442695040888963407ULL;
    dst[i] = (uint8_t)(v >> 33);
  }
}
