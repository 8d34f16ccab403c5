// TEST INFRASTRUCTURE ONLY — compiles the team-uniform codec logic of znippy_b200/csrc (the very code the CUDA
// kernels run, with a team of one and plain-loop copies) for the host, so that frame/block/entropy parsing can be
// checked against the oracle without a GPU.  Never linked into libznippy_cuda.so.
#include <cstdlib>
#include <cstring>

#include "../../znippy_b200/csrc/lz4_decode.cuh"
#include "../../znippy_b200/csrc/zstd_par.cuh"
#include "../../znippy_b200/csrc/zpipe.cuh"

extern "C" int zn_hostemu_decode(const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap,
                                 uint32_t* produced) {
  zn::DecShared* sh = (zn::DecShared*)calloc(1, sizeof(zn::DecShared));
  uint8_t* lit = (uint8_t*)malloc(zn::kZstdBlockMax + 64);
  zn::Team t{0, 1};
  uint32_t predef = 0;
  zn::zs::init_luts(t, sh);
  // the device code may read the aligned 32-bit word around any valid byte: give the input 4-byte alignment + slack
  uint8_t* in = (uint8_t*)calloc(1, (size_t)src_len + 16);
  memcpy(in + 4, src, src_len);
  uint32_t st = zn::decode_blob(t, sh, in + 4, src_len, out, cap, lit, predef, produced);
  free(in);
  free(lit);
  free(sh);
  return (int)st;
}
extern "C" int zn_hostemu_decode_at(const uint8_t* src, uint32_t src_len, uint32_t misalign, uint8_t* out,
                                    uint32_t cap, uint32_t* produced) {
  zn::DecShared* sh = (zn::DecShared*)calloc(1, sizeof(zn::DecShared));
  uint8_t* lit = (uint8_t*)malloc(zn::kZstdBlockMax + 64);
  zn::Team t{0, 1};
  uint32_t predef = 0;
  zn::zs::init_luts(t, sh);
  uint8_t* in = (uint8_t*)calloc(1, (size_t)src_len + 32);
  memcpy(in + 8 + (misalign & 7), src, src_len);
  uint32_t st = zn::decode_blob(t, sh, in + 8 + (misalign & 7), src_len, out, cap, lit, predef, produced);
  free(in);
  free(lit);
  free(sh);
  return (int)st;
}

// block-parallel pipeline (walker, symbolic repeat offsets, table provenance, chaining), serial emulation
extern "C" int zn_hostemu_decode_par(const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap, uint32_t* produced) {
  uint8_t* in = (uint8_t*)calloc(1, (size_t)src_len + 32);
  memcpy(in + 8, src, src_len);
  uint32_t st = zn::par::host_decode_frames_par(in + 8, src_len, out, cap, produced);
  free(in);
  return (int)st;
}

extern "C" int zn_hostemu_decode_lz4_block(const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap, uint32_t* produced) {
  zn::DecShared* sh = (zn::DecShared*)calloc(1, sizeof(zn::DecShared));
  zn::Team t{0, 1};
  uint8_t* in = (uint8_t*)calloc(1, (size_t)src_len + 32);
  memcpy(in + 8, src, src_len);
  uint32_t st = zn::decode_lz4_block(t, sh, in + 8, src_len, out, cap, produced);
  free(in);
  free(sh);
  return (int)st;
}

// device-wide pipeline (zpipe.cuh: walk, fat tables, lane-per-block sequences, lane-per-stream literals, chain), serial
// emulation for one blob.  Returns 0 = decoded by the pipeline, 1 = the pipeline hands the blob to the legacy decoder.
extern "C" int zn_hostemu_decode_pipe(const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap, uint64_t* stats) {
  static zn::zp::FseD predef[zn::zp::kTabSet];
  static bool init = false;
  if (!init) { zn::zp::build_predef_set(predef); init = true; }
  uint8_t* in = (uint8_t*)calloc(1, (size_t)src_len + 32);
  memcpy(in + 8, src, src_len);
  const int rc = zn::zp::host_pipeline(in + 8, src_len, out, cap, predef, stats);
  free(in);
  return rc;
}

// the same with the blob at any of 16 alignments (the sequence stage stages its bit stream in 16-byte units)
extern "C" int zn_hostemu_decode_pipe_at(const uint8_t* src, uint32_t src_len, uint32_t misalign, uint8_t* out, uint32_t cap, uint64_t* stats) {
  static zn::zp::FseD predef[zn::zp::kTabSet];
  static bool init = false;
  if (!init) { zn::zp::build_predef_set(predef); init = true; }
  uint8_t* in = (uint8_t*)calloc(1, (size_t)src_len + 64);
  uint8_t* base = (uint8_t*)(((uintptr_t)in + 31) & ~(uintptr_t)15) + (misalign & 15);
  memcpy(base, src, src_len);
  const int rc = zn::zp::host_pipeline(base, src_len, out, cap, predef, stats);
  free(in);
  return rc;
}

// FSE decoding tables: the position-by-position construction the warps of k_ztables run (zpipe.cuh: fat_symbol_at) against
// the serial spread-and-number construction, on `n` random normalized distributions (log 5..9, with "less than one"
// symbols and zero counts).  Returns the number of disagreements; *lows = how many -1 symbols were exercised.
extern "C" int zn_hostemu_check_tables(uint32_t seed, uint32_t n, uint64_t* lows) {
  using namespace zn::zp;
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x >> 16); };
  int bad = 0;
  *lows = 0;
  for (uint32_t it = 0; it < n; it++) {
    const int k = (int)(rnd() % 3);
    int max_log, max_sym;
    table_params(k, &max_log, &max_sym);
    const int log = 5 + (int)(rnd() % (uint32_t)(max_log - 4)), size = 1 << log;
    const int nsym = 1 + (int)(rnd() % (uint32_t)(max_sym + 1));
    int16_t norm[64] = {0};
    int remaining = size;
    for (int s = 0; s < nsym && remaining > 0; s++) {
      const uint32_t r = rnd() % 8;
      if (r == 0) continue;
      if (r <= 2) { norm[s] = -1; remaining -= 1; continue; }
      int c = 1 + (int)(rnd() % (uint32_t)(remaining / 2 + 1));
      if (c > remaining) c = remaining;
      norm[s] = (int16_t)c;
      remaining -= c;
    }
    if (remaining > 0) {  // the rest goes to one symbol (a -1 symbol weighs 1 already)
      const int s = (int)(rnd() % (uint32_t)nsym);
      norm[s] = (int16_t)((norm[s] < 0 ? 1 : norm[s]) + remaining);
    }
    int sum = 0;
    for (int s = 0; s < nsym; s++) { sum += norm[s] < 0 ? 1 : norm[s]; *lows += norm[s] == -1; }
    if (sum != size) continue;
    FseD a[512], b[512];
    uint16_t next[64];
    const bool oka = build_fat_table(a, k, norm, nsym, log, next);
    const bool okb = build_fat_table_by_position(b, k, norm, nsym, log);
    if (oka != okb || (oka && memcmp(a, b, sizeof(FseD) * (size_t)size) != 0)) bad++;
  }
  return bad;
}
