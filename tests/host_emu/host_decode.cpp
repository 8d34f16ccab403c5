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
