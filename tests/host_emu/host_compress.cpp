// TEST INFRASTRUCTURE ONLY — the warp-uniform compressor logic of znippy_b200/csrc/compress.cuh compiled for the
// host (a "warp" of one lane), with the frame assembly of compress_kernels.cuh restated in plain C++, so that the
// emitted LZ4 / Zstandard frames can be fed to stock liblz4 / libzstd without a GPU.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../znippy_b200/csrc/compress.cuh"

using namespace zn;

extern "C" long zn_hostemu_compress(int codec, const uint8_t* src, uint64_t n, uint8_t* dst) {
  cz::Warp w{0, 1};
  uint64_t op = 0;
  if (codec == 2) {
    std::vector<uint16_t> tab(1u << cz::kLz4HashLog);
    std::vector<uint8_t> tmp(cz::kLz4Slot);
    op = cz::lz4_frame_header(dst, n);
    for (uint64_t o = 0; o < n; o += cz::kLz4Block) {
      const uint32_t bn = (uint32_t)(n - o < cz::kLz4Block ? n - o : cz::kLz4Block);
      const uint32_t c = cz::lz4_compress_block(w, src + o, bn, tmp.data(), tab.data());
      const bool raw = c >= bn;
      const uint32_t sz = raw ? bn : c, word = sz | (raw ? 0x80000000u : 0u);
      memcpy(dst + op, &word, 4);
      memcpy(dst + op + 4, raw ? src + o : tmp.data(), sz);
      op += 4 + sz;
    }
    memset(dst + op, 0, 4);
    return (long)(op + 4);
  }
  if (n == 0) {
    const uint8_t e[9] = {0x28, 0xB5, 0x2F, 0xFD, 0x20, 0x00, 0x01, 0x00, 0x00};
    memcpy(dst, e, 9);
    return 9;
  }
  std::vector<uint16_t> tab(1u << 14);
  std::vector<cz::V16> win(cz::WinHigh::kSmem / 16 + 2);
  std::vector<uint8_t> stage(cz::kZstdSlot + 64);
  std::vector<uint64_t> seqs(cz::kZstdMaxSeq);
  op = cz::zstd_frame_header(dst, n);
  for (uint64_t o = 0; o < n; o += cz::kZstdCBlock) {
    const uint32_t bn = (uint32_t)(n - o < cz::kZstdCBlock ? n - o : cz::kZstdCBlock);
    uint32_t poff = 0;
    const uint32_t c = (codec == 12 ? cz::zstd_compress_block<cz::WinHigh> : codec == 11 ? cz::zstd_compress_block<cz::WinMid> : cz::zstd_compress_block<cz::WinFast>)(w, src, (uint32_t)o, bn, stage.data(), seqs.data(), reinterpret_cast<uint8_t*>(win.data()), tab.data(), &poff);
    const uint32_t last = o + bn == n;
    if (c == 0) {
      cz::zstd_block_header(dst + op, last, 0, bn);
      memcpy(dst + op + 3, src + o, bn);
      op += 3 + bn;
    } else {
      cz::zstd_block_header(dst + op, last, 2, c);
      memcpy(dst + op + 3, stage.data() + poff, c);
      op += 3 + c;
    }
  }
  return (long)op;
}
