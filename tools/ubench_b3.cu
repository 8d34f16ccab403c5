// Microbenchmark (development tool, not product): instruction-mix variants of the BLAKE3 compression function on
// sm_100a, to find the split between the ALU pipe (LOP3/SHF/PRMT/IADD3) and the FMA pipe (IMAD) that issues fastest.
// Each variant hashes the same 32 chunks per warp from shared memory many times; results must agree.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_b3 tools/ubench_b3.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define IV0 0x6A09E667u
#define IV1 0xBB67AE85u
#define IV2 0x3C6EF372u
#define IV3 0xA54FF53Au

__device__ __forceinline__ uint32_t rotr16(uint32_t x) { return __byte_perm(x, x, 0x1032); }
__device__ __forceinline__ uint32_t rotr8(uint32_t x) { return __byte_perm(x, x, 0x0321); }
__device__ __forceinline__ uint32_t rotr12(uint32_t x) { return __funnelshift_r(x, x, 12); }
__device__ __forceinline__ uint32_t rotr7(uint32_t x) { return __funnelshift_r(x, x, 7); }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t lop3_or_xor(uint32_t a, uint32_t b, uint32_t c) {  // (a | b) ^ c
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x1E;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int V>
struct G;
template <>
struct G<0> {  // baseline: what the product kernel compiles to today
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t,
                                           uint32_t, uint32_t) {
    a = a + b + mx; d = rotr16(d ^ a); c = c + d; b = rotr12(b ^ c);
    a = a + b + my; d = rotr8(d ^ a); c = c + d; b = rotr7(b ^ c);
  }
};
template <>
struct G<1> {  // message adds on the FMA pipe (IMAD with an opaque 1)
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t one,
                                           uint32_t, uint32_t) {
    a = imad(mx, one, a + b); d = rotr16(d ^ a); c = c + d; b = rotr12(b ^ c);
    a = imad(my, one, a + b); d = rotr8(d ^ a); c = c + d; b = rotr7(b ^ c);
  }
};
template <>
struct G<2> {  // every add on the FMA pipe
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t one,
                                           uint32_t, uint32_t) {
    a = imad(mx, one, imad(b, one, a)); d = rotr16(d ^ a); c = imad(d, one, c); b = rotr12(b ^ c);
    a = imad(my, one, imad(b, one, a)); d = rotr8(d ^ a); c = imad(d, one, c); b = rotr7(b ^ c);
  }
};
template <>
struct G<3> {  // c+d on FMA, a+b+m as IADD3 on ALU
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t one,
                                           uint32_t, uint32_t) {
    a = a + b + mx; d = rotr16(d ^ a); c = imad(d, one, c); b = rotr12(b ^ c);
    a = a + b + my; d = rotr8(d ^ a); c = imad(d, one, c); b = rotr7(b ^ c);
  }
};
template <>
struct G<4> {  // rotr12 / rotr7 through a 64-bit IMAD.WIDE (x * 2^(32-r): hi|lo = rotr), OR folded into the consumers
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t one,
                                           uint32_t m20, uint32_t m25) {
    a = imad(mx, one, a + b); d = rotr16(d ^ a); c = imad(d, one, c);
    uint64_t w = (uint64_t)(b ^ c) * m20;  // rotr12
    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    a = imad(my, one, a + lo + hi);
    d = rotr8(d ^ a); c = imad(d, one, c);
    const uint32_t t = lop3_or_xor(lo, hi, c);
    w = (uint64_t)t * m25;  // rotr7
    b = (uint32_t)w | (uint32_t)(w >> 32);
  }
};

template <>
struct G<5> {  // V1 + only rotr7 through IMAD.WIDE (x * 2^25: hi + lo = rotr 7), sum on the FMA pipe
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t one,
                                           uint32_t, uint32_t m25) {
    a = imad(mx, one, a + b); d = rotr16(d ^ a); c = c + d; b = rotr12(b ^ c);
    a = imad(my, one, a + b); d = rotr8(d ^ a); c = c + d;
    const uint64_t w = (uint64_t)(b ^ c) * m25;
    b = imad((uint32_t)(w >> 32), one, (uint32_t)w);
  }
};
template <>
struct G<6> {  // V1 + rotr12 and rotr7 through IMAD.WIDE
  static __device__ __forceinline__ void g(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t mx, uint32_t my, uint32_t one,
                                           uint32_t m20, uint32_t m25) {
    a = imad(mx, one, a + b); d = rotr16(d ^ a); c = c + d;
    uint64_t w = (uint64_t)(b ^ c) * m20;
    b = imad((uint32_t)(w >> 32), one, (uint32_t)w);
    a = imad(my, one, a + b); d = rotr8(d ^ a); c = c + d;
    w = (uint64_t)(b ^ c) * m25;
    b = imad((uint32_t)(w >> 32), one, (uint32_t)w);
  }
};

template <int V>
__device__ __forceinline__ void compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint32_t ctr, uint32_t blen, uint32_t flags,
                                         uint32_t one, uint32_t m20, uint32_t m25) {
  uint32_t v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
  uint32_t v8 = IV0, v9 = IV1, v10 = IV2, v11 = IV3, v12 = ctr, v13 = 0, v14 = blen, v15 = flags;
#define R(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15)   \
  G<V>::g(v0, v4, v8, v12, m[s0], m[s1], one, m20, m25);                          \
  G<V>::g(v1, v5, v9, v13, m[s2], m[s3], one, m20, m25);                          \
  G<V>::g(v2, v6, v10, v14, m[s4], m[s5], one, m20, m25);                         \
  G<V>::g(v3, v7, v11, v15, m[s6], m[s7], one, m20, m25);                         \
  G<V>::g(v0, v5, v10, v15, m[s8], m[s9], one, m20, m25);                         \
  G<V>::g(v1, v6, v11, v12, m[s10], m[s11], one, m20, m25);                       \
  G<V>::g(v2, v7, v8, v13, m[s12], m[s13], one, m20, m25);                        \
  G<V>::g(v3, v4, v9, v14, m[s14], m[s15], one, m20, m25);
  R(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
  R(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
  R(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
  R(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
  R(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
  R(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
  R(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
#undef R
  cv[0] = v0 ^ v8; cv[1] = v1 ^ v9; cv[2] = v2 ^ v10; cv[3] = v3 ^ v11;
  cv[4] = v4 ^ v12; cv[5] = v5 ^ v13; cv[6] = v6 ^ v14; cv[7] = v7 ^ v15;
}

// every lane hashes `iters` x 16 blocks whose words come from shared memory (conflict-free rows)
template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* seed, int iters, uint32_t one, uint32_t m20, uint32_t m25) {
  __shared__ uint4 rows[256 * 5];
  for (int i = threadIdx.x; i < 256 * 5; i += 256) rows[i] = make_uint4(seed[i & 63] + i, i * 3u, i * 7u, ~i);
  __syncthreads();
  uint32_t cv[8] = {IV0, IV1, IV2, IV3, 1, 2, 3, 4};
  const uint4* row = rows + threadIdx.x * 5;
  for (int it = 0; it < iters; it++) {
    uint32_t m[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 v = row[q];
      m[4 * q] = v.x ^ it; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
    }
    compress<V>(cv, m, it, 64, it & 3, one, m20, m25);
  }
  uint32_t x = 0;
  for (int i = 0; i < 8; i++) x ^= cv[i];
  out[blockIdx.x * 256 + threadIdx.x] = x;
}

template <int V>
double run(uint32_t* d_out, const uint32_t* d_seed, int grid, int iters, uint32_t* h_first) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<V><<<grid, 256>>>(d_out, d_seed, 64, 1, 1u << 20, 1u << 25);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<V><<<grid, 256>>>(d_out, d_seed, iters, 1, 1u << 20, 1u << 25);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(h_first, d_out, 4 * 256, cudaMemcpyDeviceToHost);
  return ms;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t *d_out, *d_seed;
  std::vector<uint32_t> seed(64);
  for (int i = 0; i < 64; i++) seed[i] = i * 2654435761u;
  cudaMalloc(&d_seed, 256);
  cudaMemcpy(d_seed, seed.data(), 256, cudaMemcpyHostToDevice);
  const int iters = 4096;
  for (int cps = 1; cps <= 4; cps *= 2) {  // resident CTAs per SM: 8 / 16 / 32 warps
    const int grid = sms * cps;
    cudaMalloc(&d_out, (size_t)grid * 256 * 4);
    std::vector<uint32_t> ref(256), got(256);
    double ms[7];
    ms[0] = run<0>(d_out, d_seed, grid, iters, ref.data());
    ms[1] = run<1>(d_out, d_seed, grid, iters, got.data()); bool ok1 = got == ref;
    ms[2] = run<2>(d_out, d_seed, grid, iters, got.data()); bool ok2 = got == ref;
    ms[3] = run<3>(d_out, d_seed, grid, iters, got.data()); bool ok3 = got == ref;
    ms[4] = run<4>(d_out, d_seed, grid, iters, got.data()); bool ok4 = got == ref;
    ms[5] = run<5>(d_out, d_seed, grid, iters, got.data()); bool ok5 = got == ref;
    ms[6] = run<6>(d_out, d_seed, grid, iters, got.data()); bool ok6 = got == ref;
    const double bytes = (double)grid * 256 * iters * 64;
    printf("ctas/sm=%d  ", cps);
    for (int v = 0; v < 7; v++) printf("V%d %.3f ms %.0f GB/s  ", v, ms[v], bytes / ms[v] / 1e6);
    printf(" agree=%d%d%d%d%d%d\n", ok1, ok2, ok3, ok4, ok5, ok6);
    cudaFree(d_out);
  }
  return 0;
}
