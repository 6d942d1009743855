// Microbenchmark: latency of one 64-bit gcd per lane (one warp, 32 different inputs), several algorithms.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ int ctz64(u64 v) { return __ffsll((long long) v) - 1; }
__device__ __forceinline__ u64 gcd_bin(u64 a, u64 b) {
  if (a == 0) return b; if (b == 0) return a;
  int sh = ctz64(a | b); a >>= ctz64(a);
  do { b >>= ctz64(b); if (a > b) { u64 t = a; a = b; b = t; } b -= a; } while (b != 0);
  return a << sh;
}
// ctz through clz of the isolated lowest bit
__device__ __forceinline__ u64 gcd_bin_clz(u64 a, u64 b) {
  if (a == 0) return b; if (b == 0) return a;
  int sh = ctz64(a | b); a >>= ctz64(a);
  do { b >>= (63 - __clzll((long long) (b & (0 - b)))); if (a > b) { u64 t = a; a = b; b = t; } b -= a; } while (b != 0);
  return a << sh;
}
// Euclid with the hardware-less 64-bit modulo
__device__ __forceinline__ u64 gcd_euclid(u64 a, u64 b) {
  while (b) { u64 t = a % b; a = b; b = t; }
  return a;
}
// Euclid, quotient from a double division, exact correction
__device__ __forceinline__ u64 gcd_euclid_fp(u64 a, u64 b) {
  while (b) {
    u64 q = (u64) ((double) a / (double) b);
    u64 t = a - q * b;                 // may be off by a few b either way (wraps)
    while ((long long) t < 0) t += b;  // q too large
    while (t >= b) t -= b;
    a = b; b = t;
  }
  return a;
}
// binary gcd, 32-bit tail once both operands fit
__device__ __forceinline__ u64 gcd_bin_3264(u64 a, u64 b) {
  if (a == 0) return b; if (b == 0) return a;
  int sh = ctz64(a | b); a >>= ctz64(a);
  do {
    b >>= ctz64(b);
    if (((a | b) >> 32) == 0) {
      uint32_t x = (uint32_t) a, y = (uint32_t) b;
      do { y >>= (__ffs(y) - 1); if (x > y) { uint32_t t = x; x = y; y = t; } y -= x; } while (y != 0);
      return (u64) x << sh;
    }
    if (a > b) { u64 t = a; a = b; b = t; } b -= a;
  } while (b != 0);
  return a << sh;
}
template <int ALG> __global__ void k(const u64* in, u64* out, long long* cyc, int reps) {
  u64 a = in[2 * threadIdx.x], b = in[2 * threadIdx.x + 1];
  u64 acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    u64 g = ALG == 0 ? gcd_bin(a + r, b) : ALG == 1 ? gcd_bin_clz(a + r, b) : ALG == 2 ? gcd_euclid(a + r, b)
          : ALG == 3 ? gcd_euclid_fp(a + r, b) : gcd_bin_3264(a + r, b);
    acc += g;
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = (t1 - t0) / reps;
}
int main() {
  u64 h[64]; uint64_t s = 88172645463325252ull;
  for (int i = 0; i < 64; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (s >> 2) | 1ull << 60; }
  u64 *d, *o; long long* c; cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"binary (ffsll)", "binary (clz of lowest bit)", "Euclid (u64 %)", "Euclid (double quotient)", "binary, 32-bit tail"};
  u64 ref[32];
  for (int alg = 0; alg < 5; alg++) {
    for (int w = 0; w < 2; w++) {
      if (alg == 0) k<0><<<1, 32>>>(d, o, c, 20); if (alg == 1) k<1><<<1, 32>>>(d, o, c, 20);
      if (alg == 2) k<2><<<1, 32>>>(d, o, c, 20); if (alg == 3) k<3><<<1, 32>>>(d, o, c, 20); if (alg == 4) k<4><<<1, 32>>>(d, o, c, 20);
      cudaDeviceSynchronize();
    }
    long long cy; u64 res[32]; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(res, o, 256, cudaMemcpyDeviceToHost);
    if (alg == 0) for (int i = 0; i < 32; i++) ref[i] = res[i];
    bool ok = true; for (int i = 0; i < 32; i++) ok &= res[i] == ref[i];
    printf("%-28s %8lld cycles per gcd (warp of 32 different 61-bit inputs) %s\n", names[alg], cy, ok ? "same" : "DIFFERENT");
  }
  return 0;
}
