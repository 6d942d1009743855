// Microbenchmark: lsi_point_finish (gcd path of a parked coordinate) with realistic operands.
#include <cstdio>
#include "rjb_exact.cuh"
using namespace rjb;
__global__ void k(const PointState* in, long long* out, long long* cyc, int reps, int what) {
  PointState st = in[threadIdx.x];
  long long acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    PointState s2 = st;
    s2.rs += 2 * r;
    if (what == 0) acc += lsi_point_finish(s2);
    else if (what == 1) acc += (long long) gcd64(s2.rs, s2.aden);
    else { const unsigned long long g = 3; acc += (long long) (s2.aden / (g + r) + s2.rs / (g + r)); }
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = (t1 - t0) / reps;
}
int main() {
  PointState h[32]; uint64_t s = 88172645463325252ull;
  for (int i = 0; i < 32; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    h[i].aden = (s >> 3) | 1ull << 60;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    h[i].rs = (i & 1) ? h[i].aden - (s >> 9) % (h[i].aden / 40) - 1 : (s >> 9) % (h[i].aden / 40) + 1;
    h[i].X0 = (long long) (s >> 20) - (1ll << 43);
  }
  PointState* d; long long *o, *c; cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 256); cudaMalloc(&c, 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"lsi_point_finish", "gcd64 alone", "two u64 divisions"};
  for (int what = 0; what < 3; what++) {
    for (int w = 0; w < 2; w++) { k<<<1, 32>>>(d, o, c, 20, what); cudaDeviceSynchronize(); }
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    printf("%-20s %8lld cycles per call (one warp, 32 different operands)\n", names[what], cy);
  }
  for (int n = 13; n <= 13; n++) { k<<<1, n>>>(d, o, c, 20, 0); cudaDeviceSynchronize(); long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); printf("13 lanes: %lld cycles\n", cy); }
  return 0;
}
