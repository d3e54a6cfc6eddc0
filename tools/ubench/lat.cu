// Dependent-issue latency of the instructions on the Smith-Waterman critical path (one warp, clock64 around a chain).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void lat(long long* out, uint32_t a, uint32_t b, int mode) {
  __shared__ uint32_t sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i + 1) & 1023;
  __syncwarp();
  uint32_t x = threadIdx.x, y = threadIdx.x + 7;
  long long t0 = clock64();
  if (mode == 0) { for (int i = 0; i < N; ++i) x = __viaddmax_s16x2(x, a, b); }
  else if (mode == 1) { for (int i = 0; i < N; ++i) x = __shfl_up_sync(0xffffffffu, x, 1) + 1; }
  else if (mode == 2) { for (int i = 0; i < N; ++i) x = sm[x & 1023]; }
  else if (mode == 3) { for (int i = 0; i < N; ++i) x = x * a + b; }
  else if (mode == 4) { for (int i = 0; i < N; ++i) x = __viaddmax_s32(x, (int)a, (int)b); }
  else if (mode == 5) { for (int i = 0; i < N; ++i) { x = __viaddmax_s16x2(x, a, b); y = __viaddmax_s16x2(y, a, b); } }   // 2 independent chains
  else if (mode == 6) { for (int i = 0; i < N; ++i) x = __byte_perm(x, a, 0x5410) ^ b; }
  else if (mode == 7) { for (int i = 0; i < N; ++i) x = __vadd2(x, a); }
  else if (mode == 8) { for (int i = 0; i < N; ++i) x = __vimax3_s16x2(x, a, b); }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = x + y; }
}
int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  const char* names[] = {"VIADDMNMX.S16x2", "SHFL.UP+IADD", "LDS (pointer chase)", "IMAD", "VIADDMNMX.S32", "2x VIADDMNMX.S16x2 (indep)", "PRMT+LOP", "VIADD.16x2", "VIMNMX3.S16x2"};
  for (int m = 0; m < 9; ++m) {
    lat<<<1, 32>>>(d, 0xFFFEFFFEu, 3, m); lat<<<1, 32>>>(d, 0xFFFEFFFEu, 3, m);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-32s %.2f cycles per loop iteration\n", names[m], (double)h[0] / N);
  }
  return 0;
}
