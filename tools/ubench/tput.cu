// Issue rate of the ALU-pipe instructions of the packed Smith-Waterman loop: 8 independent chains per thread, all SMs full.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void tput(uint32_t* out, int iters, uint32_t a, uint32_t b) {
  uint32_t x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) x[k] = __viaddmax_s32(x[k], (int)a, (int)b);
      if (MODE == 1) x[k] = __viaddmax_s16x2(x[k], a, b);
      if (MODE == 2) x[k] = __viaddmax_s16x2_relu(x[k], a, b);
      if (MODE == 3) x[k] = __vimax3_s16x2(x[k], a, b);
      if (MODE == 4) x[k] = __vadd2(x[k], a);
      if (MODE == 5) x[k] = __byte_perm(x[k], a, 0x5410);
      if (MODE == 6) x[k] = x[k] * a + b;
      if (MODE == 7) x[k] = __vimax3_s32((int)x[k], (int)a, (int)b);
      if (MODE == 8) x[k] = __vmaxs2(x[k], a);
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r ^= x[k];
  if (r == 0x12345678u) out[0] = r;
}
template <int MODE> double run(uint32_t* d, int sms) {
  const int iters = 1 << 13, ctas = sms * 8, thr = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); tput<MODE><<<ctas, thr>>>(d, iters, 0xFFFEFFFEu, 0x00030003u); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)ctas * thr * iters * 8.0 / (ms * 1e-3) / 1e12;
    if (rep) best = ops > best ? ops : best;
  }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  uint32_t* d; cudaMalloc(&d, 64);
  printf("SMs %d clock %d kHz\n", p.multiProcessorCount, p.clockRate);
  printf("VIADDMNMX.S32        %.2f T thread-ops/s\n", run<0>(d, p.multiProcessorCount));
  printf("VIADDMNMX.S16x2      %.2f\n", run<1>(d, p.multiProcessorCount));
  printf("VIADDMNMX.S16x2.RELU %.2f\n", run<2>(d, p.multiProcessorCount));
  printf("VIMNMX3.S16x2        %.2f\n", run<3>(d, p.multiProcessorCount));
  printf("VIADD.16x2           %.2f\n", run<4>(d, p.multiProcessorCount));
  printf("PRMT                 %.2f\n", run<5>(d, p.multiProcessorCount));
  printf("IMAD                 %.2f\n", run<6>(d, p.multiProcessorCount));
  printf("VIMNMX3.S32          %.2f\n", run<7>(d, p.multiProcessorCount));
  printf("VIMNMX.S16x2         %.2f\n", run<8>(d, p.multiProcessorCount));
  return 0;
}
