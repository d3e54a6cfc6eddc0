// The packed Smith-Waterman pass in isolation: W warps per CTA, one CTA per SM, synthetic profiles/windows in shared memory.
// Prints cycles per wavefront step for a warp and the aggregate cell rate (how many co-resident warps the loop needs to fill the ALU pipe).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sortmerna_b200/csrc/smr_sw.cuh"
using namespace smr;
template <int R>
__global__ void dp_kernel(unsigned long long* out, int n, int reps) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* prof = (uint32_t*)(sm + (size_t)w * (2 * kPairProfWords * 4 + 2 * 640));
  uint8_t* win = (uint8_t*)(prof + 2 * kPairProfWords);
  const SwScore sc{2, -3, -3, 5, 2, 1};
  for (int i = lane; i < 2 * kPairProfWords; i += 32) prof[i] = pack16(((i * 7) & 3) ? -3 : 2, ((i * 5) & 3) ? -3 : 2);
  for (int i = lane; i < 2 * 640; i += 32) win[i] = (uint8_t)((i * 13 + (i >> 3)) & 3);
  __syncwarp();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) acc += sw_pair_warp<R>(prof, prof + kPairProfWords, win, win + 640, n, sc);
  const long long t1 = clock64();
  if (lane == 0) { out[(blockIdx.x * (blockDim.x >> 5) + w) * 2] = (unsigned long long)(t1 - t0); out[(blockIdx.x * (blockDim.x >> 5) + w) * 2 + 1] = acc; }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  unsigned long long* d; cudaMalloc(&d, 1 << 20);
  const int n = 158, reps = 200;
  for (int W : {4, 8, 12, 16}) {
    const size_t smem = (size_t)W * (2 * kPairProfWords * 4 + 2 * 640);
    cudaFuncSetAttribute(dp_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dp_kernel<5><<<p.multiProcessorCount, W * 32, smem>>>(d, n, reps);
    cudaEventRecord(e0); dp_kernel<5><<<p.multiProcessorCount, W * 32, smem>>>(d, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double steps = (double)reps * (n + 31);
    const double cells = (double)p.multiProcessorCount * W * reps * 2.0 * n * 150.0;
    printf("R=5 warps/SMSP %d: %.1f cycles per step per warp, %.1f per SMSP-step, %.2f Tcell/s (150-row queries), err=%s\n", W / 4, h[0] / steps, h[0] / steps / (W / 4), cells / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
