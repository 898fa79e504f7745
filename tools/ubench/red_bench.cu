// Microbenchmark: cost of split-K partial-sum accumulation into a small fp32 buffer with L2 reductions.
// 128 CTAs x 128 threads; every thread adds 32 values; 16 CTAs hit the same 16 KB tile (like the O projection).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void red_scalar(float* out, int ld, int iters) {
  const int tile = blockIdx.x / 16, n = tile * 128 + threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int j = 0; j < 32; ++j) atomicAdd(out + (size_t)j * ld + n, 1.0f);  // lanes -> consecutive features
}
__global__ void red_v4(float* out, int iters) {  // transposed layout out_T[n][32]: 8 x red.v4 per thread
  const int tile = blockIdx.x / 16, n = tile * 128 + threadIdx.x;
  float* o = out + (size_t)n * 32;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(o + 4 * j), "f"(1.0f) : "memory");
}
__global__ void red_v4_coal(float* out, int iters) {  // lanes cover consecutive 16-byte groups (512 B per warp op)
  const int tile = blockIdx.x / 16;
  float* o = out + (size_t)tile * 4096;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(o + (j * 128 + threadIdx.x) * 4), "f"(1.0f) : "memory");
}
__global__ void red_v2_coal(float* out, int iters) {
  const int tile = blockIdx.x / 16;
  float* o = out + (size_t)tile * 4096;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(o + (j * 128 + threadIdx.x) * 2), "f"(1.0f) : "memory");
}
__global__ void store_v4(float* out, int iters) {  // plain partial stores (no reduction), 16 KB per CTA
  float4* o = reinterpret_cast<float4*>(out) + (size_t)blockIdx.x * 1024;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j * 128 + threadIdx.x] = make_float4(1.f, 1.f, 1.f, 1.f);
}
template <typename F>
float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 20 * 1000.f;
}
int main() {
  float* buf; cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 0, 64 << 20);
  for (int iters : {1, 8}) {
    printf("iters %d (us per launch; launch overhead ~2-3 us included):\n", iters);
    printf("  scalar coalesced [b][n]      : %.2f\n", timeit([&] { red_scalar<<<128, 128>>>(buf, 1024, iters); }));
    printf("  v4 thread-contiguous [n][b]  : %.2f\n", timeit([&] { red_v4<<<128, 128>>>(buf, iters); }));
    printf("  v4 warp-coalesced            : %.2f\n", timeit([&] { red_v4_coal<<<128, 128>>>(buf, iters); }));
    printf("  v2 warp-coalesced            : %.2f\n", timeit([&] { red_v2_coal<<<128, 128>>>(buf, iters); }));
    printf("  plain v4 stores (2 MB)       : %.2f\n", timeit([&] { store_v4<<<128, 128>>>(buf, iters); }));
    printf("  empty-ish (1 scalar red/thr) : %.2f\n", timeit([&] { red_scalar<<<128, 128>>>(buf, 1024, 0); }));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
