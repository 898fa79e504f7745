// Microbenchmark: floor of one "phase" of a persistent cooperative kernel on B200:
//   [load 32x64 fp32 activations from L2] -> [32 split-K reds per thread] -> [fence + software grid barrier]
// Reports per-phase in-kernel globaltimer deltas (mean over CTAs): load, reds issued, barrier.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(320, 1) phases(float* x, float* out, unsigned* ctr, unsigned long long* prof, int nph, int mode, int spin_warps, const float* big) {
  extern __shared__ float sm[];
  const int g = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  unsigned target = 0;
  if (t == 0) sm[0] = 0.f;
  __syncthreads();
  if (t >= 128) {
    // spin_warps == 3: warp 4 lane 0 streams HBM in the background through a 4 x 40 KB ring of bulk copies
    // (like the decode kernel's producer) until the phase loop is done
    if (spin_warps == 3 && t == 128) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8);
      for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar + s)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const char* src = reinterpret_cast<const char*>(big) + (size_t)g * (24u << 20);
      unsigned ph = 0; size_t off = 0; int s = 0; bool primed = false;
      volatile float* f = sm;
      while (f[0] == 0.f) {
        unsigned sa = (unsigned)__cvta_generic_to_shared(reinterpret_cast<char*>(sm) + 1024 + s * 40960);
        unsigned ba = (unsigned)__cvta_generic_to_shared(bar + s);
        if (primed) { unsigned done; do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(ba), "r"(ph) : "memory"); } while (!done); }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(40960u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa), "l"(src + off), "r"(40960u), "r"(ba) : "memory");
        off = (off + 40960) % (24u << 20);
        if (++s == 4) { s = 0; if (primed) ph ^= 1; primed = true; }
      }
      // drain
      for (int k = 0; k < 4; ++k) { unsigned ba = (unsigned)__cvta_generic_to_shared(bar + s); unsigned done; do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(ba), "r"(ph) : "memory"); } while (!done); if (++s == 4) { s = 0; ph ^= 1; } }
      return;
    }
    if (spin_warps < 3 && t < 128 + 32 * spin_warps) { volatile float* f = sm; while (f[0] == 0.f) {} }
    return;
  }
  float acc = 0.f;
  for (int ph = 0; ph < nph; ++ph) {
    unsigned long long t0 = gt();
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(x + ((t + i * 128) >> 4) * 1024 + (g % 16) * 64 + (t & 15) * 4));
    acc += v[0].x + v[1].y + v[2].z + v[3].w;
    unsigned long long t1 = gt() + (acc == 123.f);
    const int tile = (g / 16) % 8, n = tile * 128 + t;
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(out + (size_t)j * 1024 + n, acc);
    } else if (mode == 1) {  // plain stores of the partial (no reduction)
#pragma unroll
      for (int j = 0; j < 32; ++j) out[(size_t)(g * 32 + j) * 128 + t] = acc;
    }
    unsigned long long t2 = gt();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    target += G;
    if (t == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      unsigned vv;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(vv) : "l"(ctr) : "memory"); } while (vv < target);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    unsigned long long t3 = gt();
    if (t == 0) { unsigned long long* p = prof + ((size_t)g * nph + ph) * 4; p[0] = t0; p[1] = t1; p[2] = t2; p[3] = t3; }
  }
  if (t == 0) { sm[0] = 1.f; x[1 << 20] = acc; }
  // note: sm[0] must start at 0

}
int main() {
  float *x, *out; unsigned* ctr; unsigned long long* prof;
  const int nph = 64, G = 148;
  float* big; cudaMalloc(&big, (size_t)148 * (24u << 20)); cudaMemset(big, 0, (size_t)148 * (24u << 20)); cudaMalloc(&x, 8 << 20); cudaMalloc(&out, 64 << 20); cudaMalloc(&ctr, 4); cudaMalloc(&prof, sizeof(unsigned long long) * G * nph * 4);
  cudaMemset(x, 0, 8 << 20); cudaMemset(out, 0, 64 << 20);
  cudaFuncSetAttribute(phases, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int spin = 0; spin <= 3; spin += 3)
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(ctr, 0, 4);
    int nph_ = nph, spin_ = spin; float* xx = x; float* oo = out; unsigned* cc = ctr; unsigned long long* pp = prof; int mm = mode;
    const float* bb = big; void* args[] = {&xx, &oo, &cc, &pp, &nph_, &mm, &spin_, &bb};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)phases, dim3(G), dim3(320), args, 200 * 1024, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("error %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
    std::vector<unsigned long long> h((size_t)G * nph * 4);
    cudaMemcpy(h.data(), prof, h.size() * 8, cudaMemcpyDeviceToHost);
    double ld = 0, red = 0, bar = 0, tot = 0; int cnt = 0;
    for (int g = 0; g < G; ++g) for (int ph = 8; ph < nph; ++ph) {
      unsigned long long* p = &h[((size_t)g * nph + ph) * 4];
      ld += p[1] - p[0]; red += p[2] - p[1]; bar += p[3] - p[2]; tot += p[3] - p[0]; ++cnt;
    }
    const char* nm[] = {"32 scalar reds/thread (16-way contended)", "32 plain stores/thread", "no output"};
    printf("spin_warps %d  %-42s load %.2f us  out %.2f us  fence+barrier %.2f us  phase %.2f us\n", spin, nm[mode], ld / cnt / 1e3, red / cnt / 1e3, bar / cnt / 1e3, tot / cnt / 1e3);
  }
  return 0;
}
