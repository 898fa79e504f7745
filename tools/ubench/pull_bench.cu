// Microbenchmark: how fast can ONE SM (and all 148 together) pull a read-once stream through a shared-memory ring of
// cp.async.bulk copies -- the access pattern of the fused decode step's attention phases (decode_mega.cu).
// Question it answers (DESIGN.md decision 9): the cross-attention phase streams at ~39-43 GB/s per SM and did not get
// faster when 23 % of its bytes were L2-resident.  Is that the SM-side bulk-copy pipeline, or DRAM?
//   for every (stages, chunk KB) x (source in DRAM | source L2-resident) x (CTAs = 1, 37, 74, 148):
//     one producer lane per CTA issues chunk after chunk into the ring, one consumer warp "consumes" a chunk by
//     reading 16 bytes per lane from it and releasing the slot; reports GB/s per SM and in total.
// Build / run: make -C tools/ubench pull_bench && tools/ubench/pull_bench
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_par(uint32_t bar, uint32_t par) {
  uint32_t done;
  do {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
  } while (!done);
}

// per CTA: bytes_per_cta bytes starting at src + cta * stride (wrapping inside `span` bytes so the source can be made
// L2-resident by choosing a small span)
__global__ void __launch_bounds__(64, 1) pull(const uint8_t* __restrict__ src, size_t span, size_t bytes_per_cta, int stages,
                                              uint32_t chunk, int evict_first, unsigned long long* t_out, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint8_t* ring = sm + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(empty + s)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t n_chunks = bytes_per_cta / chunk;
  const size_t base = ((size_t)blockIdx.x * bytes_per_cta) % span;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (warp == 0) {
    if (lane == 0) {
      uint64_t pol;
      if (evict_first)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      else
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
      int s = 0;
      uint32_t ph = 0;
      for (size_t c = 0; c < n_chunks; ++c) {
        wait_par(s32(empty + s), ph ^ 1);
        const size_t off = (base + c * chunk) % span;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + s)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                         s32(ring + (size_t)s * chunk)),
                     "l"(reinterpret_cast<uint64_t>(src + off)), "r"(chunk), "r"(s32(full + s)), "l"(pol)
                     : "memory");
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    int s = 0;
    uint32_t ph = 0;
    float acc = 0.f;
    for (size_t c = 0; c < n_chunks; ++c) {
      wait_par(s32(full + s), ph);
      acc += reinterpret_cast<const float4*>(ring + (size_t)s * chunk)[lane].x;  // touch the chunk
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + s)) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    if (acc == 1.2345e-30f) sink[0] = acc;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (lane == 0) { t_out[2 * blockIdx.x] = t0; t_out[2 * blockIdx.x + 1] = t1; }
  }
}

int main() {
  const size_t big = (size_t)6 << 30;   // DRAM-resident source (>> 126 MB L2)
  const size_t small = (size_t)48 << 20;  // L2-resident source
  uint8_t* buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  unsigned long long* t;
  float* sink;
  CK(cudaMalloc(&t, 2 * 148 * sizeof(unsigned long long)));
  CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(pull, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int cfgs[][2] = {{5, 40}, {4, 48}, {3, 64}, {6, 32}, {8, 24}, {12, 16}, {2, 96}, {5, 20}, {5, 10}};
  const int ctas[] = {1, 37, 74, 148};
  printf("%8s %8s %6s %6s | %10s %10s\n", "source", "ring", "chunk", "CTAs", "GB/s/SM", "GB/s total");
  for (int srcmode = 0; srcmode < 2; ++srcmode)
    for (auto& c : cfgs)
      for (int G : ctas) {
        const int stages = c[0];
        const uint32_t chunk = (uint32_t)c[1] * 1024;
        const size_t span = srcmode ? small : big;
        size_t per = ((size_t)24 << 20) / chunk * chunk;  // 24 MB per CTA
        const size_t smem = 1024 + (size_t)stages * chunk;
        if (srcmode) {  // warm L2 with the span
          pull<<<148, 64, smem>>>(buf, span, (span / 148) / chunk * chunk, stages, chunk, 0, t, sink);
        }
        pull<<<G, 64, smem>>>(buf, span, per, stages, chunk, srcmode ? 0 : 1, t, sink);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(2 * G);
        CK(cudaMemcpy(h.data(), t, 2 * G * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long lo = ~0ull, hi = 0;
        double per_sm = 0;
        for (int g = 0; g < G; ++g) {
          lo = h[2 * g] < lo ? h[2 * g] : lo;
          hi = h[2 * g + 1] > hi ? h[2 * g + 1] : hi;
          per_sm += (double)per / (double)(h[2 * g + 1] - h[2 * g]);
        }
        printf("%8s %5d st %4d K %6d | %10.1f %10.1f\n", srcmode ? "L2" : "DRAM", stages, c[1], G, per_sm / G,
               (double)per * G / (double)(hi - lo));
      }
  return 0;
}
