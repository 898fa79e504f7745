// Microbenchmark: pace of back-to-back tcgen05.mma (kind::f16, bf16 operands, fp32 accumulator in TMEM, cta_group::1,
// M = 128, K = 16) as a function of N and of where the operands live -- the instruction mix of the fused decode step's
// linears (decode_mega.cu: weights = A operand from shared memory, <= 32 activation rows = B operand, N = 32).
// Question it answers: the MMA warp spends ~56 cycles per M128 x N32 x K16 instruction (nominal 16) and
// the time scales with N, not with the instruction count.  Is that the shared-memory operand read (SS form), i.e.
// would another operand assignment (A = activations in TMEM, B = 256 weight rows) move more weight bytes per cycle?
//   variants: SS form with N = 16 .. 256; the same with all instructions reading ONE k-slice (operand reuse);
//   reports cycles per instruction and weight bytes per cycle (A bytes for the SS/weights-as-A form, B bytes when
//   the weights are the B operand).
// Build / run: make -C tools/ubench mma_bench && tools/ubench/mma_bench
#include <cstdint>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../markushgrapher_b200/csrc/ptx.cuh"

using namespace mg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// one CTA per SM, one warp: issues `iters` rounds of 4 k-slices x `per_k` instructions against a [128 x 64] bf16 A tile
// and a [N x 64] bf16 B tile (both K-major, 128-byte swizzle, contents irrelevant), then commits and waits.
template <int N>
__global__ void __launch_bounds__(32, 1) mma_pace(int iters, int same_slice, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  uint8_t* a_tile = sm;                       // 128 rows x 128 B = 16 KB
  uint8_t* b_tile = sm + 16384;               // N rows x 128 B (<= 32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 32) reinterpret_cast<uint32_t*>(sm)[i] = 0x3f803f80u;  // bf16 1.0
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncwarp();
  tmem_alloc(slot, 256);
  tmem_relinquish();
  tc_fence_before();
  __syncwarp();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, N);
  const uint64_t da = make_sw128_kmajor_desc(smem_u32(a_tile)), db = make_sw128_kmajor_desc(smem_u32(b_tile));
  long long t0 = 0, t1 = 0;
  for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up
    t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int kk = same_slice ? 0 : k;
          umma_bf16(tmem, da + 2 * kk, db + 2 * kk, idesc, (it | k) ? 1u : 0u);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, rep & 1);
    t1 = clock64();
  }
  tc_fence_before();
  __syncwarp();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tmem_dealloc(tmem, 256);
}

template <int N>
int run(int ctas, int same_slice, long long* d_cyc) {
  const int iters = 512;
  CK(cudaFuncSetAttribute(mma_pace<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  mma_pace<N><<<ctas, 32, 16384 + 32768 + 2048>>>(iters, same_slice, d_cyc);
  CK(cudaDeviceSynchronize());
  long long h[148];
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < ctas; ++i) mean += (double)h[i];
  mean /= ctas;
  const double per = mean / (iters * 4.0);
  printf("N %3d  CTAs %3d  %s : %7.1f cycles / tcgen05.mma   A(weights) %6.1f B/cycle   B(weights) %6.1f B/cycle   nominal %5.1f\n", N,
         ctas, same_slice ? "one k-slice " : "four k-slices", per, 128 * 32.0 / per, N * 32.0 / per, 128.0 * N / 256.0);
  return 0;
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  for (int ctas : {1, 148})
    for (int same : {0, 1}) {
      if (run<16>(ctas, same, d)) return 1;
      if (run<32>(ctas, same, d)) return 1;
      if (run<64>(ctas, same, d)) return 1;
      if (run<128>(ctas, same, d)) return 1;
      if (run<256>(ctas, same, d)) return 1;
    }
  return 0;
}
