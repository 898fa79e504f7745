// Autoregressive decode-step kernels (T5/UDOP decoder, one new token per image per step):
// fused KV-cache append + self-attention with the T5 unidirectional bucket bias, cross-attention over the
// encoder memory, ReLU+split, LM-head argmax with finished-row bookkeeping and next-token embedding.
// Every kernel reads the current step from a device counter so one captured CUDA graph replays for all steps.
// Arithmetic follows transformers/models/udop/modeling_udop.py (UdopAttention.forward :531-622,
// compute_bias :514-529, UdopStack decoder path :1146-1256, lm head :1585-1590) and
// transformers/generation/utils.py::_sample (:2762-2805).
#include <algorithm>

#include <stdio.h>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

// =====================================================================================================
// Single-query attention for one (head, image).   HD = 64.
//   K^T cache : kt[b][h*64 + d][key]   (row stride kt_ld, image stride kt_bs)   -> coalesced over keys
//   V   cache : v [b][key][h*64 + d]   (row stride v_ld,  image stride v_bs)    -> coalesced over d
// SELF : n_keys = step+1; the new k/v row (from the fused QKV GEMM output qkv[b][3*D]) is appended to the
//        caches by this kernel and used from shared memory; bias = dec_bias[lut[step - j]][h].
// CROSS: n_keys = Mp; additive mask (1-mask)*finfo.min, no positional bias (:588-593).
// Scores are NOT scaled by 1/sqrt(d) (T5).  Output ctx[b][h*64+d] fp32 feeds the O projection.
template <bool SELF>
__global__ void __launch_bounds__(256) dec_attn_kernel(const float* __restrict__ qsrc, int q_ld, float* __restrict__ kt,
                                                       int64_t kt_ld, int64_t kt_bs, float* __restrict__ v,
                                                       int64_t v_ld, int64_t v_bs, const int* __restrict__ step_ptr,
                                                       int n_keys_cross, const int* __restrict__ mask, int mask_ld,
                                                       const float* __restrict__ dec_bias, const int* __restrict__ lut,
                                                       int H, int D, float* __restrict__ ctx, int kpad_keys) {
  constexpr int HD = 64;
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;             // [64]
  float* snew = sq + HD;      // [64] new v row (self)
  float* sred = snew + HD;    // [16*64] PV partials / reductions
  float* sc = sred + 16 * HD; // [n_keys padded to 4]
  __shared__ float s_bcast[2];
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  griddep_launch();
  griddep_wait();
  const int step = SELF ? *step_ptr : 0;
  const int nk = SELF ? step + 1 : n_keys_cross;
  const int ncache = SELF ? step : nk;  // keys already resident in the caches
  float* ktb = kt + (int64_t)b * kt_bs + (int64_t)h * HD * kt_ld;
  float* vb = v + (int64_t)b * v_bs + (int64_t)h * HD;

  if (tid < HD) {
    sq[tid] = qsrc[(int64_t)b * q_ld + h * HD + tid];
    if (SELF) {
      const float kn = qsrc[(int64_t)b * q_ld + D + h * HD + tid];
      const float vn = qsrc[(int64_t)b * q_ld + 2 * D + h * HD + tid];
      snew[tid] = vn;
      ktb[(int64_t)tid * kt_ld + step] = kn;  // append
      vb[(int64_t)step * v_ld + tid] = vn;
      sred[tid] = sq[tid] * kn;  // partial products of q . k_new
    }
  }
  __syncthreads();
  // ---- scores over cached keys.  Work item = (group of 4 consecutive keys, quarter of the 64 d-rows): at a few
  // hundred cached keys this keeps all 256 threads busy with ONE batch of 16 independent float4 loads each
  // (coalesced rows of K^T) instead of a quarter of the threads running four dependent batches.  The four partial
  // sums of a key are combined in the bias pass below.
  const int n4 = ncache >> 2;
  float* scp = sc + ((kpad_keys + 3) & ~3);  // [3][kpad] partials of d-quarters 1..3 (quarter 0 lives in sc)
  for (int u = tid; u < n4 * 4; u += 256) {
    const int dq = u / n4, g = u - dq * n4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* col = reinterpret_cast<const float4*>(ktb + (int64_t)(dq * 16) * kt_ld) + g;
    const int64_t ld4 = kt_ld >> 2;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float4 kk = __ldg(col + d * ld4);
      const float qd = sq[dq * 16 + d];
      a.x += qd * kk.x; a.y += qd * kk.y; a.z += qd * kk.z; a.w += qd * kk.w;
    }
    float* dst = dq == 0 ? sc : scp + (int64_t)(dq - 1) * kpad_keys;
    reinterpret_cast<float4*>(dst)[g] = a;
  }
  for (int j = (n4 << 2) + tid; j < ncache; j += 256) {  // ragged tail (< 4 keys)
    float a = 0.f;
#pragma unroll 16
    for (int d = 0; d < HD; ++d) a += sq[d] * ktb[(int64_t)d * kt_ld + j];
    sc[j] = a;
    scp[j] = 0.f;
    scp[kpad_keys + j] = 0.f;
    scp[2 * kpad_keys + j] = 0.f;
  }
  if (SELF && tid == 0) {
    float a = 0.f;
    for (int d = 0; d < HD; ++d) a += sred[d];
    sc[step] = a;
    scp[step] = 0.f;
    scp[kpad_keys + step] = 0.f;
    scp[2 * kpad_keys + step] = 0.f;
  }
  __syncthreads();
  // ---- bias / mask, max
  float mx = -INFINITY;
  for (int j = tid; j < nk; j += 256) {
    float s = (sc[j] + scp[j]) + (scp[kpad_keys + j] + scp[2 * kpad_keys + j]);
    if (SELF) {
      s += dec_bias[lut[step - j] * H + h];  // causal mask is all-visible for the newest token
    } else {
      s += (mask[(int64_t)b * mask_ld + j] ? 0.f : -3.4028234663852886e38f);
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float* wred = sred + 64;  // [8] scratch (sred[0..63] no longer needed)
  if ((tid & 31) == 0) wred[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = wred[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, wred[i]);
    s_bcast[0] = m;
  }
  __syncthreads();
  mx = s_bcast[0];
  float sum = 0.f;
  for (int j = tid; j < nk; j += 256) {
    const float p = expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if ((tid & 31) == 0) wred[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += wred[i];
    s_bcast[1] = s;
  }
  __syncthreads();
  const float inv = 1.f / s_bcast[1];
  // ---- P.V: thread (r = tid/16, c = tid%16) accumulates float4 column c over keys j == r (mod 16)
  const int r = tid >> 4, c = tid & 15;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* v4 = reinterpret_cast<const float4*>(vb) + c;
  const int64_t vld4 = v_ld >> 2;
#pragma unroll 16
  for (int j = r; j < ncache; j += 16) {
    const float4 vv = __ldg(v4 + (int64_t)j * vld4);
    const float p = sc[j];
    acc.x += p * vv.x; acc.y += p * vv.y; acc.z += p * vv.z; acc.w += p * vv.w;
  }
  if (SELF && r == 0) {
    const float p = sc[step];
    acc.x += p * snew[4 * c]; acc.y += p * snew[4 * c + 1]; acc.z += p * snew[4 * c + 2]; acc.w += p * snew[4 * c + 3];
  }
  __syncthreads();
  reinterpret_cast<float4*>(sred)[r * 16 + c] = acc;
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) o += sred[rr * HD + tid];
    ctx[(int64_t)b * D + h * HD + tid] = o * inv;
  }
}

static size_t dec_attn_smem(int max_keys) { return (size_t)(64 + 64 + 16 * 64 + 4 * (((max_keys + 3) / 4) * 4)) * sizeof(float); }

void launch_dec_self_attn(cudaStream_t st, const float* qkv, int B, int H, int D, float* kt, int64_t kt_ld,
                          int64_t kt_bs, float* v, int64_t v_ld, int64_t v_bs, const int* step_ptr, int max_keys,
                          const float* dec_bias, const int* lut, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  dim3 grid(H, B);
  launch_pdl(dec_attn_kernel<true>, grid, dim3(256), dec_attn_smem(max_keys), st, qkv, 3 * D, kt, kt_ld, kt_bs, v, v_ld,
             v_bs, step_ptr, 0, (const int*)nullptr, 0, dec_bias, lut, H, D, ctx, ((max_keys + 3) / 4) * 4);
}

void launch_dec_cross_attn(cudaStream_t st, const float* q, int B, int H, int D, const float* kt, int64_t kt_ld,
                           int64_t kt_bs, const float* v, int64_t v_ld, int64_t v_bs, int n_keys, const int* mask,
                           int mask_ld, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  MG_REQUIRE(n_keys % 4 == 0 && kt_ld % 4 == 0, "cross-attention memory length must be a multiple of 4");
  dim3 grid(H, B);
  dec_attn_kernel<false><<<grid, 256, dec_attn_smem(n_keys), st>>>(q, D, const_cast<float*>(kt), kt_ld, kt_bs,
                                                                   const_cast<float*>(v), v_ld, v_bs, nullptr, n_keys,
                                                                   mask, mask_ld, nullptr, nullptr, H, D, ctx,
                                                                   ((n_keys + 3) / 4) * 4);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Cross-attention for one (head, image) as a bulk-copy streaming pipeline (the dominant, HBM-bound kernel of
// the whole path: it reads the image's cross K and V of every layer once per generated token).
//   K^T : kt[b][h][64][Mp]   one contiguous 64*Mp block   -> chunks of RK whole d-rows
//   V   : v [b][h][Mp][64]   one contiguous Mp*64 block   -> chunks of 64 keys
// A producer thread streams both blocks through a ring of shared-memory stages with cp.async.bulk, completing
// on "full" mbarriers; 8 consumer warps drain a stage and release it through an "empty" mbarrier, so the
// deep queue of bulk copies (not warp occupancy) keeps HBM busy and the V stream starts while the softmax of
// the scores is still being reduced.  Scores: thread = key (conflict-free smem reads), fp32 accumulate over d;
// additive mask (1-mask)*finfo.min (modeling_udop.py:1202-1205), no positional bias, no 1/sqrt(d) scale.
constexpr int CA_STAGE_BYTES = 15360;  // 3 stages + scores: 4 CTAs (one wave of 592 slots) per SM at Mp ~ 1232
constexpr int CA_NST = 3;
constexpr int CA_MAXK = 8;  // keys per consumer thread: Mp <= 2048

__global__ void __launch_bounds__(288) cross_attn_stream_kernel(const float* __restrict__ q, const float* __restrict__ kt,
                                                                const float* __restrict__ v,
                                                                const int* __restrict__ mask, int Mp, int H, int D,
                                                                float* __restrict__ ctx) {
  constexpr int HD = 64;
  extern __shared__ __align__(128) uint8_t smc[];
  float* ring = reinterpret_cast<float*>(smc);                                   // [NST][4096 floats]
  float* sc = ring + CA_NST * (CA_STAGE_BYTES / 4);                              // [Mp]
  float* sq = sc + Mp;                                                           // [64]
  float* sred = sq + HD;                                                         // [16*64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sred + 16 * HD);              // [NST]
  uint64_t* empty_bar = full_bar + CA_NST;                                       // [NST]
  float* s_b = reinterpret_cast<float*>(empty_bar + CA_NST);                     // [2 + 8]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* ktb = kt + ((int64_t)b * H + h) * HD * Mp;
  const float* vb = v + ((int64_t)b * H + h) * (int64_t)Mp * HD;
  const int RK = min(HD, CA_STAGE_BYTES / (Mp * 4));  // d-rows per K chunk
  const int nkc = (HD + RK - 1) / RK;
  constexpr int VR = CA_STAGE_BYTES / (HD * 4);  // keys per V chunk
  const int nvc = (Mp + VR - 1) / VR;

  griddep_launch();
  if (tid == 0) {
    for (int s = 0; s < CA_NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // The cross K/V blocks were written once before the decode loop, so the producer starts streaming them right
  // away; only the consumers depend on the preceding kernel (q) and wait for it.
  if (warp != 8) {
    griddep_wait();
    if (tid < HD) sq[tid] = q[(int64_t)b * D + h * HD + tid];
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

  if (warp == 8) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int c = 0; c < nkc + nvc; ++c) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        const float* src;
        uint32_t bytes;
        if (c < nkc) {
          const int r0 = c * RK, rows = min(RK, HD - r0);
          src = ktb + (int64_t)r0 * Mp;
          bytes = (uint32_t)rows * Mp * 4;
        } else {
          const int m0 = (c - nkc) * VR, rows = min(VR, Mp - m0);
          src = vb + (int64_t)m0 * HD;
          bytes = (uint32_t)rows * HD * 4;
        }
        mbar_expect_tx(&full_bar[s], bytes);
        bulk_load_1d(ring + s * (CA_STAGE_BYTES / 4), src, bytes, &full_bar[s]);
        if (++s == CA_NST) { s = 0; ph ^= 1; }
      }
    }
    griddep_wait();  // every thread of every kernel in the chain waits once: keeps completion order transitive
    return;
  }
  // -------------------------------------------------------------------- consumers (256 threads)
  int s = 0;
  uint32_t ph = 0;
  float acc[CA_MAXK];
#pragma unroll
  for (int i = 0; i < CA_MAXK; ++i) acc[i] = 0.f;
  for (int c = 0; c < nkc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const float* buf = ring + s * (CA_STAGE_BYTES / 4);
    const int r0 = c * RK, rows = min(RK, HD - r0);
    for (int rr = 0; rr < rows; ++rr) {
      const float qd = sq[r0 + rr];
      const float* row = buf + rr * Mp;
#pragma unroll
      for (int i = 0; i < CA_MAXK; ++i) {
        const int m = tid + 256 * i;
        if (m < Mp) acc[i] += qd * row[m];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == CA_NST) { s = 0; ph ^= 1; }
  }
  // mask, max
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < CA_MAXK; ++i) {
    const int m = tid + 256 * i;
    if (m < Mp) {
      acc[i] += (mask[(int64_t)b * Mp + m] ? 0.f : -3.4028234663852886e38f);
      mx = fmaxf(mx, acc[i]);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_b[2 + warp] = mx;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  mx = s_b[2];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_b[2 + w]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < CA_MAXK; ++i) {
    const int m = tid + 256 * i;
    if (m < Mp) {
      const float p = expf(acc[i] - mx);
      sc[m] = p;
      sum += p;
    }
  }
  sum = warp_sum(sum);
  asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read s_b[2..9]
  if (lane == 0) s_b[2 + warp] = sum;
  asm volatile("bar.sync 1, 256;" ::: "memory");  // sc[] and partial sums visible
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += s_b[2 + w];
  const float inv = 1.f / sum;
  // P.V over the V chunks: thread (r = tid/16, c = tid%16) -> float4 column c, keys j == r (mod 16)
  const int r = tid >> 4, cc = tid & 15;
  float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nvc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const float4* buf4 = reinterpret_cast<const float4*>(ring + s * (CA_STAGE_BYTES / 4));
    const int m0 = c * VR, rows = min(VR, Mp - m0);
#pragma unroll
    for (int j = 0; j < (VR + 15) / 16; ++j) {
      const int jj = r + 16 * j;
      if (jj < rows) {
        const float4 vv = buf4[jj * 16 + cc];
        const float p = sc[m0 + jj];
        a4.x += p * vv.x; a4.y += p * vv.y; a4.z += p * vv.z; a4.w += p * vv.w;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == CA_NST) { s = 0; ph ^= 1; }
  }
  reinterpret_cast<float4*>(sred)[r * 16 + cc] = a4;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) o += sred[rr * HD + tid];
    ctx[(int64_t)b * D + h * HD + tid] = o * inv;
  }
}

void launch_cross_attn_stream(cudaStream_t st, const float* q, int B, int H, int D, const float* kt, const float* v,
                              int Mp, const int* mask, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  MG_REQUIRE(Mp % 4 == 0 && Mp <= 256 * CA_MAXK, "cross-attention memory length must be a multiple of 4, <= 2048");
  const size_t smem = (size_t)CA_NST * CA_STAGE_BYTES + (size_t)(Mp + 64 + 16 * 64) * 4 + 2 * CA_NST * 8 + 64;
  static bool attr = false;
  if (!attr) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  dim3 grid(H, B);
  launch_pdl(cross_attn_stream_kernel, grid, dim3(288), smem, st, q, kt, v, mask, Mp, H, D, ctx);
}

// =====================================================================================================
// "kv24" cross K/V: the decode loop re-reads the image's cross K and V of every layer once per generated token,
// which is the bulk of the step's HBM traffic.  Both are stored with 24 significant bits per element -- fp32
// rounded to nearest-even to 15 stored mantissa bits (relative error <= 2^-16, the same order as the split-bf16 GEMM
// operands that produced them) -- as a 16-bit plane (sign, exponent, 7 mantissa bits = the bf16 truncation) plus an 8-bit
// plane (the next 8 mantissa bits).  3 bytes instead of 4: the dominant kernel moves 25 % fewer bytes.
// Per (image, head) block of 384*Mp bytes, n = 64*Mp:
//   [K^T hi: 64 x Mp u16][K^T lo: 64 x Mp u8][V hi: Mp x 64 u16][V lo: Mp x 64 u8]
__device__ __forceinline__ uint32_t f32_to_kv24(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x7Fu + ((u >> 8) & 1u);  // round to nearest even at bit 8
  return u >> 8;
}
__global__ void kv24_pack_kernel(const float* __restrict__ kt, const float* __restrict__ v, int64_t n, int64_t n_bh,
                                 uint8_t* __restrict__ out) {
  const int64_t total4 = n_bh * n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bh = (i * 4) / n, e = (i * 4) % n;
    uint8_t* blk = out + bh * 6 * n;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float4 x = *reinterpret_cast<const float4*>((which ? v : kt) + i * 4);
      const uint32_t t0 = f32_to_kv24(x.x), t1 = f32_to_kv24(x.y), t2 = f32_to_kv24(x.z), t3 = f32_to_kv24(x.w);
      uint2 hi;
      hi.x = (t0 >> 8) | ((t1 >> 8) << 16);
      hi.y = (t2 >> 8) | ((t3 >> 8) << 16);
      const uint32_t lo = (t0 & 0xffu) | ((t1 & 0xffu) << 8) | ((t2 & 0xffu) << 16) | ((t3 & 0xffu) << 24);
      uint8_t* base = blk + (which ? 3 * n : 0);
      *reinterpret_cast<uint2*>(base + e * 2) = hi;
      *reinterpret_cast<uint32_t*>(base + 2 * n + e) = lo;
    }
  }
}
void launch_kv24_pack(cudaStream_t st, const float* kt, const float* v, int B, int H, int Mp, uint8_t* out) {
  MG_REQUIRE(Mp % 8 == 0, "kv24: memory length must be a multiple of 8");
  const int64_t n = (int64_t)64 * Mp, n_bh = (int64_t)B * H;
  const int blocks = (int)std::min<int64_t>((n_bh * n / 4 + 255) / 256, 148 * 16);
  kv24_pack_kernel<<<blocks, 256, 0, st>>>(kt, v, n, n_bh, out);
  MG_CHECK_CUDA(cudaGetLastError());
}

// fp32 -> kv24 block -> fp32 (parity tests of the storage format): x holds n_bh blocks of n = 64*Mp floats, laid out
// like one layer's cross K^T; `packed` (6*n bytes per block) is scratch.
__global__ void kv24_unpack_k_kernel(const uint8_t* __restrict__ packed, int64_t n, int64_t n_bh, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_bh * n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bh = i / n, e = i % n;
    const uint8_t* blk = packed + bh * 6 * n;
    const uint32_t hi = reinterpret_cast<const uint16_t*>(blk)[e], lo = blk[2 * n + e];
    const uint32_t hv = reinterpret_cast<const uint16_t*>(blk + 3 * n)[e], lv = blk[5 * n + e];
    out[i] = __uint_as_float((hi << 16) | (lo << 8));
    out[n_bh * n + i] = __uint_as_float((hv << 16) | (lv << 8));
  }
}
void launch_kv24_roundtrip(cudaStream_t st, const float* kt, const float* v, int B, int H, int Mp, uint8_t* packed,
                           float* out) {
  launch_kv24_pack(st, kt, v, B, H, Mp, packed);
  const int64_t n = (int64_t)64 * Mp, n_bh = (int64_t)B * H;
  const int blocks = (int)std::min<int64_t>((n_bh * n + 255) / 256, 148 * 16);
  kv24_unpack_k_kernel<<<blocks, 256, 0, st>>>(packed, n, n_bh, out);
  MG_CHECK_CUDA(cudaGetLastError());
}

// two adjacent kv24 elements from a 32-bit word of the hi plane and a (zero-extended) 16-bit word of the lo plane
__device__ __forceinline__ float kv24_lo_elem(uint32_t hi2, uint32_t lo2) { return __uint_as_float(__byte_perm(hi2, lo2, 0x1046)); }
__device__ __forceinline__ float kv24_hi_elem(uint32_t hi2, uint32_t lo2) { return __uint_as_float(__byte_perm(hi2, lo2, 0x3256)); }

// Streaming cross-attention over kv24 blocks: same pipeline as cross_attn_stream_kernel (producer thread + ring of
// bulk copies + 8 consumer warps); a chunk is two bulk copies (hi rows, lo rows) landing back to back in one stage.
// Scores: thread = PAIR of adjacent keys (one 32-bit + one 16-bit shared load per pair and d-row).
__global__ void __launch_bounds__(288) cross_attn_stream24_kernel(const float* __restrict__ q, const uint8_t* __restrict__ kv,
                                                                  const int* __restrict__ mask, int Mp, int H, int D,
                                                                  float* __restrict__ ctx) {
  constexpr int HD = 64;
  extern __shared__ __align__(128) uint8_t smc[];
  uint8_t* ring = smc;                                                           // [NST][CA_STAGE_BYTES]
  float* sc = reinterpret_cast<float*>(smc + CA_NST * CA_STAGE_BYTES);           // [Mp]
  float* sq = sc + Mp;                                                           // [64]
  float* sred = sq + HD;                                                         // [16*64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sred + 16 * HD);              // [NST]
  uint64_t* empty_bar = full_bar + CA_NST;                                       // [NST]
  float* s_b = reinterpret_cast<float*>(empty_bar + CA_NST);                     // [2 + 8]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n = (int64_t)HD * Mp;
  const uint8_t* blk = kv + ((int64_t)b * H + h) * 6 * n;
  const int RK = min(HD, (CA_STAGE_BYTES / (Mp * 3)) & ~1);  // d-rows per K chunk (even: both copies stay 16-byte sized)
  const int nkc = (HD + RK - 1) / RK;
  constexpr int VR = CA_STAGE_BYTES / (HD * 3);  // keys per V chunk
  const int nvc = (Mp + VR - 1) / VR;

  griddep_launch();
  if (tid == 0) {
    for (int s = 0; s < CA_NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp != 8) {
    griddep_wait();
    if (tid < HD) sq[tid] = q[(int64_t)b * D + h * HD + tid];
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

  if (warp == 8) {
    // ------------------------------------------------------------------ producer (K/V were written before the loop)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int c = 0; c < nkc + nvc; ++c) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint8_t *src_hi, *src_lo;
        uint32_t bytes_lo;
        if (c < nkc) {
          const int r0 = c * RK, rows = min(RK, HD - r0);
          src_hi = blk + (int64_t)r0 * Mp * 2;
          src_lo = blk + 2 * n + (int64_t)r0 * Mp;
          bytes_lo = (uint32_t)rows * Mp;
        } else {
          const int m0 = (c - nkc) * VR, rows = min(VR, Mp - m0);
          src_hi = blk + 3 * n + (int64_t)m0 * HD * 2;
          src_lo = blk + 5 * n + (int64_t)m0 * HD;
          bytes_lo = (uint32_t)rows * HD;
        }
        uint8_t* dst = ring + s * CA_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], 3 * bytes_lo);
        bulk_load_1d(dst, src_hi, 2 * bytes_lo, &full_bar[s]);
        bulk_load_1d(dst + 2 * bytes_lo, src_lo, bytes_lo, &full_bar[s]);
        if (++s == CA_NST) { s = 0; ph ^= 1; }
      }
    }
    griddep_wait();
    return;
  }
  // -------------------------------------------------------------------- consumers (256 threads)
  int s = 0;
  uint32_t ph = 0;
  const int npair = Mp >> 1;
  float acc[8];  // acc[2i], acc[2i+1] = keys 2*(tid + 256 i), +1
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int c = 0; c < nkc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const uint8_t* buf = ring + s * CA_STAGE_BYTES;
    const int r0 = c * RK, rows = min(RK, HD - r0);
    const uint8_t* lo_base = buf + (size_t)rows * Mp * 2;
    for (int rr = 0; rr < rows; ++rr) {
      const float qd = sq[r0 + rr];
      const uint32_t* hrow = reinterpret_cast<const uint32_t*>(buf + (size_t)rr * Mp * 2);
      const uint16_t* lrow = reinterpret_cast<const uint16_t*>(lo_base + (size_t)rr * Mp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pi = tid + 256 * i;
        if (pi < npair) {
          const uint32_t h2 = hrow[pi], l2 = lrow[pi];
          acc[2 * i] += qd * kv24_lo_elem(h2, l2);
          acc[2 * i + 1] += qd * kv24_hi_elem(h2, l2);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == CA_NST) { s = 0; ph ^= 1; }
  }
  // mask, max
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = 2 * (tid + 256 * (i >> 1)) + (i & 1);
    if (m < Mp) {
      acc[i] += (mask[(int64_t)b * Mp + m] ? 0.f : -3.4028234663852886e38f);
      mx = fmaxf(mx, acc[i]);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_b[2 + warp] = mx;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  mx = s_b[2];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_b[2 + w]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = 2 * (tid + 256 * (i >> 1)) + (i & 1);
    if (m < Mp) {
      const float pe = expf(acc[i] - mx);
      sc[m] = pe;
      sum += pe;
    }
  }
  sum = warp_sum(sum);
  asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read s_b[2..9]
  if (lane == 0) s_b[2 + warp] = sum;
  asm volatile("bar.sync 1, 256;" ::: "memory");  // sc[] and partial sums visible
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += s_b[2 + w];
  const float inv = 1.f / sum;
  // P.V: thread (r = tid/16, c = tid%16) -> 4 adjacent d of keys j == r (mod 16)
  const int r = tid >> 4, cc = tid & 15;
  float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nvc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const uint8_t* buf = ring + s * CA_STAGE_BYTES;
    const int m0 = c * VR, rows = min(VR, Mp - m0);
    const uint8_t* lo_base = buf + (size_t)rows * HD * 2;
#pragma unroll
    for (int j = 0; j < (VR + 15) / 16; ++j) {
      const int jj = r + 16 * j;
      if (jj < rows) {
        const uint2 h4 = *reinterpret_cast<const uint2*>(buf + ((size_t)jj * HD + 4 * cc) * 2);
        const uint32_t l4 = *reinterpret_cast<const uint32_t*>(lo_base + (size_t)jj * HD + 4 * cc);
        const float pj = sc[m0 + jj];
        a4.x += pj * __uint_as_float((h4.x << 16) | ((l4 & 0xffu) << 8));
        a4.y += pj * __uint_as_float((h4.x & 0xffff0000u) | (l4 & 0xff00u));
        a4.z += pj * __uint_as_float((h4.y << 16) | ((l4 >> 8) & 0xff00u));
        a4.w += pj * __uint_as_float((h4.y & 0xffff0000u) | ((l4 >> 16) & 0xff00u));
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == CA_NST) { s = 0; ph ^= 1; }
  }
  reinterpret_cast<float4*>(sred)[r * 16 + cc] = a4;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) o += sred[rr * HD + tid];
    ctx[(int64_t)b * D + h * HD + tid] = o * inv;
  }
}

void launch_cross_attn_stream24(cudaStream_t st, const float* q, int B, int H, int D, const uint8_t* kv, int Mp,
                                const int* mask, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  MG_REQUIRE(Mp % 8 == 0 && Mp <= 2048, "cross-attention memory length must be a multiple of 8, <= 2048");
  const size_t smem = (size_t)CA_NST * CA_STAGE_BYTES + (size_t)(Mp + 64 + 16 * 64) * 4 + 2 * CA_NST * 8 + 64;
  static bool attr = false;
  if (!attr) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_stream24_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  dim3 grid(H, B);
  launch_pdl(cross_attn_stream24_kernel, grid, dim3(288), smem, st, q, kv, mask, Mp, H, D, ctx);
}

// =====================================================================================================
// h = relu(x) -> split planes (decode FF: the wi GEMM is split-K/atomic so the activation cannot live in its epilogue)
__global__ void relu_split_kernel(const float* __restrict__ x, int64_t n, bf16* hi, bf16* lo) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    bf16 h, l;
    split_bf16(fmaxf(x[i], 0.f), h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}
void launch_relu_split(cudaStream_t st, const float* x, int64_t n, Planes out) {
  if (!n) return;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  relu_split_kernel<<<blocks, 256, 0, st>>>(x, n, out.hi, out.lo);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Greedy selection (GenerationMixin._sample :2762-2805): first-max argmax over fp32 logits, finished rows
// emit pad, EOS bookkeeping, token append, next-step input embedding.  One CTA per image.  The last CTA to
// finish advances the device step counter and publishes the "all rows finished" flag.
// Ordering of torch.argmax: the first maximum wins and NaN compares greater than every number (a row of NaN / -inf
// logits still yields a valid index, never an out-of-range embedding gather).
__device__ __forceinline__ bool argmax_better(float x, int xi, float best, int bi) {
  const bool xn = x != x, bn = best != best;
  if (xn || bn) return xn && (!bn || xi < bi);
  return x > best || (x == best && xi < bi);
}
// consumer half of the peer exchange: wait until every rank has published decode step `step`, then scatter the
// world x B tokens into column step + 1 of the global id matrix and keep the global finished flags / unfinished count
__device__ void peer_consume(const PeerExchange& px, int step) {
  const int* mine = px.peers[px.rank];
  const unsigned want = px.base + (unsigned)step + 1u;
  const int ring = step % PX_RING;
  if ((int)threadIdx.x < px.world) {
    const unsigned* f = reinterpret_cast<const unsigned*>(mine) + ring * PX_MAXW + threadIdx.x;
    unsigned v;
    long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - want) >= 0) break;
      if (clock64() - t0 > 20000000000LL) {  // a dead peer must not hang this GPU for ever
        printf("peer exchange: rank %d never saw step %d of rank %d (flag %u, want %u)\n", px.rank, step, (int)threadIdx.x, v, want);
        __trap();
      }
    }
  }
  __syncthreads();
  const int* slots = mine + PX_RING * PX_MAXW + (size_t)ring * px.world * px.bcap;
  for (int i = threadIdx.x; i < px.world * px.B; i += blockDim.x) {
    const int r = i / px.B, b = i - r * px.B;
    const int tok = __ldcv(slots + (size_t)r * px.bcap + b);
    px.all_ids[(int64_t)i * px.ld + step + 1] = tok;
    if (!px.gfinished[i] && tok == px.eos) {
      px.gfinished[i] = 1;
      atomicSub(px.g_unfinished, 1);
    }
  }
}
__global__ void peer_drain_kernel(PeerExchange px, int step) { peer_consume(px, step); }
void launch_peer_drain(cudaStream_t st, const PeerExchange& px, int step) {
  peer_drain_kernel<<<1, 256, 0, st>>>(px, step);
  MG_CHECK_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) greedy_select_kernel(const float* __restrict__ part_val,
                                                            const int* __restrict__ part_idx, int n_part,
                                                            const float* __restrict__ logits, int V, int64_t ld,
                                                            const float* __restrict__ emb, int D, int eos, int pad,
                                                            int64_t* __restrict__ out_ids, int out_ld,
                                                            int* __restrict__ finished, int* __restrict__ step_ptr,
                                                            int* __restrict__ n_unfinished, int* __restrict__ ticket,
                                                            float* __restrict__ x_next, float* __restrict__ logits_dump,
                                                            int64_t dump_bs, int64_t dump_ss,
                                                            const int64_t* __restrict__ forced, int forced_ld,
                                                            int* __restrict__ step_tok,
                                                            unsigned long long* __restrict__ step_ts,
                                                            const PeerExchange px) {
  const int b = blockIdx.x;
  griddep_launch();
  griddep_wait();
  const int step = *step_ptr;
  const int n_img = px.peers ? (int)gridDim.x - 1 : (int)gridDim.x;
  if (b == n_img) {
    // peer exchange: the extra CTA consumes the PREVIOUS step's tokens of all ranks.  It keeps its own launch counter
    // (*step_ptr is advanced by the last image CTA of this very launch and may already show the next step)
    const int k = *px.consumed;
    if (k > 0) peer_consume(px, k - 1);
    __syncthreads();
    if (threadIdx.x == 0) *px.consumed = k + 1;
    return;
  }
  const float* lg = logits + (int64_t)b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (part_val) {  // per-tile maxima were produced by the LM-head epilogue: reduce n_part candidates
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) {
      const float x = part_val[(int64_t)b * n_part + i];
      const int xi = part_idx[(int64_t)b * n_part + i];
      if (argmax_better(x, xi, best, bi)) {
        best = x;
        bi = xi;
      }
    }
    if (logits_dump)
      for (int i = threadIdx.x; i < V; i += blockDim.x)
        logits_dump[(int64_t)b * dump_bs + (int64_t)step * dump_ss + i] = lg[i];
  } else {
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float x = lg[i];
      if (logits_dump) logits_dump[(int64_t)b * dump_bs + (int64_t)step * dump_ss + i] = x;
      if (argmax_better(x, i, best, bi)) {
        best = x;
        bi = i;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (argmax_better(ob, oi, best, bi)) {
      best = ob;
      bi = oi;
    }
  }
  __shared__ float sb[8];
  __shared__ int si[8];
  __shared__ int s_tok;
  if ((threadIdx.x & 31) == 0) {
    sb[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (argmax_better(sb[w], si[w], best, bi)) {
        best = sb[w];
        bi = si[w];
      }
    bi = min(max(bi, 0), V - 1);
    const int fin = finished[b];
    int tok = fin ? pad : bi;
    out_ids[(int64_t)b * out_ld + step + 1] = tok;
    if (step_tok) step_tok[b] = tok;
    if (px.peers) {  // this image's token into slot (step % ring, my rank) of every rank's exchange buffer, over NVLink
      const size_t off = PX_RING * PX_MAXW + ((size_t)(step % PX_RING) * px.world + px.rank) * px.bcap + b;
      for (int r = 0; r < px.world; ++r) *reinterpret_cast<volatile int*>(px.peers[r] + off) = tok;
    }
    if (forced) {
      // teacher forcing (model(**batch).logits): the next decoder input is given, nothing ever "finishes"
      tok = (int)forced[(int64_t)b * forced_ld + min(step + 1, forced_ld - 1)];
    } else if (!fin && tok == eos) {
      finished[b] = 1;
      atomicSub(n_unfinished, 1);
    }
    s_tok = tok;
  }
  __syncthreads();
  const int tok = s_tok;
  const float4* e4 = reinterpret_cast<const float4*>(emb + (int64_t)tok * D);
  float4* x4 = reinterpret_cast<float4*>(x_next + (int64_t)b * D);
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) x4[c] = e4[c];
  __syncthreads();
  if (threadIdx.x == 0) {
    if (px.peers) __threadfence_system(); else __threadfence();
    const int t = atomicAdd(ticket, 1);
    if (t == n_img - 1) {
      if (px.peers) {  // all images of this rank are stored everywhere: raise this rank's flag on every rank
        __threadfence_system();
        const unsigned val = px.base + (unsigned)step + 1u;
        for (int r = 0; r < px.world; ++r) {
          unsigned* f = reinterpret_cast<unsigned*>(px.peers[r]) + (step % PX_RING) * PX_MAXW + px.rank;
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(val) : "memory");
        }
      }
      *ticket = 0;
      *step_ptr = step + 1;
      if (step_ts) {  // %globaltimer at the end of every decode step: true per-step latencies (mg_last_decode_p50)
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        step_ts[step] = now;
      }
    }
  }
}

void launch_greedy_select(cudaStream_t st, const float* part_val, const int* part_idx, int n_part, const float* logits,
                          int B, int V, int64_t ld, const float* emb, int D,
                          int eos, int pad, int64_t* out_ids, int out_ld, int* finished, int* step_ptr,
                          int* n_unfinished, int* ticket, float* x_next, float* logits_dump, int64_t dump_bs,
                          int64_t dump_ss, const int64_t* forced, int forced_ld, int* step_tok,
                          unsigned long long* step_ts, const PeerExchange* px) {
  const PeerExchange pe = px ? *px : PeerExchange{};
  launch_pdl(greedy_select_kernel, dim3(B + (pe.peers ? 1 : 0)), dim3(256), (size_t)0, st, part_val, part_idx, n_part, logits,
             V, ld, emb, D, eos, pad, out_ids, out_ld,
             finished, step_ptr, n_unfinished, ticket, x_next, logits_dump, dump_bs, dump_ss, forced, forced_ld,
             step_tok, step_ts, pe);
}

// decode state reset: ids[:,0] = start token, x = emb[start], finished = 0, step = 0
__global__ void decode_init_kernel(const float* __restrict__ emb, int D, int start, int pad, int B, int64_t* out_ids,
                                   int out_ld, int* finished, int* step_ptr, int* n_unfinished, int* ticket, float* x,
                                   const int64_t* __restrict__ forced, int forced_ld) {
  const int b = blockIdx.x;
  if (forced) start = (int)forced[(int64_t)b * forced_ld];
  for (int i = threadIdx.x; i < out_ld; i += blockDim.x) out_ids[(int64_t)b * out_ld + i] = (i == 0) ? start : pad;
  for (int c = threadIdx.x; c < D; c += blockDim.x) x[(int64_t)b * D + c] = emb[(int64_t)start * D + c];
  if (threadIdx.x == 0) {
    finished[b] = 0;
    if (b == 0) {
      *step_ptr = 0;
      *n_unfinished = B;
      *ticket = 0;
    }
  }
}
void launch_decode_init(cudaStream_t st, const float* emb, int D, int start, int pad, int B, int64_t* out_ids,
                        int out_ld, int* finished, int* step_ptr, int* n_unfinished, int* ticket, float* x,
                        const int64_t* forced, int forced_ld) {
  decode_init_kernel<<<B, 256, 0, st>>>(emb, D, start, pad, B, out_ids, out_ld, finished, step_ptr, n_unfinished, ticket, x,
                                        forced, forced_ld);
  MG_CHECK_CUDA(cudaGetLastError());
}

// per-row output length: index of the first EOS + 1, else ncols
__global__ void out_len_kernel(const int64_t* __restrict__ ids, int B, int ld, int ncols, int eos, int* __restrict__ len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = ncols;
  for (int t = 1; t < ncols; ++t)
    if (ids[(int64_t)b * ld + t] == eos) {
      n = t + 1;
      break;
    }
  len[b] = n;
}
void launch_out_len(cudaStream_t st, const int64_t* ids, int B, int ld, int ncols, int eos, int* len) {
  out_len_kernel<<<(B + 127) / 128, 128, 0, st>>>(ids, B, ld, ncols, eos, len);
  MG_CHECK_CUDA(cudaGetLastError());
}

// After the per-step all-gather: gathered[r][b] = token rank r emitted for its image b at column `col`.
// Writes the global id matrix and maintains the global finished flags / unfinished count (same stop rule on all ranks).
__global__ void scatter_step_kernel(const int* __restrict__ gathered, int n_rows, int col, int ld, int eos,
                                    int64_t* __restrict__ all_ids, int* __restrict__ gfinished,
                                    int* __restrict__ g_unfinished) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int tok = gathered[i];
  all_ids[(int64_t)i * ld + col] = tok;
  if (!gfinished[i] && tok == eos) {
    gfinished[i] = 1;
    atomicSub(g_unfinished, 1);
  }
}
void launch_scatter_step(cudaStream_t st, const int* gathered, int n_rows, int col, int ld, int eos, int64_t* all_ids,
                         int* gfinished, int* g_unfinished) {
  scatter_step_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(gathered, n_rows, col, ld, eos, all_ids, gfinished,
                                                          g_unfinished);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
