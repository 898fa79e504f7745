// Beam-search decode kernels.  Semantics follow transformers/generation/utils.py::_beam_search (:3076-3370) and
// its helpers (_get_top_k_continuations :2945-3000, _get_running_beams_for_next_iteration :3002-3022,
// _update_finished_beams :3024-3073, _check_early_stop_heuristic :2862-2917,
// _beam_search_has_unfinished_sequences :2919-2943) with the defaults the reference call site leaves in place
// (utils_evaluation.py:279-285 passes only num_beams / max_length): length_penalty 1.0, early_stopping False,
// one EOS id, num_return_sequences 1, decoder prompt length 1.
//
// B200 design: the nb beams of an image are nb rows of the decoder batch (row = image*nb + beam).
//  * cross-attention K/V are per IMAGE: one CTA streams an (image, head) block once and serves all nb queries
//    (the reference repeats the encoder output nb times, _expand_inputs_for_generation :866);
//  * the self-attention cache is never reordered: each logical row keeps an ancestry table anc[row][pos] = the
//    physical row whose cache slot holds position pos (the reference copies the whole cache by beam index every
//    step, DynamicCache.reorder_cache);
//  * selection (log-softmax, + running score, top-2nb over nb*V, finished-beam merge, early-stop heuristic)
//    runs in one CTA per image.
#include <algorithm>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

constexpr int BM_MAXNB = 8;    // beams per image supported
constexpr float BM_NEG = -1.0e9f;

// =====================================================================================================
// self-attention for one (head, logical row) with ancestry-indirected cache reads; appends this step's k/v at
// (physical row = logical row, pos = step).  Same math as dec_attn_kernel<SELF> (T5 bucket bias, no scale).
__global__ void __launch_bounds__(256) beam_self_attn_kernel(const float* __restrict__ qkv, float* __restrict__ kt,
                                                             int64_t kt_ld, int64_t kt_bs, float* __restrict__ v,
                                                             int64_t v_ld, int64_t v_bs,
                                                             const int* __restrict__ step_ptr,
                                                             const int* __restrict__ anc_sel, const int* anc0,
                                                             const int* anc1, int anc_ld,
                                                             const float* __restrict__ dec_bias,
                                                             const int* __restrict__ lut, int H, int D,
                                                             float* __restrict__ ctx) {
  constexpr int HD = 64;
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;
  float* snew = sq + HD;
  float* sred = snew + HD;     // [16*64]
  float* sc = sred + 16 * HD;  // [keys]
  int* sphys = reinterpret_cast<int*>(sc + anc_ld);  // [keys] physical rows
  __shared__ float s_bcast[2];
  const int h = blockIdx.x, r = blockIdx.y;
  const int tid = threadIdx.x;
  griddep_launch();
  griddep_wait();
  const int step = *step_ptr;
  const int* anc = (*anc_sel ? anc1 : anc0) + (int64_t)r * anc_ld;
  const int nk = step + 1;
  float* kt_self = kt + (int64_t)r * kt_bs + (int64_t)h * HD * kt_ld;
  float* v_self = v + (int64_t)r * v_bs + (int64_t)h * HD;
  if (tid < HD) {
    const float qv = qkv[(int64_t)r * 3 * D + h * HD + tid];
    const float kn = qkv[(int64_t)r * 3 * D + D + h * HD + tid];
    const float vn = qkv[(int64_t)r * 3 * D + 2 * D + h * HD + tid];
    sq[tid] = qv;
    snew[tid] = vn;
    kt_self[(int64_t)tid * kt_ld + step] = kn;
    v_self[(int64_t)step * v_ld + tid] = vn;
    sred[tid] = qv * kn;
  }
  for (int j = tid; j < step; j += 256) sphys[j] = anc[j];
  __syncthreads();
  for (int j = tid; j < step; j += 256) {
    const float* kp = kt + (int64_t)sphys[j] * kt_bs + (int64_t)h * HD * kt_ld + j;
    float a = 0.f;
#pragma unroll 16
    for (int d = 0; d < HD; ++d) a += sq[d] * kp[(int64_t)d * kt_ld];
    sc[j] = a;
  }
  if (tid == 0) {
    float a = 0.f;
    for (int d = 0; d < HD; ++d) a += sred[d];
    sc[step] = a;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < nk; j += 256) {
    const float s = sc[j] + dec_bias[lut[step - j] * H + h];
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float* wred = sred + 64;
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
  const int rr = tid >> 4, c = tid & 15;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = rr; j < step; j += 16) {
    const float4 vv = *reinterpret_cast<const float4*>(v + (int64_t)sphys[j] * v_bs + (int64_t)j * v_ld + h * HD + c * 4);
    const float p = sc[j];
    acc.x += p * vv.x; acc.y += p * vv.y; acc.z += p * vv.z; acc.w += p * vv.w;
  }
  if (rr == 0) {
    const float p = sc[step];
    acc.x += p * snew[4 * c]; acc.y += p * snew[4 * c + 1]; acc.z += p * snew[4 * c + 2]; acc.w += p * snew[4 * c + 3];
  }
  __syncthreads();
  reinterpret_cast<float4*>(sred)[rr * 16 + c] = acc;
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) o += sred[q * HD + tid];
    ctx[(int64_t)r * D + h * HD + tid] = o * inv;
  }
}

void launch_beam_self_attn(cudaStream_t st, const float* qkv, int R, int H, int D, float* kt, int64_t kt_ld,
                           int64_t kt_bs, float* v, int64_t v_ld, int64_t v_bs, const int* step_ptr,
                           const int* anc_sel, const int* anc0, const int* anc1, int anc_ld, const float* dec_bias,
                           const int* lut, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  const size_t smem = (size_t)(64 + 64 + 16 * 64 + 2 * anc_ld) * sizeof(float);
  dim3 grid(H, R);
  launch_pdl(beam_self_attn_kernel, grid, dim3(256), smem, st, qkv, kt, kt_ld, kt_bs, v, v_ld, v_bs, step_ptr, anc_sel,
             anc0, anc1, anc_ld, dec_bias, lut, H, D, ctx);
}

// =====================================================================================================
// streaming cross-attention, one CTA per (head, IMAGE) serving nq <= 8 queries (the image's beams) from a single
// pass over the K^T / V blocks.  Same pipeline as cross_attn_stream_kernel (decode.cu).
constexpr int BC_STAGE_BYTES = 24576;  // (15360 until round 2: fewer, larger chunks -- the per-chunk hand-off is what a streaming attention kernel pays for)
constexpr int BC_NST = 3;
constexpr int BC_MAXK = 8;

__global__ void __launch_bounds__(288) beam_cross_attn_kernel(const float* __restrict__ q, const float* __restrict__ kt,
                                                              const float* __restrict__ v,
                                                              const int* __restrict__ mask, int Mp, int H, int D,
                                                              int nq, float* __restrict__ ctx) {
  constexpr int HD = 64;
  extern __shared__ __align__(128) uint8_t smc[];
  float* ring = reinterpret_cast<float*>(smc);
  float* sc = ring + BC_NST * (BC_STAGE_BYTES / 4);             // [nq][Mp]
  float* sq = sc + (int64_t)nq * Mp;                            // [nq][64]
  float* sred = sq + nq * HD;                                   // [16*64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sred + 16 * HD);
  uint64_t* empty_bar = full_bar + BC_NST;
  float* s_b = reinterpret_cast<float*>(empty_bar + BC_NST);    // [8 warps][nq]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* ktb = kt + ((int64_t)b * H + h) * HD * Mp;
  const float* vb = v + ((int64_t)b * H + h) * (int64_t)Mp * HD;
  const int RK = min(HD, BC_STAGE_BYTES / (Mp * 4));
  const int nkc = (HD + RK - 1) / RK;
  constexpr int VR = BC_STAGE_BYTES / (HD * 4);
  const int nvc = (Mp + VR - 1) / VR;
  griddep_launch();
  if (tid == 0) {
    for (int s = 0; s < BC_NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 8) {
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
        bulk_load_1d(ring + s * (BC_STAGE_BYTES / 4), src, bytes, &full_bar[s]);
        if (++s == BC_NST) { s = 0; ph ^= 1; }
      }
    }
    griddep_wait();
    return;
  }
  griddep_wait();
  for (int i = tid; i < nq * HD; i += 256) sq[i] = q[((int64_t)b * nq + i / HD) * D + h * HD + (i % HD)];
  asm volatile("bar.sync 1, 256;" ::: "memory");
  int s = 0;
  uint32_t ph = 0;
  float acc[BM_MAXNB][BC_MAXK];
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k)
#pragma unroll
    for (int i = 0; i < BC_MAXK; ++i) acc[k][i] = 0.f;
  for (int c = 0; c < nkc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const float* buf = ring + s * (BC_STAGE_BYTES / 4);
    const int r0 = c * RK, rows = min(RK, HD - r0);
    for (int rr = 0; rr < rows; ++rr) {
      const float* row = buf + rr * Mp;
      float kv[BC_MAXK];
#pragma unroll
      for (int i = 0; i < BC_MAXK; ++i) {
        const int m = tid + 256 * i;
        kv[i] = m < Mp ? row[m] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < BM_MAXNB; ++k) {
        if (k < nq) {
          const float qd = sq[k * HD + r0 + rr];
#pragma unroll
          for (int i = 0; i < BC_MAXK; ++i) acc[k][i] += qd * kv[i];
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == BC_NST) { s = 0; ph ^= 1; }
  }
  float inv[BM_MAXNB];
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k) {
    inv[k] = 0.f;
    if (k < nq) {  // nq is uniform across the block: the named barriers below are reached by all 256 threads
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < BC_MAXK; ++i) {
        const int m = tid + 256 * i;
        if (m < Mp) {
          acc[k][i] += (mask[(int64_t)b * Mp + m] ? 0.f : -3.4028234663852886e38f);
          mx = fmaxf(mx, acc[k][i]);
        }
      }
      mx = warp_max(mx);
      if (lane == 0) s_b[warp] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = s_b[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_b[w]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < BC_MAXK; ++i) {
        const int m = tid + 256 * i;
        if (m < Mp) {
          const float p = expf(acc[k][i] - mx);
          sc[(int64_t)k * Mp + m] = p;
          sum += p;
        }
      }
      sum = warp_sum(sum);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) s_b[warp] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += s_b[w];
      inv[k] = 1.f / sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  const int r = tid >> 4, cc = tid & 15;
  float4 a4[BM_MAXNB];
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k) a4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nvc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const float4* buf4 = reinterpret_cast<const float4*>(ring + s * (BC_STAGE_BYTES / 4));
    const int m0 = c * VR, rows = min(VR, Mp - m0);
#pragma unroll
    for (int j = 0; j < (VR + 15) / 16; ++j) {
      const int jj = r + 16 * j;
      if (jj < rows) {
        const float4 vv = buf4[jj * 16 + cc];
#pragma unroll
        for (int k = 0; k < BM_MAXNB; ++k) {
          if (k < nq) {
            const float p = sc[(int64_t)k * Mp + m0 + jj];
            a4[k].x += p * vv.x; a4[k].y += p * vv.y; a4[k].z += p * vv.z; a4[k].w += p * vv.w;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == BC_NST) { s = 0; ph ^= 1; }
  }
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k) {
    if (k < nq) {
      reinterpret_cast<float4*>(sred)[r * 16 + cc] = a4[k];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < HD) {
        float o = 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) o += sred[rr * HD + tid];
        ctx[((int64_t)b * nq + k) * D + h * HD + tid] = o * inv[k];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
}

void launch_beam_cross_attn(cudaStream_t st, const float* q, int B, int nq, int H, int D, const float* kt,
                            const float* v, int Mp, const int* mask, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  MG_REQUIRE(nq >= 1 && nq <= BM_MAXNB, "1 <= num_beams <= 8");
  MG_REQUIRE(Mp % 4 == 0 && Mp <= 256 * BC_MAXK, "cross-attention memory length must be a multiple of 4, <= 2048");
  const size_t smem = (size_t)BC_NST * BC_STAGE_BYTES + ((size_t)nq * Mp + (size_t)nq * 64 + 16 * 64) * 4 +
                      2 * BC_NST * 8 + 64;
  static bool attr = false;
  if (!attr) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(beam_cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  MG_REQUIRE(smem <= 200 * 1024, "num_beams * memory length too large for the beam cross-attention kernel");
  dim3 grid(H, B);
  launch_pdl(beam_cross_attn_kernel, grid, dim3(288), smem, st, q, kt, v, mask, Mp, H, D, nq, ctx);
}

// kv24 variant (decode.cu: 16-bit + 8-bit planes, 3 bytes per element, the format the greedy paths stream): one CTA
// per (head, IMAGE), the image's block streamed ONCE for its nq beams -- per generated token a beam-4 decode moves
// 3/16 of the bytes of the reference's nb-times-repeated fp32 encoder K/V.  Scores: thread = quad of adjacent keys.
__global__ void __launch_bounds__(288) beam_cross_attn24_kernel(const float* __restrict__ q, const uint8_t* __restrict__ kv,
                                                                const int* __restrict__ mask, int Mp, int H, int D,
                                                                int nq, float* __restrict__ ctx) {
  constexpr int HD = 64;
  extern __shared__ __align__(128) uint8_t smc[];
  uint8_t* ring = smc;
  float* sc = reinterpret_cast<float*>(smc + BC_NST * BC_STAGE_BYTES);  // [nq][Mp]
  float* sq = sc + (int64_t)nq * Mp;                                    // [nq][64]
  float* sred = sq + nq * HD;                                           // [16*64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sred + 16 * HD);
  uint64_t* empty_bar = full_bar + BC_NST;
  float* s_b = reinterpret_cast<float*>(empty_bar + BC_NST);            // [8 warps]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n = (int64_t)HD * Mp;
  const uint8_t* blk = kv + ((int64_t)b * H + h) * 6 * n;
  const int RK = min(HD, (BC_STAGE_BYTES / (Mp * 3)) & ~1);
  const int nkc = (HD + RK - 1) / RK;
  constexpr int VR = BC_STAGE_BYTES / (HD * 3);
  const int nvc = (Mp + VR - 1) / VR;
  griddep_launch();
  if (tid == 0) {
    for (int s = 0; s < BC_NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 8) {
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
        uint8_t* dst = ring + s * BC_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], 3 * bytes_lo);
        bulk_load_1d(dst, src_hi, 2 * bytes_lo, &full_bar[s]);
        bulk_load_1d(dst + 2 * bytes_lo, src_lo, bytes_lo, &full_bar[s]);
        if (++s == BC_NST) { s = 0; ph ^= 1; }
      }
    }
    griddep_wait();
    return;
  }
  griddep_wait();
  for (int i = tid; i < nq * HD; i += 256) sq[i] = q[((int64_t)b * nq + i / HD) * D + h * HD + (i % HD)];
  asm volatile("bar.sync 1, 256;" ::: "memory");
  int s = 0;
  uint32_t ph = 0;
  const bool kq0 = 128 * warp < Mp, kq1 = 1024 + 128 * warp < Mp;  // this warp's key quads exist (warp-uniform)
  float acc[BM_MAXNB][BC_MAXK];  // acc[k][i] = key 4 * tid + (i & 3) + 1024 * (i >> 2) of beam k
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k)
#pragma unroll
    for (int i = 0; i < BC_MAXK; ++i) acc[k][i] = 0.f;
  for (int c = 0; c < nkc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const uint8_t* buf = ring + s * BC_STAGE_BYTES;
    const int r0 = c * RK, rows = min(RK, HD - r0);
    const uint8_t* lo_base = buf + (size_t)rows * Mp * 2;
    for (int rr = 0; rr < rows; ++rr) {
      // thread = 4 adjacent keys (+ the 4 keys 1024 further on): an 8-byte load of the 16-bit plane and a 4-byte load of
      // the 8-bit plane per quad (3 shared-memory wavefronts per 128 keys; key pairs with 2-byte loads needed 4 and ran
      // every warp over all 2048 key slots -- decode_mega.cu, DESIGN.md decision 10).  kq0 / kq1 are warp-uniform; lanes
      // past the row's end inside an active warp read the next row / plane, their sums are never used.
      const uint2* hrow = reinterpret_cast<const uint2*>(buf + (size_t)rr * Mp * 2) + tid;
      const uint32_t* lrow = reinterpret_cast<const uint32_t*>(lo_base + (size_t)rr * Mp) + tid;
      float kvv[BC_MAXK];
      uint2 h0 = make_uint2(0u, 0u), h1 = make_uint2(0u, 0u);
      uint32_t l0 = 0u, l1 = 0u;
      if (kq0) { h0 = hrow[0]; l0 = lrow[0]; }
      if (kq1) { h1 = hrow[256]; l1 = lrow[256]; }
      {
        const uint32_t la = __byte_perm(l0, 0u, 0x4240), lb = __byte_perm(l0, 0u, 0x4341);
        kvv[0] = __uint_as_float(__byte_perm(h0.x, la, 0x1045));
        kvv[1] = __uint_as_float(__byte_perm(h0.x, lb, 0x3245));
        kvv[2] = __uint_as_float(__byte_perm(h0.y, la, 0x1065));
        kvv[3] = __uint_as_float(__byte_perm(h0.y, lb, 0x3265));
        const uint32_t lc = __byte_perm(l1, 0u, 0x4240), ld = __byte_perm(l1, 0u, 0x4341);
        kvv[4] = __uint_as_float(__byte_perm(h1.x, lc, 0x1045));
        kvv[5] = __uint_as_float(__byte_perm(h1.x, ld, 0x3245));
        kvv[6] = __uint_as_float(__byte_perm(h1.y, lc, 0x1065));
        kvv[7] = __uint_as_float(__byte_perm(h1.y, ld, 0x3265));
      }
#pragma unroll
      for (int k = 0; k < BM_MAXNB; ++k) {
        if (k < nq) {
          const float qd = sq[k * HD + r0 + rr];
#pragma unroll
          for (int i = 0; i < BC_MAXK; ++i) acc[k][i] += qd * kvv[i];
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == BC_NST) { s = 0; ph ^= 1; }
  }
  float inv[BM_MAXNB];
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k) {
    inv[k] = 0.f;
    if (k < nq) {  // nq is uniform across the block: the named barriers below are reached by all 256 threads
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < BC_MAXK; ++i) {
        const int m = 4 * tid + (i & 3) + 1024 * (i >> 2);
        if (m < Mp) {
          acc[k][i] += (mask[(int64_t)b * Mp + m] ? 0.f : -3.4028234663852886e38f);
          mx = fmaxf(mx, acc[k][i]);
        }
      }
      mx = warp_max(mx);
      if (lane == 0) s_b[warp] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = s_b[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_b[w]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < BC_MAXK; ++i) {
        const int m = 4 * tid + (i & 3) + 1024 * (i >> 2);
        if (m < Mp) {
          const float p = expf(acc[k][i] - mx);
          sc[(int64_t)k * Mp + m] = p;
          sum += p;
        }
      }
      sum = warp_sum(sum);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) s_b[warp] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += s_b[w];
      inv[k] = 1.f / sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  const int r = tid >> 4, cc = tid & 15;
  float4 a4[BM_MAXNB];
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k) a4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nvc; ++c) {
    mbar_wait(&full_bar[s], ph);
    const uint8_t* buf = ring + s * BC_STAGE_BYTES;
    const int m0 = c * VR, rows = min(VR, Mp - m0);
    const uint8_t* lo_base = buf + (size_t)rows * HD * 2;
#pragma unroll
    for (int j = 0; j < (VR + 15) / 16; ++j) {
      const int jj = r + 16 * j;
      if (jj < rows) {
        const uint2 h4 = *reinterpret_cast<const uint2*>(buf + ((size_t)jj * HD + 4 * cc) * 2);
        const uint32_t l4 = *reinterpret_cast<const uint32_t*>(lo_base + (size_t)jj * HD + 4 * cc);
        float4 vv;
        vv.x = __uint_as_float((h4.x << 16) | ((l4 & 0xffu) << 8));
        vv.y = __uint_as_float((h4.x & 0xffff0000u) | (l4 & 0xff00u));
        vv.z = __uint_as_float((h4.y << 16) | ((l4 >> 8) & 0xff00u));
        vv.w = __uint_as_float((h4.y & 0xffff0000u) | ((l4 >> 16) & 0xff00u));
#pragma unroll
        for (int k = 0; k < BM_MAXNB; ++k) {
          if (k < nq) {
            const float p = sc[(int64_t)k * Mp + m0 + jj];
            a4[k].x += p * vv.x; a4[k].y += p * vv.y; a4[k].z += p * vv.z; a4[k].w += p * vv.w;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (++s == BC_NST) { s = 0; ph ^= 1; }
  }
#pragma unroll
  for (int k = 0; k < BM_MAXNB; ++k) {
    if (k < nq) {
      reinterpret_cast<float4*>(sred)[r * 16 + cc] = a4[k];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < HD) {
        float o = 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) o += sred[rr * HD + tid];
        ctx[((int64_t)b * nq + k) * D + h * HD + tid] = o * inv[k];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
}

void launch_beam_cross_attn24(cudaStream_t st, const float* q, int B, int nq, int H, int D, const uint8_t* kv, int Mp,
                              const int* mask, float* ctx) {
  MG_REQUIRE(D == H * 64, "decoder head_dim must be 64");
  MG_REQUIRE(nq >= 1 && nq <= BM_MAXNB, "1 <= num_beams <= 8");
  MG_REQUIRE(Mp % 8 == 0 && Mp <= 256 * BC_MAXK, "cross-attention memory length must be a multiple of 8, <= 2048");
  const size_t smem = (size_t)BC_NST * BC_STAGE_BYTES + ((size_t)nq * Mp + (size_t)nq * 64 + 16 * 64) * 4 +
                      2 * BC_NST * 8 + 64;
  static bool attr = false;
  if (!attr) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(beam_cross_attn24_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  MG_REQUIRE(smem <= 200 * 1024, "num_beams * memory length too large for the beam cross-attention kernel");
  dim3 grid(H, B);
  launch_pdl(beam_cross_attn24_kernel, grid, dim3(288), smem, st, q, kv, mask, Mp, H, D, nq, ctx);
}

// =====================================================================================================
// Beam state (per image b, beam k):  run_seq / fin_seq [B][nb][L] i64, run_score / fin_score [B][nb] f32,
// fin_flag [B][nb] (is_sent_finished), fin_len [B][nb] (generated tokens of the finished hypothesis),
// unsat [B] (is_early_stop_heuristic_unsatisfied).  ctrl: [0]=step [1]=anc_sel [2]=done [3]=ticket
// [4]=n_unsat (this step) [5]=n_all_hit (this step) [6]=steps executed.

__global__ void beam_init_kernel(BeamState s, const float* __restrict__ emb, int D, int start, int pad,
                                 float* __restrict__ x) {
  const int b = blockIdx.x;
  const int nb = s.nb;
  for (int i = threadIdx.x; i < nb * s.L; i += blockDim.x) {
    const int pos = i % s.L;
    const int64_t val = pos == 0 ? start : pad;
    s.run_seq[(int64_t)b * nb * s.L + i] = val;
    s.fin_seq[(int64_t)b * nb * s.L + i] = val;
  }
  for (int i = threadIdx.x; i < nb * s.anc_ld; i += blockDim.x) {
    const int k = i / s.anc_ld;
    s.anc0[(int64_t)(b * nb) * s.anc_ld + i] = b * nb + k;
    s.anc1[(int64_t)(b * nb) * s.anc_ld + i] = b * nb + k;
  }
  for (int i = threadIdx.x; i < nb * D; i += blockDim.x) x[(int64_t)b * nb * D + i] = emb[(int64_t)start * D + (i % D)];
  if (threadIdx.x < nb) {
    const int k = threadIdx.x;
    s.run_score[b * nb + k] = k == 0 ? 0.f : BM_NEG;
    s.fin_score[b * nb + k] = BM_NEG;
    s.fin_flag[b * nb + k] = 0;
    s.fin_len[b * nb + k] = 0;
  }
  if (threadIdx.x == 0) {
    s.unsat[b] = 1;
    if (b == 0)
      for (int i = 0; i < 8; ++i) s.ctrl[i] = 0;
  }
}

void launch_beam_init(cudaStream_t st, const BeamState& s, const float* emb, int D, int start, int pad, float* x) {
  beam_init_kernel<<<s.B, 256, 0, st>>>(s, emb, D, start, pad, x);
  MG_CHECK_CUDA(cudaGetLastError());
}

// one CTA (1024 threads) per image
__global__ void __launch_bounds__(1024) beam_select_kernel(BeamState s, const float* __restrict__ logits, int V,
                                                           int64_t ld, const float* __restrict__ emb, int D, int eos,
                                                           int max_length, float* __restrict__ x_next) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float s_lse[BM_MAXNB], s_rs[BM_MAXNB];
  __shared__ float c_lp[2 * BM_MAXNB];
  __shared__ int c_beam[2 * BM_MAXNB], c_tok[2 * BM_MAXNB];
  __shared__ int n_src[BM_MAXNB], n_tok[BM_MAXNB];
  __shared__ int s_skip;
  const int b = blockIdx.x, nb = s.nb, L = s.L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  griddep_launch();
  griddep_wait();
  const int step = s.ctrl[0];
  const int sel = s.ctrl[1];
  if (tid == 0) s_skip = s.ctrl[2];
  __syncthreads();
  const bool frozen = s_skip != 0;   // the global stop condition was reached in an earlier step: state is final
  const int cur_len = step + 1;      // tokens in the running sequences before this step (prompt length 1)
  if (!frozen) {
    // ---- log-softmax statistics per beam: lse_k = max + log(sum exp(x - max))
    for (int k = 0; k < nb; ++k) {
      const float* lg = logits + (int64_t)(b * nb + k) * ld;
      float mx = -INFINITY;
      for (int i = tid; i < V; i += 1024) mx = fmaxf(mx, lg[i]);
      mx = warp_max(mx);
      if (lane == 0) s_val[warp] = mx;
      __syncthreads();
      mx = s_val[0];
      for (int w = 1; w < 32; ++w) mx = fmaxf(mx, s_val[w]);
      __syncthreads();
      float sum = 0.f;
      for (int i = tid; i < V; i += 1024) sum += expf(lg[i] - mx);
      sum = warp_sum(sum);
      if (lane == 0) s_val[warp] = sum;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < 32; ++w) t += s_val[w];
        s_lse[k] = mx + logf(t);
        s_rs[k] = s.run_score[b * nb + k];
      }
      __syncthreads();
    }
    // ---- top-2nb of (log_softmax + running score) over nb*V candidates, in (value desc, flat index asc) order
    float last_v = INFINITY;
    int last_i = -1;
    for (int c = 0; c < 2 * nb; ++c) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int k = 0; k < nb; ++k) {
        const float* lg = logits + (int64_t)(b * nb + k) * ld;
        const float off = s_lse[k], rs = s_rs[k];
        for (int i = tid; i < V; i += 1024) {
          const float val = (lg[i] - off) + rs;
          const int idx = k * V + i;
          const bool after = (val < last_v) || (val == last_v && idx > last_i);
          if (after && (val > bv || (val == bv && idx < bi))) {
            bv = val;
            bi = idx;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_val[warp] = bv;
        s_idx[warp] = bi;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < 32; ++w)
          if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) {
            bv = s_val[w];
            bi = s_idx[w];
          }
        c_lp[c] = bv;
        c_beam[c] = bi / V;
        c_tok[c] = bi % V;
        s_val[0] = bv;
        s_idx[0] = bi;
      }
      __syncthreads();
      last_v = s_val[0];
      last_i = s_idx[0];
      __syncthreads();
    }
    // ---- sequential bookkeeping (tiny): thread 0 decides, then all threads move the sequences
    __shared__ int f_from[BM_MAXNB];     // new finished slot <- index into merged list (0..nb-1 old, nb.. candidates)
    __shared__ float f_score[BM_MAXNB];
    __shared__ int f_flag[BM_MAXNB], f_len[BM_MAXNB];
    if (tid == 0) {
      const int K2 = 2 * nb;
      bool hit[2 * BM_MAXNB];
      bool all_hit = true;
      for (int c = 0; c < K2; ++c) {
        hit[c] = (c_tok[c] == eos) || (cur_len + 1 >= max_length);
        all_hit = all_hit && hit[c];
      }
      // running beams for the next iteration: top-nb of lp + hit * -1e9 (stable in candidate order)
      float rl[2 * BM_MAXNB];
      bool used[2 * BM_MAXNB];
      for (int c = 0; c < K2; ++c) {
        rl[c] = c_lp[c] + (hit[c] ? 1.f : 0.f) * BM_NEG;
        used[c] = false;
      }
      float new_rs[BM_MAXNB];
      for (int k = 0; k < nb; ++k) {
        int best = -1;
        for (int c = 0; c < K2; ++c)
          if (!used[c] && (best < 0 || rl[c] > rl[best])) best = c;
        used[best] = true;
        n_src[k] = c_beam[best];
        n_tok[k] = c_tok[best];
        new_rs[k] = rl[best];
      }
      // finished beams: merge old finished with the candidates that just finished inside the top nb
      const float denom = (float)(cur_len + 1 - 1);  // (cur_len + 1 - decoder_prompt_len) ** length_penalty(=1)
      float ms[3 * BM_MAXNB];
      int mflag[3 * BM_MAXNB];
      const bool unsat = s.unsat[b] != 0;
      for (int k = 0; k < nb; ++k) {
        ms[k] = s.fin_score[b * nb + k];
        mflag[k] = s.fin_flag[b * nb + k];
      }
      for (int c = 0; c < K2; ++c) {
        const bool did = hit[c] && c < nb;
        float sc = c_lp[c] / denom;
        // early_stopping is False: the "beams full" term never applies
        sc += (unsat ? 0.f : 1.f) * BM_NEG;
        sc += (did ? 0.f : 1.f) * BM_NEG;
        ms[nb + c] = sc;
        mflag[nb + c] = did ? 1 : 0;
      }
      bool mused[3 * BM_MAXNB];
      for (int i = 0; i < nb + K2; ++i) mused[i] = false;
      for (int k = 0; k < nb; ++k) {
        int best = -1;
        for (int i = 0; i < nb + K2; ++i)
          if (!mused[i] && (best < 0 || ms[i] > ms[best])) best = i;
        mused[best] = true;
        f_from[k] = best;
        f_score[k] = ms[best];
        f_flag[k] = mflag[best];
        f_len[k] = best < nb ? s.fin_len[b * nb + best] : cur_len;  // generated tokens incl. the new one
      }
      // early-stop heuristic for the next iteration (cur_len already advanced by one)
      bool all_fin = true, any_better = false;
      float worst = INFINITY;
      for (int k = 0; k < nb; ++k) worst = fminf(worst, f_score[k]);
      const float best_running = new_rs[0] / (float)(cur_len + 1 - 1);
      for (int k = 0; k < nb; ++k) {
        const float wf = f_flag[k] ? worst : BM_NEG;
        any_better = any_better || (best_running > wf);
        all_fin = all_fin && f_flag[k];
      }
      (void)all_fin;
      const int new_unsat = (unsat && any_better) ? 1 : 0;
      s.unsat[b] = new_unsat;
      for (int k = 0; k < nb; ++k) s.run_score[b * nb + k] = new_rs[k];
      if (new_unsat) atomicAdd(&s.ctrl[4], 1);
      if (all_hit) atomicAdd(&s.ctrl[5], 1);
    }
    __syncthreads();
    // ---- finished sequences: gather from (old finished | candidate = running[c_beam] + c_tok) into a scratch
    // copy first (slots permute), using the not-yet-updated running sequences
    int64_t* run = s.run_seq + (int64_t)b * nb * L;
    int64_t* fin = s.fin_seq + (int64_t)b * nb * L;
    int64_t* tmp = s.tmp_seq + (int64_t)b * nb * L;
    for (int k = 0; k < nb; ++k) {
      const int from = f_from[k];
      for (int p = tid; p < L; p += 1024) {
        int64_t val;
        if (from < nb) {
          val = fin[(int64_t)from * L + p];
        } else {
          const int c = from - nb;
          val = (p == cur_len) ? (int64_t)c_tok[c] : run[(int64_t)c_beam[c] * L + p];
        }
        tmp[(int64_t)k * L + p] = val;
      }
    }
    __syncthreads();
    for (int i = tid; i < nb * L; i += 1024) fin[i] = tmp[i];
    if (tid < nb) {
      s.fin_score[b * nb + tid] = f_score[tid];
      s.fin_flag[b * nb + tid] = f_flag[tid];
      s.fin_len[b * nb + tid] = f_len[tid];
    }
    __syncthreads();
    // ---- running sequences for the next step (same two-phase move) + ancestry + next-step embeddings
    for (int k = 0; k < nb; ++k)
      for (int p = tid; p < L; p += 1024)
        tmp[(int64_t)k * L + p] = (p == cur_len) ? (int64_t)n_tok[k] : run[(int64_t)n_src[k] * L + p];
    __syncthreads();
    for (int i = tid; i < nb * L; i += 1024) run[i] = tmp[i];
    const int* anc_old = (sel ? s.anc1 : s.anc0) + (int64_t)(b * nb) * s.anc_ld;
    int* anc_new = (sel ? s.anc0 : s.anc1) + (int64_t)(b * nb) * s.anc_ld;
    for (int k = 0; k < nb; ++k)
      for (int p = tid; p <= step; p += 1024) anc_new[(int64_t)k * s.anc_ld + p] = anc_old[(int64_t)n_src[k] * s.anc_ld + p];
    for (int i = tid; i < nb * (D / 4); i += 1024) {
      const int k = i / (D / 4), c = i % (D / 4);
      reinterpret_cast<float4*>(x_next + (int64_t)(b * nb + k) * D)[c] =
          reinterpret_cast<const float4*>(emb + (int64_t)n_tok[k] * D)[c];
    }
    if (s.step_tok && tid == 0) s.step_tok[b] = n_tok[0];  // provisional token of the best running beam
  } else if (s.step_tok && tid == 0) {
    s.step_tok[b] = -1;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int t = atomicAdd(&s.ctrl[3], 1);
    if (t == (int)gridDim.x - 1) {
      s.ctrl[3] = 0;
      if (!frozen) {
        // _beam_search_has_unfinished_sequences: any image can still improve AND not every candidate of every
        // image hit a stopping criterion (early_stopping False -> "exists open beam" is always true)
        const bool improvement_possible = s.ctrl[4] > 0;
        const bool valid_continuations = s.ctrl[5] < (int)gridDim.x;
        if (!(improvement_possible && valid_continuations)) s.ctrl[2] = 1;
        s.ctrl[6] = step + 1;
        s.ctrl[1] = sel ^ 1;
      }
      s.ctrl[4] = 0;
      s.ctrl[5] = 0;
      s.ctrl[0] = step + 1;
      if (s.step_tok) s.step_tok[gridDim.x] = s.ctrl[2];  // this rank's "done" flag travels with the step's tokens
    }
  }
}

void launch_beam_select(cudaStream_t st, const BeamState& s, const float* logits, int V, int64_t ld, const float* emb,
                        int D, int eos, int max_length, float* x_next) {
  MG_REQUIRE(s.nb <= BM_MAXNB, "1 <= num_beams <= 8");
  launch_pdl(beam_select_kernel, dim3(s.B), dim3(1024), (size_t)0, st, s, logits, V, ld, emb, D, eos, max_length, x_next);
}

// best finished hypothesis of every image -> out_ids (B, L) padded with pad; out_len = 1 + generated length
__global__ void beam_finalize_kernel(BeamState s, int pad, int64_t* __restrict__ out_ids, int* __restrict__ out_len) {
  const int b = blockIdx.x;
  const int len = 1 + s.fin_len[b * s.nb];
  for (int p = threadIdx.x; p < s.L; p += blockDim.x)
    out_ids[(int64_t)b * s.L + p] = p < len ? s.fin_seq[(int64_t)b * s.nb * s.L + p] : (int64_t)pad;
  if (threadIdx.x == 0 && out_len) out_len[b] = len;
}
void launch_beam_finalize(cudaStream_t st, const BeamState& s, int pad, int64_t* out_ids, int* out_len) {
  beam_finalize_kernel<<<s.B, 256, 0, st>>>(s, pad, out_ids, out_len);
  MG_CHECK_CUDA(cudaGetLastError());
}

// multi-GPU beam search, after the per-step all-gather: gathered[r] = {token of the best running beam of each of
// rank r's B images (-1 once that rank's search is frozen), rank r's done flag}.  Fills the provisional column of the
// global id matrix (the final sequences replace it after the last step) and counts the ranks still searching.
__global__ void beam_scatter_step_kernel(const int* __restrict__ gathered, int world, int B, int col, int ld,
                                         int64_t* __restrict__ all_ids, int* __restrict__ n_not_done) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < world * B) {
    const int r = i / B, b = i - r * B;
    const int tok = gathered[r * (B + 1) + b];
    if (tok >= 0 && col < ld) all_ids[(int64_t)i * ld + col] = tok;
  }
  if (i == 0) {
    int n = 0;
    for (int r = 0; r < world; ++r) n += gathered[r * (B + 1) + B] == 0;
    *n_not_done = n;
  }
}
void launch_beam_scatter_step(cudaStream_t st, const int* gathered, int world, int B, int col, int ld, int64_t* all_ids,
                              int* n_not_done) {
  beam_scatter_step_kernel<<<(world * B + 127) / 128, 128, 0, st>>>(gathered, world, B, col, ld, all_ids, n_not_done);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
