// OCSR branch (MolScribe Swin-B) non-GEMM kernels: bilinear resize, LayerNorm with window-partition /
// cyclic-shift / patch-merge gathers folded into the row index, and shifted-window attention.
// Arithmetic follows transformers/models/swin/modeling_swin.py (SwinLayer :534-653, SwinSelfAttention
// :383-459, SwinPatchMerging :298-350), the stock restatement of timm-0.4.12 swin_base_patch4_window12_384
// that MolScribe's encoder wraps (reference requirements.txt:25).
#include <algorithm>

#include <string>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

// =====================================================================================================
// F.interpolate(mode="bilinear", align_corners=False, antialias=False)   (B,3,Hi,Wi) -> (B,3,Ho,Wo)
__global__ void resize_bilinear_kernel(const float* __restrict__ in, int BC, int Hi, int Wi, int Ho, int Wo,
                                       float* __restrict__ out) {
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  const int64_t total = (int64_t)BC * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const int64_t bc = i / ((int64_t)Wo * Ho);
    float sy = sh * ((float)oy + 0.5f) - 0.5f;
    float sx = sw * ((float)ox + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    float ly = sy - (float)y0, lx = sx - (float)x0;
    ly = fminf(fmaxf(ly, 0.f), 1.f);
    lx = fminf(fmaxf(lx, 0.f), 1.f);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = in + bc * Hi * Wi;
    const float top = hx * p[(int64_t)y0 * Wi + x0] + lx * p[(int64_t)y0 * Wi + x1];
    const float bot = hx * p[(int64_t)y1 * Wi + x0] + lx * p[(int64_t)y1 * Wi + x1];
    out[i] = hy * top + ly * bot;
  }
}

void launch_resize_bilinear(cudaStream_t st, const float* in, int B, int Hi, int Wi, int Ho, int Wo, float* out) {
  const int64_t total = (int64_t)B * 3 * Ho * Wo;
  if (!total) return;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  resize_bilinear_kernel<<<blocks, 256, 0, st>>>(in, B * 3, Hi, Wi, Ho, Wo, out);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Row maps.  Window-order row m = ((b*nWy + wy)*nWx + wx)*ws*ws + iy*ws + ix  <->  spatial row
// b*Hs*Ws + ((wy*ws+iy+shift)%Hs)*Ws + (wx*ws+ix+shift)%Ws   (torch.roll by -shift then window_partition;
// window_reverse + roll by +shift is the same bijection read backwards, modeling_swin.py:606-636).
__global__ void window_rowmap_kernel(int B, int Hs, int Ws, int ws, int shift, int* __restrict__ map) {
  const int nWx = Ws / ws, nWy = Hs / ws;
  const int64_t total = (int64_t)B * Hs * Ws;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(m % (ws * ws));
    const int64_t w = m / (ws * ws);
    const int wx = (int)(w % nWx), wy = (int)((w / nWx) % nWy), b = (int)(w / ((int64_t)nWx * nWy));
    const int iy = t / ws, ix = t % ws;
    const int y = (wy * ws + iy + shift) % Hs, x = (wx * ws + ix + shift) % Ws;
    map[m] = (int)((int64_t)b * Hs * Ws + (int64_t)y * Ws + x);
  }
}
void launch_window_rowmap(cudaStream_t st, int B, int Hs, int Ws, int ws, int shift, int* map) {
  const int64_t total = (int64_t)B * Hs * Ws;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  window_rowmap_kernel<<<blocks, 256, 0, st>>>(B, Hs, Ws, ws, shift, map);
  MG_CHECK_CUDA(cudaGetLastError());
}

// Patch-merge gather (SwinPatchMerging :332-343): output row (b, i, j) = cat[x(2i,2j), x(2i+1,2j), x(2i,2j+1), x(2i+1,2j+1)]
__global__ void merge_rowmap_kernel(int B, int Hs, int Ws, int* __restrict__ map) {
  const int Ho = Hs / 2, Wo = Ws / 2;
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < total; m += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(m % Wo), i = (int)((m / Wo) % Ho), b = (int)(m / ((int64_t)Wo * Ho));
    const int64_t base = (int64_t)b * Hs * Ws;
    map[m * 4 + 0] = (int)(base + (int64_t)(2 * i) * Ws + 2 * j);
    map[m * 4 + 1] = (int)(base + (int64_t)(2 * i + 1) * Ws + 2 * j);
    map[m * 4 + 2] = (int)(base + (int64_t)(2 * i) * Ws + 2 * j + 1);
    map[m * 4 + 3] = (int)(base + (int64_t)(2 * i + 1) * Ws + 2 * j + 1);
  }
}
void launch_merge_rowmap(cudaStream_t st, int B, int Hs, int Ws, int* map) {
  const int64_t total = (int64_t)B * (Hs / 2) * (Ws / 2);
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  merge_rowmap_kernel<<<blocks, 256, 0, st>>>(B, Hs, Ws, map);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// LayerNorm over rows assembled from G gathered source rows of C channels each (G=1: plain / window gather,
// G=4: patch merge).  y = (x - mean) * rsqrt(var + eps) * w + b, biased variance (torch.nn.LayerNorm).
// One warp per output row; the row lives in registers (G*C <= 4096).
template <int MAXV>
__global__ void layernorm_gather_kernel(const float* __restrict__ x, const int* __restrict__ src_rows, int G, int C,
                                        int64_t rows, const float* __restrict__ w, const float* __restrict__ bvec,
                                        float eps, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                                        float* __restrict__ out_f32) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int D = G * C;
  const int nv = D / 128;  // float4 per lane
  const int cv = C / 4;    // float4 per source row
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const int c4 = lane + 32 * i;
      const int g = c4 / cv;
      const int64_t sr = src_rows ? (int64_t)src_rows[r * G + g] : r;
      v[i] = reinterpret_cast<const float4*>(x + sr * C)[c4 - g * cv];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  s = warp_sum(s);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const int c4 = lane + 32 * i;
      const float4 ww = reinterpret_cast<const float4*>(w)[c4];
      const float4 bb = reinterpret_cast<const float4*>(bvec)[c4];
      float o[4] = {(v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y,
                    (v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w};
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * D + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
      if (out_hi) {
        bf16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16(o[k], h[k], l[k]);
        uint2 ph, pl;
        ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
        pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
        pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
        *reinterpret_cast<uint2*>(out_hi + r * D + c4 * 4) = ph;
        if (out_lo) *reinterpret_cast<uint2*>(out_lo + r * D + c4 * 4) = pl;
      }
    }
  }
}

void launch_layernorm(cudaStream_t st, const float* x, const int* src_rows, int G, int C, int64_t rows, const float* w,
                      const float* b, float eps, Planes out, float* out_f32) {
  const int D = G * C;
  MG_REQUIRE(D % 128 == 0 && C % 4 == 0 && D <= 4096, "layernorm: row width must be a multiple of 128 and <= 4096");
  if (!rows) return;
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (D <= 1024)
    layernorm_gather_kernel<8><<<grid, wpb * 32, 0, st>>>(x, src_rows, G, C, rows, w, b, eps, out.hi, out.lo, out_f32);
  else
    layernorm_gather_kernel<32><<<grid, wpb * 32, 0, st>>>(x, src_rows, G, C, rows, w, b, eps, out.hi, out.lo, out_f32);
  MG_CHECK_CUDA(cudaGetLastError());
}

// LayerNorm for narrow rows (C < 128 or not a multiple of 128, e.g. tiny test configs): scalar loads
__global__ void layernorm_small_kernel(const float* __restrict__ x, const int* __restrict__ src_rows, int G, int C,
                                       int64_t rows, const float* __restrict__ w, const float* __restrict__ bvec,
                                       float eps, bf16* out_hi, bf16* out_lo, float* out_f32) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int D = G * C;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) {
    const int g = c / C;
    const int64_t sr = src_rows ? (int64_t)src_rows[r * G + g] : r;
    s += x[sr * C + (c - g * C)];
  }
  s = warp_sum(s);
  const float mean = s / (float)D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) {
    const int g = c / C;
    const int64_t sr = src_rows ? (int64_t)src_rows[r * G + g] : r;
    const float a = x[sr * C + (c - g * C)] - mean;
    q += a * a;
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)D + eps);
  for (int c = lane; c < D; c += 32) {
    const int g = c / C;
    const int64_t sr = src_rows ? (int64_t)src_rows[r * G + g] : r;
    const float o = (x[sr * C + (c - g * C)] - mean) * rstd * w[c] + bvec[c];
    if (out_f32) out_f32[r * D + c] = o;
    if (out_hi) {
      bf16 h, l;
      split_bf16(o, h, l);
      out_hi[r * D + c] = h;
      if (out_lo) out_lo[r * D + c] = l;
    }
  }
}

void launch_layernorm_any(cudaStream_t st, const float* x, const int* src_rows, int G, int C, int64_t rows,
                          const float* w, const float* b, float eps, Planes out, float* out_f32) {
  const int D = G * C;
  if (D % 128 == 0 && C % 4 == 0 && D <= 4096) {
    launch_layernorm(st, x, src_rows, G, C, rows, w, b, eps, out, out_f32);
    return;
  }
  if (!rows) return;
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  layernorm_small_kernel<<<grid, wpb * 32, 0, st>>>(x, src_rows, G, C, rows, w, b, eps, out.hi, out.lo, out_f32);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Shifted-window attention (SwinSelfAttention.forward :410-459 + get_attn_mask :556-582), one CTA per
// (window, head).  qkv: fp32 [rows(window order), 3C] = [q | k | v] with bias already added by the GEMM.
//   s_ij = (q_i . k_j) / sqrt(hd) + table[(yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1)][h] + (label_i != label_j ? -100 : 0)
// Generic fp32 path (any window size; the Swin-B geometry runs window_attn_mma_kernel below): K and V of the window
// are copied into shared memory with plain vector loads; each thread owns one query row and runs an online softmax
// over the ws*ws keys with fp32 FMAs.
template <int HD>
__global__ void __launch_bounds__(160) window_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ table,
                                                          int C, int heads, int ws, int Hs, int Ws, int shift,
                                                          bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
  extern __shared__ __align__(16) float smw[];
  const int T = ws * ws;
  float* sk = smw;            // [T][HD]
  float* sv = sk + T * HD;    // [T][HD]
  float* stab = sv + T * HD;  // [(2ws-1)^2]
  int* slab = reinterpret_cast<int*>(stab + (2 * ws - 1) * (2 * ws - 1));  // [T]
  const int h = blockIdx.x % heads;
  const int64_t win = blockIdx.x / heads;
  const int nWx = Ws / ws, nWy = Hs / ws;
  const int wx = (int)(win % nWx), wy = (int)((win / nWx) % nWy);
  const int64_t row0 = win * T;
  const int ld = 3 * C;
  for (int i = threadIdx.x; i < T * (HD / 4); i += blockDim.x) {
    const int t = i / (HD / 4), c = i % (HD / 4);
    const float4* src = reinterpret_cast<const float4*>(qkv + (row0 + t) * ld + h * HD);
    reinterpret_cast<float4*>(sk + t * HD)[c] = src[(C) / 4 + c];
    reinterpret_cast<float4*>(sv + t * HD)[c] = src[(2 * C) / 4 + c];
  }
  const int ntab = (2 * ws - 1) * (2 * ws - 1);
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) stab[i] = table[(int64_t)i * heads + h];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int lab = 0;
    if (shift > 0) {
      const int y = wy * ws + t / ws, x = wx * ws + t % ws;  // coordinates in the shifted image
      const int ry = y < Hs - ws ? 0 : (y < Hs - shift ? 1 : 2);
      const int rx = x < Ws - ws ? 0 : (x < Ws - shift ? 1 : 2);
      lab = ry * 3 + rx;
    }
    slab[t] = lab;
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= T) return;
  float q[HD];
  {
    const float4* src = reinterpret_cast<const float4*>(qkv + (row0 + i) * ld + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 t = src[c];
      q[4 * c] = t.x; q[4 * c + 1] = t.y; q[4 * c + 2] = t.z; q[4 * c + 3] = t.w;
    }
  }
  const float inv_div = sqrtf((float)HD);
  const int yi = i / ws, xi = i % ws;
  const int li = slab[i];
  float acc[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) acc[c] = 0.f;
  float mx = -INFINITY, sum = 0.f;
  for (int j = 0; j < T; ++j) {
    const float4* kj = reinterpret_cast<const float4*>(sk + j * HD);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 kk = kj[c];
      d0 += q[4 * c] * kk.x;
      d1 += q[4 * c + 1] * kk.y;
      d2 += q[4 * c + 2] * kk.z;
      d3 += q[4 * c + 3] * kk.w;
    }
    float s = ((d0 + d1) + (d2 + d3)) / inv_div;
    const int yj = j / ws, xj = j % ws;
    s += stab[(yi - yj + ws - 1) * (2 * ws - 1) + (xi - xj + ws - 1)];
    if (slab[j] != li) s += -100.0f;
    const float nm = fmaxf(mx, s);
    const float corr = expf(mx - nm);
    const float p = expf(s - nm);
    sum = sum * corr + p;
    const float4* vj = reinterpret_cast<const float4*>(sv + j * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 vv = vj[c];
      acc[4 * c] = acc[4 * c] * corr + p * vv.x;
      acc[4 * c + 1] = acc[4 * c + 1] * corr + p * vv.y;
      acc[4 * c + 2] = acc[4 * c + 2] * corr + p * vv.z;
      acc[4 * c + 3] = acc[4 * c + 3] * corr + p * vv.w;
    }
    mx = nm;
  }
  const float inv = 1.f / sum;
  bf16* oh = out_hi + (row0 + i) * C + h * HD;
  bf16* ol = out_lo ? out_lo + (row0 + i) * C + h * HD : nullptr;
#pragma unroll
  for (int c = 0; c < HD; c += 2) {
    bf16 h0, l0, h1, l1;
    split_bf16(acc[c] * inv, h0, l0);
    split_bf16(acc[c + 1] * inv, h1, l1);
    *reinterpret_cast<uint32_t*>(oh + c) = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    if (ol)
      *reinterpret_cast<uint32_t*>(ol + c) = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
}


// =====================================================================================================
// Shifted-window attention on the tensor cores (the path of every Swin-B block: window 12 -> 144 tokens, head_dim 32).
// One CTA per (window, head), one warp per 16-row strip of the window (9 warps at 144 tokens):
//   * the window's q / k / v slices -- 144 rows x 128 bytes each, strided in the fp32 qkv matrix -- are staged into
//     shared memory by THREE TMA tile loads (cp.async.bulk.tensor.2d, 128-byte swizzle, one mbarrier);
//   * S = Q K^T and O = P V run as warp-level bf16 MMAs (m16n8k16, fp32 accumulate) on split planes: every fp32
//     operand is split into hi + lo bf16 while its fragment is built and three products (lo*hi + hi*lo + hi*hi)
//     restore ~fp32 accuracy, the same numerics policy as the tcgen05 GEMMs.  A 144 x 144 x 32 problem per (window,
//     head) fills 16-row strips exactly; a 128-row tcgen05 tile would idle 44 % of its rows;
//   * the whole score row (144 keys) of a strip lives in registers: scale 1/sqrt(32), relative-position bias from the
//     (2 ws - 1)^2 table, shift mask (-100 across region labels), exact fp32 softmax -- no online rescaling -- and the
//     probabilities feed the P V MMAs straight from the accumulator fragments.
// SwinSelfAttention.forward (TF/models/swin/modeling_swin.py:410-459), get_attn_mask (:556-582).
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) fp32 -> packed bf16x2 hi and lo planes (x in the low half)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 f = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// element (row r, column c) of a [rows][32] fp32 tile stored by TMA with the 128-byte swizzle
__device__ __forceinline__ const float* sw_at(const float* tile, int r, int c) {
  return tile + r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3));
}

constexpr int WA_T = 144;      // tokens per window (12 x 12)
constexpr int WA_NT = WA_T / 8;   // 18 key tiles of 8
constexpr int WA_WARPS = WA_T / 16;

struct alignas(64) WinAttnParams {
  CUtensorMap tq;  // fp32 qkv [rows][3C], box 32 x 144
  const float* table;
  bf16 *out_hi, *out_lo;
  int C, heads, ws, Hs, Ws, shift;
};

// MINB = CTAs per SM the register allocation aims at: 1 -> 154 registers, no spills; 2 -> 96 registers, ~300 bytes of
// spills but twice the warps to hide the TMA / LDS / MMA latencies behind (A/B: MG_SWIN_MINB)
template <int MINB>
__global__ void __launch_bounds__(WA_WARPS * 32, MINB) window_attn_mma_kernel(const __grid_constant__ WinAttnParams p) {
  extern __shared__ uint8_t smw_raw[];
  uint8_t* const sm = smw_raw + ((1024u - (smem_u32(smw_raw) & 1023u)) & 1023u);
  float* const sq = reinterpret_cast<float*>(sm);                  // [144][32] swizzled
  float* const sk = sq + WA_T * 32;
  float* const sv = sk + WA_T * 32;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(sv + WA_T * 32);  // 8-byte aligned: right behind the tiles
  float* const stab = reinterpret_cast<float*>(bar + 2);             // [(2 ws - 1)^2]
  int* const slab = reinterpret_cast<int*>(stab + 23 * 23);          // [144] region label (shift mask)
  int* const scol = slab + WA_T;                                     // [144] yj * (2 ws - 1) + xj
  const int ws = p.ws, heads = p.heads, C = p.C;
  const int h = blockIdx.x % heads;
  const int64_t win = blockIdx.x / heads;
  const int nWx = p.Ws / ws, nWy = p.Hs / ws;
  const int wx = (int)(win % nWx), wy = (int)((win / nWx) % nWy);
  const int64_t row0 = win * WA_T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, 3 * WA_T * 32 * 4);
    for (int which = 0; which < 3; ++which)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(sq + which * WA_T * 32)),
                   "l"(reinterpret_cast<uint64_t>(&p.tq)), "r"(smem_u32(bar)), "r"(which * C + h * 32), "r"((int)row0)
                   : "memory");
  }
  const int tw = 2 * ws - 1;
  for (int i = tid; i < tw * tw; i += blockDim.x) stab[i] = p.table[(int64_t)i * heads + h];
  for (int t = tid; t < WA_T; t += blockDim.x) {
    int lab = 0;
    const int ty = t / ws, tx = t % ws;
    if (p.shift > 0) {
      const int y = wy * ws + ty, x = wx * ws + tx;  // coordinates in the shifted image
      const int ry = y < p.Hs - ws ? 0 : (y < p.Hs - p.shift ? 1 : 2);
      const int rx = x < p.Ws - ws ? 0 : (x < p.Ws - p.shift ? 1 : 2);
      lab = ry * 3 + rx;
    }
    slab[t] = lab;
    scol[t] = ty * tw + tx;
  }
  __syncthreads();
  mbar_wait(bar, 0);

  const int g = lane >> 2, t4 = lane & 3;
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;  // this thread's two query rows
  // ---- Q fragments (A operand): 2 k-steps x {hi, lo}
  uint32_t qh[2][4], ql[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int c = ks * 16 + 2 * t4;
    const float2 a0 = *reinterpret_cast<const float2*>(sw_at(sq, r_lo, c));
    const float2 a1 = *reinterpret_cast<const float2*>(sw_at(sq, r_hi, c));
    const float2 a2 = *reinterpret_cast<const float2*>(sw_at(sq, r_lo, c + 8));
    const float2 a3 = *reinterpret_cast<const float2*>(sw_at(sq, r_hi, c + 8));
    split2(a0.x, a0.y, qh[ks][0], ql[ks][0]);
    split2(a1.x, a1.y, qh[ks][1], ql[ks][1]);
    split2(a2.x, a2.y, qh[ks][2], ql[ks][2]);
    split2(a3.x, a3.y, qh[ks][3], ql[ks][3]);
  }
  // ---- S = Q K^T: 18 key tiles x 2 k-steps x 3 products
  float s[WA_NT][4];
#pragma unroll
  for (int nt = 0; nt < WA_NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const int j = nt * 8 + g;  // key row of this thread's B fragment
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + 2 * t4;
      const float2 k0 = *reinterpret_cast<const float2*>(sw_at(sk, j, c));
      const float2 k1 = *reinterpret_cast<const float2*>(sw_at(sk, j, c + 8));
      uint32_t bh0, bl0, bh1, bl1;
      split2(k0.x, k0.y, bh0, bl0);
      split2(k1.x, k1.y, bh1, bl1);
      mma_bf16_16816(s[nt], ql[ks], bh0, bh1);
      mma_bf16_16816(s[nt], qh[ks], bl0, bl1);
      mma_bf16_16816(s[nt], qh[ks], bh0, bh1);
    }
  }
  // ---- scale, relative-position bias, shift mask, exact softmax over the 144 keys of rows r_lo / r_hi
  const float div = sqrtf(32.f);
  const int base_lo = (r_lo / ws + ws - 1) * tw + (r_lo % ws + ws - 1), base_hi = (r_hi / ws + ws - 1) * tw + (r_hi % ws + ws - 1);
  const int lab_lo = slab[r_lo], lab_hi = slab[r_hi];
  float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < WA_NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + 2 * t4 + e;
      const int cj = scol[j], lj = slab[j];
      float a = s[nt][e] / div + stab[base_lo - cj];
      float b = s[nt][2 + e] / div + stab[base_hi - cj];
      if (lj != lab_lo) a += -100.0f;
      if (lj != lab_hi) b += -100.0f;
      s[nt][e] = a;
      s[nt][2 + e] = b;
      m_lo = fmaxf(m_lo, a);
      m_hi = fmaxf(m_hi, b);
    }
  }
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < WA_NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[nt][e] = expf(s[nt][e] - m_lo);
      s[nt][2 + e] = expf(s[nt][2 + e] - m_hi);
      sum_lo += s[nt][e];
      sum_hi += s[nt][2 + e];
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
  // ---- O = P V: 9 k-steps of 16 keys, 4 output tiles of 8 dims, 3 products; P fragments come from the S accumulators
  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < WA_T / 16; ++kk) {
    uint32_t ph[4], pl[4];
    split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
    split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
    split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
    split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
    const int j0 = kk * 16 + 2 * t4;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int d = nt * 8 + g;
      uint32_t bh0, bl0, bh1, bl1;
      split2(*sw_at(sv, j0, d), *sw_at(sv, j0 + 1, d), bh0, bl0);
      split2(*sw_at(sv, j0 + 8, d), *sw_at(sv, j0 + 9, d), bh1, bl1);
      mma_bf16_16816(o[nt], pl, bh0, bh1);
      mma_bf16_16816(o[nt], ph, bl0, bl1);
      mma_bf16_16816(o[nt], ph, bh0, bh1);
    }
  }
  // ---- normalise and write the context planes
  const float inv_lo = 1.f / sum_lo, inv_hi = 1.f / sum_hi;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int d = h * 32 + nt * 8 + 2 * t4;
    uint32_t hh, ll;
    split2(o[nt][0] * inv_lo, o[nt][1] * inv_lo, hh, ll);
    *reinterpret_cast<uint32_t*>(p.out_hi + (row0 + r_lo) * C + d) = hh;
    if (p.out_lo) *reinterpret_cast<uint32_t*>(p.out_lo + (row0 + r_lo) * C + d) = ll;
    split2(o[nt][2] * inv_hi, o[nt][3] * inv_hi, hh, ll);
    *reinterpret_cast<uint32_t*>(p.out_hi + (row0 + r_hi) * C + d) = hh;
    if (p.out_lo) *reinterpret_cast<uint32_t*>(p.out_lo + (row0 + r_hi) * C + d) = ll;
  }
}

void launch_window_attn(cudaStream_t st, const float* qkv, const float* table, int64_t n_windows, int C, int heads,
                        int ws, int Hs, int Ws, int shift, Planes out) {
  MG_REQUIRE(C / heads == 32, "Swin head_dim must be 32");
  MG_REQUIRE(n_windows * heads < (1ll << 31), "too many windows");
  static const bool env_fma = getenv("MG_SWIN_ATTN") && std::string(getenv("MG_SWIN_ATTN")) == "fma";  // A/B switch
  if (ws * ws == WA_T && out.hi && !env_fma) {  // window 12: the tensor-core kernel
    WinAttnParams p;
    make_tmap_f32_2d(&p.tq, qkv, 3 * C, n_windows * WA_T, 3 * C, 32, WA_T);
    p.table = table; p.out_hi = out.hi; p.out_lo = out.lo;
    p.C = C; p.heads = heads; p.ws = ws; p.Hs = Hs; p.Ws = Ws; p.shift = shift;
    const size_t smem = 1024 + (size_t)(3 * WA_T * 32 + 23 * 23) * 4 + 2 * WA_T * 4 + 64;
    static bool attr2 = false;
    if (!attr2) {
      MG_CHECK_CUDA(cudaFuncSetAttribute(window_attn_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      MG_CHECK_CUDA(cudaFuncSetAttribute(window_attn_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr2 = true;
    }
    static const int minb = getenv("MG_SWIN_MINB") ? atoi(getenv("MG_SWIN_MINB")) : 2;
    if (minb >= 2)
      window_attn_mma_kernel<2><<<(unsigned)(n_windows * heads), WA_WARPS * 32, smem, st>>>(p);
    else
      window_attn_mma_kernel<1><<<(unsigned)(n_windows * heads), WA_WARPS * 32, smem, st>>>(p);
    MG_CHECK_CUDA(cudaGetLastError());
    return;
  }
  MG_REQUIRE(ws * ws <= 160, "Swin window too large for the attention kernel");
  const int T = ws * ws;
  const size_t smem = (size_t)(2 * T * 32 + (2 * ws - 1) * (2 * ws - 1)) * sizeof(float) + T * sizeof(int);
  static bool attr = false;
  if (!attr) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(window_attn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  MG_REQUIRE(n_windows * heads < (1ll << 31), "too many windows");
  window_attn_kernel<32><<<(unsigned)(n_windows * heads), 160, smem, st>>>(qkv, table, C, heads, ws, Hs, Ws, shift,
                                                                         out.hi, out.lo);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
