// OCSR branch (MolScribe Swin-B) non-GEMM kernels: bilinear resize, LayerNorm with window-partition /
// cyclic-shift / patch-merge gathers folded into the row index, and shifted-window attention.
// Arithmetic follows transformers/models/swin/modeling_swin.py (SwinLayer :534-653, SwinSelfAttention
// :383-459, SwinPatchMerging :298-350), the stock restatement of timm-0.4.12 swin_base_patch4_window12_384
// that MolScribe's encoder wraps (reference requirements.txt:25).
#include <algorithm>

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
// K and V of the window sit in shared memory (staged with cp.async.bulk + mbarrier when aligned);
// each thread owns one query row and runs an online softmax over the ws*ws keys in fp32.
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

void launch_window_attn(cudaStream_t st, const float* qkv, const float* table, int64_t n_windows, int C, int heads,
                        int ws, int Hs, int Ws, int shift, Planes out) {
  MG_REQUIRE(C / heads == 32, "Swin head_dim must be 32");
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
