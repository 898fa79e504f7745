// Non-GEMM kernels of the VTL (UDOP) encoder and shared helpers: RMSNorm, embedding fusion/compaction,
// cell-2D embeddings, relative-position buckets, bias+mask+softmax. All fp32 math; GEMM inputs are emitted
// directly as split-bf16 planes so no separate cast pass exists.
#include <math.h>

#include <algorithm>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

// =====================================================================================================
// T5 / UDOP bucket function as an exact integer LUT (host).  Follows
// transformers/models/udop/modeling_udop.py:466-512 (_relative_position_bucket), fp32 op order kept:
//   large = max_exact + (log(n.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
// lut[n] for n in [0, n_entries): bucket of |relative_position| = n WITHOUT the bidirectional sign offset.
void rel_bucket_lut(int bidirectional, int num_buckets, int max_distance, int n_entries, int32_t* lut) {
  int nb = bidirectional ? num_buckets / 2 : num_buckets;
  const int max_exact = nb / 2;
  const float denom = (float)log((double)max_distance / (double)max_exact);
  for (int n = 0; n < n_entries; ++n) {
    int v;
    if (n < max_exact) {
      v = n;
    } else {
      const float f = logf((float)n / (float)max_exact) / denom * (float)(nb - max_exact);
      long l = (long)f;  // trunc toward zero like .to(torch.long)
      l += max_exact;
      if (l > nb - 1) l = nb - 1;
      v = (int)l;
    }
    lut[n] = v;
  }
}

// =====================================================================================================
// RMSNorm (UdopLayerNorm, modeling_udop.py:333-355): y = w * (x * rsqrt(mean(x^2) + eps)) [* scale]
// one warp per row; D % 128 == 0 (float4 per lane).
template <int MAXV>  // float4 chunks per lane held in registers
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t rows, int D, float eps,
                               float scale, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                               float* __restrict__ out_f32, int64_t rows_per_b, int64_t out_bs, int64_t out_off) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = D / 128;
  const float4* xr = reinterpret_cast<const float4*>(x + r * D);
  float4 v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      v[i] = xr[lane + 32 * i];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  const float rs = rsqrtf(ss / (float)D + eps);
  const float4* wr = reinterpret_cast<const float4*>(w);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const float4 ww = wr[lane + 32 * i];
      float o[4] = {ww.x * (v[i].x * rs), ww.y * (v[i].y * rs), ww.z * (v[i].z * rs), ww.w * (v[i].w * rs)};
      if (scale != 1.f) {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] *= scale;
      }
      const int c = (lane + 32 * i) * 4;
      if (out_hi) {
        bf16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16(o[k], h[k], l[k]);
        uint2 ph, pl;
        ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
        pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
        pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
        *reinterpret_cast<uint2*>(out_hi + r * D + c) = ph;
        if (out_lo) *reinterpret_cast<uint2*>(out_lo + r * D + c) = pl;
      }
      if (out_f32) {
        const int64_t o_row = (r / rows_per_b) * out_bs + (r % rows_per_b) * D + out_off;
        *reinterpret_cast<float4*>(out_f32 + o_row + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

void launch_rmsnorm(cudaStream_t st, const float* x, const float* w, int64_t rows, int D, float eps, float scale,
                    Planes out, float* out_f32, int64_t rows_per_b, int64_t out_bs, int64_t out_off) {
  MG_REQUIRE(D % 128 == 0 && D <= 128 * 8, "rmsnorm: d_model must be a multiple of 128 and <= 1024");
  if (rows == 0) return;
  if (rows_per_b <= 0) {
    rows_per_b = rows;
    out_bs = 0;
  }
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  rmsnorm_kernel<8><<<grid, wpb * 32, 0, st>>>(x, w, rows, D, eps, scale, out.hi, out.lo, out_f32, rows_per_b, out_bs,
                                               out_off);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// im2col for a stride==kernel patch conv (UdopPatchEmbeddings modeling_udop.py:217-243; SwinPatchEmbeddings
// modeling_swin.py:255-296): pixel (B,3,H,W) fp32 -> planes [B*gh*gw, ldk], column order (c, ky, kx) = the
// flattening of the conv weight [out, 3, p, p]; columns >= 3*p*p are zero.
__global__ void im2col_kernel(const float* __restrict__ px, int B, int H, int W, int p, int ldk, bf16* hi, bf16* lo) {
  const int gh = H / p, gw = W / p;
  const int K = 3 * p * p;
  const int64_t total = (int64_t)B * gh * gw * ldk;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % ldk);
    const int64_t row = i / ldk;
    float v = 0.f;
    if (col < K) {
      const int kx = col % p, ky = (col / p) % p, c = col / (p * p);
      const int gx = (int)(row % gw), gy = (int)((row / gw) % gh), b = (int)(row / ((int64_t)gw * gh));
      v = px[(((int64_t)b * 3 + c) * H + (gy * p + ky)) * W + gx * p + kx];
    }
    bf16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

void launch_im2col(cudaStream_t st, const float* px, int B, int H, int W, int p, int ldk, Planes out) {
  const int64_t total = (int64_t)B * (H / p) * (W / p) * ldk;
  if (!total) return;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  im2col_kernel<<<blocks, 256, 0, st>>>(px, B, H, W, p, ldk, out.hi, out.lo);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// combine_image_text_embeddings (modeling_udop.py:133-214), step 1: per image, which patches are absorbed
// by a text token and the compacted (stable-order) list of surviving patches.
//   ocr_point = clip(floor((x0+x2)/2*np),0,np-1) + np*clip(floor((y0+y2)/2*np),0,np-1)     (fp32, :149-156)
//   every text token (masked or not, "target" or not) removes its patch                       (:169-181)
// out: ocr_point[B,Lt] (int), vis_src[B,NP] = source patch of the k-th surviving patch or -1, n_vis[B]
__global__ void combine_plan_kernel(const float* __restrict__ bbox, int Lt, int np, int* __restrict__ ocr_point,
                                    int* __restrict__ vis_src, int* __restrict__ n_vis) {
  extern __shared__ int sm[];  // removed[NP], then scan
  const int b = blockIdx.x;
  const int NP = np * np;
  int* removed = sm;
  for (int i = threadIdx.x; i < NP; i += blockDim.x) removed[i] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < Lt; t += blockDim.x) {
    const float* bb = bbox + ((int64_t)b * Lt + t) * 4;
    const float fx = floorf((bb[0] + bb[2]) / 2.0f * (float)np);
    const float fy = floorf((bb[1] + bb[3]) / 2.0f * (float)np);
    long long ix = (long long)fx, iy = (long long)fy;
    ix = ix < 0 ? 0 : (ix > np - 1 ? np - 1 : ix);
    iy = iy < 0 ? 0 : (iy > np - 1 ? np - 1 : iy);
    const int pt = (int)(ix + iy * np);
    ocr_point[(int64_t)b * Lt + t] = pt;
    removed[pt] = 1;  // benign race: all writers store 1
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // NP <= 1024: a serial stable compaction per image is negligible next to the encoder
    int k = 0;
    for (int i = 0; i < NP; ++i)
      if (!removed[i]) vis_src[(int64_t)b * NP + (k++)] = i;
    n_vis[b] = k;
    for (; k < NP; ++k) vis_src[(int64_t)b * NP + k] = -1;
  }
}

// step 2: build the encoder input sequence  x[B,Sp,D], bbox_ext[B,Sp,4] (float64 like the reference after
// :158), mask[B,Sp]:   s <  Lt            : tok_emb[id] + (target_seg ? 0 : patch_emb[ocr_point])     (:159-164)
//                      Lt <= s < Lt+NP    : k-th surviving patch embedding / zero padding              (:183-206)
//                      s >= Lt+NP         : alignment padding (mask 0), not part of the reference sequence
// then x += cell_2d_embedding(bbox_ext)  (UdopCellEmbeddings :784-807, UdopStack :1138-1139).
__global__ void combine_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ bbox,
                                     const int64_t* __restrict__ attn_mask, const float* __restrict__ tok_emb,
                                     const float* __restrict__ patch_emb, const int* __restrict__ ocr_point,
                                     const int* __restrict__ vis_src, const float* __restrict__ cell_x,
                                     const float* __restrict__ cell_y, int Lt, int np, int Sp, int D, int max_2d,
                                     int vocab, float* __restrict__ x, double* __restrict__ bbox_ext,
                                     int* __restrict__ mask) {
  const int s = blockIdx.x, b = blockIdx.y;
  const int NP = np * np;
  double bb[4] = {0., 0., 0., 0.};
  const float* src_tok = nullptr;
  const float* src_patch = nullptr;
  int m = 0;
  if (s < Lt) {
    const int64_t tix = (int64_t)b * Lt + s;
    for (int k = 0; k < 4; ++k) bb[k] = (double)bbox[tix * 4 + k];
    int64_t id = ids[tix];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    src_tok = tok_emb + id * D;
    const double mean = (bb[0] + bb[1] + bb[2] + bb[3]) / 4.0;
    const bool target = (mean == 0.0) || (mean == 1.0);
    if (!target) src_patch = patch_emb + ((int64_t)b * NP + ocr_point[tix]) * D;
    m = attn_mask ? (attn_mask[tix] != 0) : 1;
  } else if (s < Lt + NP) {
    const int src = vis_src[(int64_t)b * NP + (s - Lt)];
    if (src >= 0) {
      src_patch = patch_emb + ((int64_t)b * NP + src) * D;
      const int gx = src % np, gy = src / np;
      // get_visual_bbox (:94-117): fp32 k/np, promoted to float64 by the cat with the text boxes
      bb[0] = (double)((float)gx / (float)np);
      bb[1] = (double)((float)gy / (float)np);
      bb[2] = (double)((float)(gx + 1) / (float)np);
      bb[3] = (double)((float)(gy + 1) / (float)np);
      m = 1;
    }
  }
  int ci[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double c = bb[k] < 0.0 ? 0.0 : (bb[k] > 1.0 ? 1.0 : bb[k]);
    ci[k] = (int)(long long)(c * (double)(max_2d - 1));
  }
  const int64_t row = (int64_t)b * Sp + s;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) bbox_ext[row * 4 + k] = bb[k];
    mask[row] = m;
  }
  const float4* t4 = reinterpret_cast<const float4*>(src_tok);
  const float4* p4 = reinterpret_cast<const float4*>(src_patch);
  const float4* cx0 = reinterpret_cast<const float4*>(cell_x + (int64_t)ci[0] * D);
  const float4* cy1 = reinterpret_cast<const float4*>(cell_y + (int64_t)ci[1] * D);
  const float4* cx2 = reinterpret_cast<const float4*>(cell_x + (int64_t)ci[2] * D);
  const float4* cy3 = reinterpret_cast<const float4*>(cell_y + (int64_t)ci[3] * D);
  float4* o4 = reinterpret_cast<float4*>(x + row * D);
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src_tok) {
      e = t4[c];
      if (src_patch) {
        const float4 q = p4[c];
        e.x += q.x; e.y += q.y; e.z += q.z; e.w += q.w;
      }
    } else if (src_patch) {
      e = p4[c];
    }
    const float4 a = cx0[c], bq = cy1[c], cq = cx2[c], dq = cy3[c];
    float4 cell;
    cell.x = ((a.x + bq.x) + cq.x) + dq.x;
    cell.y = ((a.y + bq.y) + cq.y) + dq.y;
    cell.z = ((a.z + bq.z) + cq.z) + dq.z;
    cell.w = ((a.w + bq.w) + cq.w) + dq.w;
    e.x += cell.x; e.y += cell.y; e.z += cell.z; e.w += cell.w;
    o4[c] = e;
  }
}

void launch_combine(cudaStream_t st, const int64_t* ids, const float* bbox, const int64_t* attn_mask,
                    const float* tok_emb, const float* patch_emb, const float* cell_x, const float* cell_y, int B,
                    int Lt, int np, int Sp, int D, int max_2d, int vocab, int* ocr_point, int* vis_src, int* n_vis,
                    float* x, double* bbox_ext, int* mask) {
  const int NP = np * np;
  MG_REQUIRE(D % 4 == 0, "d_model must be a multiple of 4");
  combine_plan_kernel<<<B, 256, NP * sizeof(int), st>>>(bbox, Lt, np, ocr_point, vis_src, n_vis);
  MG_CHECK_CUDA(cudaGetLastError());
  dim3 grid(Sp, B);
  combine_embed_kernel<<<grid, 128, 0, st>>>(ids, bbox, attn_mask, tok_emb, patch_emb, ocr_point, vis_src, cell_x,
                                             cell_y, Lt, np, Sp, D, max_2d, vocab, x, bbox_ext, mask);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Horizontal / vertical relative-position buckets (RelativePositionBiasHorizontal/Vertical :935-970,
// get_relative_position :887-895, bucket :466-512), computed ONCE per forward and shared by all layers/heads:
//   pos = (b0+b2)/2 (float64);  rel = ((pos_j - pos_i) * 100).long();  bucket = (rel>0)*16 + lut[min(|rel|,cap)]
// out hv[B,Sp,Sp] as uchar2 (h, v).
__global__ void relbucket_hv_kernel(const double* __restrict__ bbox_ext, int Sp, const int* __restrict__ lut_hv,
                                    int lut_n, int half_buckets, double scaling, uchar2* __restrict__ hv) {
  const int i = blockIdx.x, b = blockIdx.y;
  const double* bi = bbox_ext + ((int64_t)b * Sp + i) * 4;
  const double xi = (bi[0] + bi[2]) / 2.0, yi = (bi[1] + bi[3]) / 2.0;
  uchar2* out = hv + ((int64_t)b * Sp + i) * Sp;
  for (int j = threadIdx.x; j < Sp; j += blockDim.x) {
    const double* bj = bbox_ext + ((int64_t)b * Sp + j) * 4;
    const double xj = (bj[0] + bj[2]) / 2.0, yj = (bj[1] + bj[3]) / 2.0;
    long long rx = (long long)((xj - xi) * scaling);
    long long ry = (long long)((yj - yi) * scaling);
    const int ox = rx > 0 ? half_buckets : 0, oy = ry > 0 ? half_buckets : 0;
    rx = rx < 0 ? -rx : rx;
    ry = ry < 0 ? -ry : ry;
    if (rx > lut_n - 1) rx = lut_n - 1;
    if (ry > lut_n - 1) ry = lut_n - 1;
    out[j] = make_uchar2((unsigned char)(ox + lut_hv[rx]), (unsigned char)(oy + lut_hv[ry]));
  }
}

void launch_relbucket_hv(cudaStream_t st, const double* bbox_ext, int B, int Sp, const int* lut_hv, int lut_n,
                         int half_buckets, uchar2* hv) {
  dim3 grid(Sp, B);
  relbucket_hv_kernel<<<grid, 256, 0, st>>>(bbox_ext, Sp, lut_hv, lut_n, half_buckets, 100.0, hv);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Encoder attention probabilities: P = softmax(scores + (bias_v + (bias_h + bias_1d)) + (1-mask)*finfo.min)
// (UdopStack :1173-1190, RelativePositionBiasAggregated :973-989, UdopAttention :608-613).  The (B,H,S,S)
// bias tensor of the reference is never materialised: buckets come from `hv` (shared over heads/layers) and
// an integer LUT over j-i.  One warp per (b,h,i) row; output = split planes for the P*V GEMM.
#ifndef ENC_SM_THREADS
#define ENC_SM_THREADS 256
#define ENC_SM_MINB 2
#endif
template <int MAXQ>  // quads (4 consecutive keys) per lane
__global__ void __launch_bounds__(ENC_SM_THREADS, ENC_SM_MINB) enc_softmax_kernel(const float* __restrict__ scores, const uchar2* __restrict__ hv,
                                   const int* __restrict__ mask, const float* __restrict__ tab1d,
                                   const float* __restrict__ tabh, const float* __restrict__ tabv,
                                   const int* __restrict__ lut1d, int lut1d_n, int half_buckets, int nbuckets, int H,
                                   int Sp, bf16* __restrict__ p_hi, bf16* __restrict__ p_lo, int64_t n_rows) {
  extern __shared__ float smf[];
  // tables transposed to [head][bucket]: a warp works on ONE head at a time, so its lanes index by bucket only and
  // distinct buckets fall in distinct banks (the [bucket][head] layout of the weights is a 16-way conflict)
  float* t1 = smf;                    // [H*nbuckets]
  float* th = t1 + nbuckets * H;
  float* tv = th + nbuckets * H;
  int* l1 = reinterpret_cast<int*>(tv + nbuckets * H);  // [lut1d_n]
  for (int i = threadIdx.x; i < nbuckets * H; i += blockDim.x) {
    const int bk = i / H, hh = i % H;
    t1[hh * nbuckets + bk] = tab1d[i];
    th[hh * nbuckets + bk] = tabh[i];
    tv[hh * nbuckets + bk] = tabv[i];
  }
  for (int i = threadIdx.x; i < lut1d_n; i += blockDim.x) l1[i] = lut1d[i];
  __syncthreads();
  // One warp per (image, query row), looping over the heads: the three bucket ids and the mask bit of every key --
  // the part of the work that does not depend on the head -- are decoded ONCE into a packed word per key and reused
  // by all H heads (the kernel is instruction-bound: ~40 instructions per score when every head redid the decode).
  const int64_t rowbi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // b*Sp + i
  if (rowbi >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int i = (int)(rowbi % Sp);
  const int64_t b = rowbi / Sp;
  const uint2* hvr = reinterpret_cast<const uint2*>(hv + rowbi * Sp);   // 4 x uchar2
  const int4* mr = reinterpret_cast<const int4*>(mask + b * Sp);
  const int nq = Sp >> 2;  // Sp is a multiple of 8
  uint32_t pk[MAXQ][4];    // bh | bv << 8 | b1 << 16 | visible << 24
#pragma unroll
  for (int e = 0; e < MAXQ; ++e) {
    const int q = lane + 32 * e;
    if (q < nq) {
      const uint2 hq = __ldg(hvr + q);
      const int4 mq = __ldg(mr + q);
      const uint32_t hw[2] = {hq.x, hq.y};
      const int mm[4] = {mq.x, mq.y, mq.z, mq.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = q * 4 + k;
        const uint32_t pr = (hw[k >> 1] >> ((k & 1) * 16)) & 0xffffu;  // bh | bv << 8
        int rel = j - i;
        const int o1 = rel > 0 ? half_buckets : 0;
        rel = rel < 0 ? -rel : rel;
        if (rel > lut1d_n - 1) rel = lut1d_n - 1;
        pk[e][k] = pr | ((uint32_t)(o1 + l1[rel]) << 16) | (mm[k] ? (1u << 24) : 0u);
      }
    }
  }
  for (int h = 0; h < H; ++h) {
    const int64_t row = (b * H + h) * Sp + i;
    const float* t1h = t1 + h * nbuckets;
    const float* thh = th + h * nbuckets;
    const float* tvh = tv + h * nbuckets;
    const float4* sr = reinterpret_cast<const float4*>(scores + row * Sp);
    // issue every global load of the row before touching the data (memory-level parallelism)
    float4 v[MAXQ];
#pragma unroll
    for (int e = 0; e < MAXQ; ++e) {
      const int q = lane + 32 * e;
      if (q < nq) v[e] = __ldcs(sr + q);   // streamed once
    }
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < MAXQ; ++e) {
      const int q = lane + 32 * e;
      if (q < nq) {
        float* vv = reinterpret_cast<float*>(&v[e]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t w = pk[e][k];
          float bias = tvh[(w >> 8) & 0xffu] + (thh[w & 0xffu] + t1h[(w >> 16) & 0xffu]);
          bias = bias + ((w >> 24) ? 0.f : -3.4028234663852886e38f);
          vv[k] = vv[k] + bias;
          mx = fmaxf(mx, vv[k]);
        }
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < MAXQ; ++e) {
      const int q = lane + 32 * e;
      if (q < nq) {
        float* vv = reinterpret_cast<float*>(&v[e]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          vv[k] = expf(vv[k] - mx);
          sum += vv[k];
        }
      }
    }
    sum = warp_sum(sum);
    const float inv_sum = 1.f / sum;
    uint2* ph = reinterpret_cast<uint2*>(p_hi + row * Sp);
    uint2* pl = p_lo ? reinterpret_cast<uint2*>(p_lo + row * Sp) : nullptr;
#pragma unroll
    for (int e = 0; e < MAXQ; ++e) {
      const int q = lane + 32 * e;
      if (q < nq) {
        const float* vv = reinterpret_cast<const float*>(&v[e]);
        // fp32 -> bf16 hi + bf16 lo, two elements per conversion (same rounding as split_bf16)
        const float p0 = vv[0] * inv_sum, p1 = vv[1] * inv_sum, p2 = vv[2] * inv_sum, p3 = vv[3] * inv_sum;
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1), h23 = __floats2bfloat162_rn(p2, p3);
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(p0 - f01.x, p1 - f01.y);
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(p2 - f23.x, p3 - f23.y);
        uint2 oh, ol;
        oh.x = *reinterpret_cast<const uint32_t*>(&h01); oh.y = *reinterpret_cast<const uint32_t*>(&h23);
        ol.x = *reinterpret_cast<const uint32_t*>(&l01); ol.y = *reinterpret_cast<const uint32_t*>(&l23);
        ph[q] = oh;
        if (pl) pl[q] = ol;
      }
    }
  }
}

void launch_enc_softmax(cudaStream_t st, const float* scores, const uchar2* hv, const int* mask, const float* tab1d,
                        const float* tabh, const float* tabv, const int* lut1d, int lut1d_n, int half_buckets,
                        int nbuckets, int B, int H, int Sp, Planes P) {
  MG_REQUIRE(Sp % 8 == 0 && Sp <= 128 * 13, "encoder sequence must be a multiple of 8 and <= 1664");
  MG_REQUIRE(nbuckets <= 256, "too many relative-position buckets");
  const int64_t rows = (int64_t)B * Sp;  // one warp per (image, query row), all heads
  const int wpb = ENC_SM_THREADS / 32;
  const size_t smem = (size_t)3 * nbuckets * H * sizeof(float) + (size_t)lut1d_n * sizeof(int);
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  if (Sp <= 128 * 5)
    enc_softmax_kernel<5><<<grid, wpb * 32, smem, st>>>(scores, hv, mask, tab1d, tabh, tabv, lut1d, lut1d_n,
                                                        half_buckets, nbuckets, H, Sp, P.hi, P.lo, rows);
  else if (Sp <= 128 * 9)
    enc_softmax_kernel<9><<<grid, wpb * 32, smem, st>>>(scores, hv, mask, tab1d, tabh, tabv, lut1d, lut1d_n,
                                                        half_buckets, nbuckets, H, Sp, P.hi, P.lo, rows);
  else
    enc_softmax_kernel<13><<<grid, wpb * 32, smem, st>>>(scores, hv, mask, tab1d, tabh, tabv, lut1d, lut1d_n,
                                                         half_buckets, nbuckets, H, Sp, P.hi, P.lo, rows);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// small utilities
__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill_f32(cudaStream_t st, float* p, int64_t n, float v) {
  if (!n) return;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  fill_f32_kernel<<<blocks, 256, 0, st>>>(p, n, v);
  MG_CHECK_CUDA(cudaGetLastError());
}

// memory mask: [1]*n_sw + vtl mask (UDOP extended mask), padded positions 0.  out [B, Mp] int32
__global__ void build_mem_mask_kernel(const int* __restrict__ vtl_mask, int Sp, int S, int n_sw, int Mp, int* out) {
  const int b = blockIdx.y;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Mp; m += gridDim.x * blockDim.x) {
    int v = 0;
    if (m < n_sw)
      v = 1;
    else if (m - n_sw < S)
      v = vtl_mask[(int64_t)b * Sp + (m - n_sw)];
    out[(int64_t)b * Mp + m] = v;
  }
}
void launch_build_mem_mask(cudaStream_t st, const int* vtl_mask, int B, int Sp, int S, int n_sw, int Mp, int* out) {
  dim3 grid((Mp + 255) / 256, B);
  build_mem_mask_kernel<<<grid, 256, 0, st>>>(vtl_mask, Sp, S, n_sw, Mp, out);
  MG_CHECK_CUDA(cudaGetLastError());
}

// =====================================================================================================
// Decoder-side view of the encoder memory: masked positions (text padding, the padded tail behind the surviving
// patches) never contribute to cross-attention -- exp(finfo.min - max) is exactly 0 in fp32 -- so they are dropped
// BEFORE the cross K/V projection instead of being projected, stored and re-streamed once per generated token.
// Valid rows keep their order (a softmax over keys does not depend on it anyway) and every image is padded with masked
// rows to the batch maximum, so all decode kernels keep one uniform memory length.
// plan: one CTA per image, src[b][k] = k-th valid position (stable), n_valid[b].
__global__ void compact_plan_kernel(const int* __restrict__ mask, int Mp, int* __restrict__ src, int* __restrict__ n_valid) {
  __shared__ int s_cnt[256];
  const int b = blockIdx.x, t = threadIdx.x;
  const int per = (Mp + 255) / 256;
  const int lo = min(t * per, Mp), hi = min(lo + per, Mp);
  int c = 0;
  for (int i = lo; i < hi; ++i) c += mask[(int64_t)b * Mp + i] != 0;
  s_cnt[t] = c;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) {
      const int v = s_cnt[i];
      s_cnt[i] = run;
      run += v;
    }
    n_valid[b] = run;
  }
  __syncthreads();
  int k = s_cnt[t];
  for (int i = lo; i < hi; ++i)
    if (mask[(int64_t)b * Mp + i] != 0) src[(int64_t)b * Mp + k++] = i;
}
__global__ void compact_gather_kernel(const float* __restrict__ mem, const int* __restrict__ src,
                                      const int* __restrict__ n_valid, int Mp, int Mc, int D, float* __restrict__ out,
                                      int* __restrict__ mask_out) {
  const int k = blockIdx.x, b = blockIdx.y;
  const bool ok = k < n_valid[b];
  float4* o = reinterpret_cast<float4*>(out + ((int64_t)b * Mc + k) * D);
  if (ok) {
    const float4* in = reinterpret_cast<const float4*>(mem + ((int64_t)b * Mp + src[(int64_t)b * Mp + k]) * D);
    for (int c = threadIdx.x; c < D / 4; c += blockDim.x) o[c] = in[c];
  } else {
    for (int c = threadIdx.x; c < D / 4; c += blockDim.x) o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) mask_out[(int64_t)b * Mc + k] = ok ? 1 : 0;
}
void launch_compact_plan(cudaStream_t st, const int* mask, int B, int Mp, int* src, int* n_valid) {
  compact_plan_kernel<<<B, 256, 0, st>>>(mask, Mp, src, n_valid);
  MG_CHECK_CUDA(cudaGetLastError());
}
void launch_compact_gather(cudaStream_t st, const float* mem, const int* src, const int* n_valid, int B, int Mp, int Mc,
                           int D, float* out, int* mask_out) {
  MG_REQUIRE(D % 4 == 0, "compact_gather: D must be a multiple of 4");
  dim3 grid(Mc, B);
  compact_gather_kernel<<<grid, 128, 0, st>>>(mem, src, n_valid, Mp, Mc, D, out, mask_out);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
