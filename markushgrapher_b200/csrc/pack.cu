// GPU input packing: page image (uint8 RGB, any size) -> pixel_values (B,3,512,512) fp32, i.e.
//   row["page_image"].resize((512, 512), resample=Image.LANCZOS)            reference mdu_dataset.py:118
//   image processor: resize (identity at 512x512), * 1/255, (x - mean) / std   reference utils/common.py:34-42
// The resize is Pillow's two-pass separable resampler restated exactly (Pillow src/libImaging/Resample.c:
// precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc): coefficients computed in
// double on the host, rounded to 22-bit fixed point, integer accumulation, clip to uint8 after EACH pass -- so the
// result is bit-identical to PIL (tests/test_pack_*.py compare against PIL itself).  The normalisation uses explicitly
// rounded fp32 multiply / subtract / divide (no FMA contraction) to match torch's CPU arithmetic bit for bit.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "kernels.h"

namespace mg {

static constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

static double rs_bilinear(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}
static double rs_sinc(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
static double rs_lanczos(double x) {
  if (-3.0 <= x && x < 3.0) return rs_sinc(x) * rs_sinc(x / 3);
  return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the full-image box [0, in_size)
int resample_coeffs(int in_size, int out_size, int filter, std::vector<int>& bounds, std::vector<int>& kk) {
  MG_REQUIRE(in_size > 0 && out_size > 0, "resample: sizes must be positive");
  MG_REQUIRE(filter == 0 || filter == 1, "resample: filter 0 = BILINEAR, 1 = LANCZOS");
  double (*fn)(double) = filter == 0 ? rs_bilinear : rs_lanczos;
  const double fsupport = filter == 0 ? 1.0 : 3.0;
  const float in0 = 0.f, in1 = (float)in_size;
  double scale, filterscale;
  filterscale = scale = (double)(in1 - in0) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsupport * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = fn((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    for (x = 0; x < ksize; ++x) {
      const double pk = k[x];
      kk[(size_t)xx * ksize + x] = pk < 0 ? (int)(-0.5 + pk * (1 << RS_PRECISION_BITS)) : (int)(0.5 + pk * (1 << RS_PRECISION_BITS));
    }
    bounds[(size_t)xx * 2] = xmin;
    bounds[(size_t)xx * 2 + 1] = xmax;
  }
  return ksize;
}

__device__ __forceinline__ int rs_clip8(int v) {
  v >>= RS_PRECISION_BITS;  // arithmetic shift, like Pillow's table lookup on (in >> PRECISION_BITS)
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: src [B][Hin][Win][3] u8 -> tmp [B][Hin][Wout][3] u8
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, int Hin, int Win, int Wout, const int* __restrict__ bounds,
                                  const int* __restrict__ kk, int ksize, uint8_t* __restrict__ tmp, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % Wout);
    const int64_t row = i / Wout;  // b * Hin + y
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int* k = kk + (int64_t)xx * ksize;
    const uint8_t* s = src + (row * Win + xmin) * 3;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
      const int w = k[x];
      s0 += s[3 * x] * w;
      s1 += s[3 * x + 1] * w;
      s2 += s[3 * x + 2] * w;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = (uint8_t)rs_clip8(s0);
    o[1] = (uint8_t)rs_clip8(s1);
    o[2] = (uint8_t)rs_clip8(s2);
  }
}

// vertical pass (or plain copy when Hin == Hout) fused with the image processor's rescale + normalise:
// tmp [B][Hin][W][3] u8 -> out [B][3][Hout][W] fp32
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int Hin, int W, int Hout,
                                       const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, float m0,
                                       float m1, float m2, float d0, float d1, float d2, float* __restrict__ out,
                                       int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const int yy = (int)((i / W) % Hout);
    const int64_t b = i / ((int64_t)W * Hout);
    int c0, c1, c2;
    if (kk) {
      const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
      const int* k = kk + (int64_t)yy * ksize;
      const uint8_t* s = tmp + ((b * Hin + ymin) * W + xx) * 3;
      int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < ymax; ++y) {
        const int w = k[y];
        const uint8_t* q = s + (int64_t)y * W * 3;
        s0 += q[0] * w;
        s1 += q[1] * w;
        s2 += q[2] * w;
      }
      c0 = rs_clip8(s0); c1 = rs_clip8(s1); c2 = rs_clip8(s2);
    } else {
      const uint8_t* q = tmp + ((b * Hin + yy) * W + xx) * 3;
      c0 = q[0]; c1 = q[1]; c2 = q[2];
    }
    const float inv255 = 1.0f / 255.0f;  // torch: t.float() * (1.0 / 255.0), then (t - mean) / std, each op rounded
    const int64_t plane = (int64_t)Hout * W;
    float* o = out + b * 3 * plane + (int64_t)yy * W + xx;
    o[0] = __fdiv_rn(__fsub_rn(__fmul_rn((float)c0, inv255), m0), d0);
    o[plane] = __fdiv_rn(__fsub_rn(__fmul_rn((float)c1, inv255), m1), d1);
    o[2 * plane] = __fdiv_rn(__fsub_rn(__fmul_rn((float)c2, inv255), m2), d2);
  }
}

// device copies of the coefficient tables, cached per (in, out, filter); the hot path re-uses a handful of shapes
struct DevCoeffs {
  int* bounds = nullptr;
  int* kk = nullptr;
  int ksize = 0;
};
static const DevCoeffs& dev_coeffs(int in_size, int out_size, int filter) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int, int>, DevCoeffs> cache;
  int dev = 0;
  MG_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple(dev, in_size, out_size, filter);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  std::vector<int> b, k;
  DevCoeffs d;
  d.ksize = resample_coeffs(in_size, out_size, filter, b, k);
  MG_CHECK_CUDA(cudaMalloc((void**)&d.bounds, b.size() * sizeof(int)));
  MG_CHECK_CUDA(cudaMalloc((void**)&d.kk, k.size() * sizeof(int)));
  MG_CHECK_CUDA(cudaMemcpy(d.bounds, b.data(), b.size() * sizeof(int), cudaMemcpyHostToDevice));
  MG_CHECK_CUDA(cudaMemcpy(d.kk, k.data(), k.size() * sizeof(int), cudaMemcpyHostToDevice));
  return cache.emplace(key, d).first->second;
}

void launch_pack_pixels(cudaStream_t st, int B, int Hin, int Win, const uint8_t* src, int Hout, int Wout, int filter,
                        const float* mean3, const float* std3, uint8_t* tmp, float* out) {
  MG_REQUIRE(B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "pack_pixels: empty image");
  const uint8_t* vin = src;
  if (Win != Wout) {  // Pillow: horizontal pass first, only when the width changes
    const DevCoeffs& ch = dev_coeffs(Win, Wout, filter);
    const int64_t total = (int64_t)B * Hin * Wout;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    resample_h_kernel<<<blocks, 256, 0, st>>>(src, Hin, Win, Wout, ch.bounds, ch.kk, ch.ksize, tmp, total);
    MG_CHECK_CUDA(cudaGetLastError());
    vin = tmp;
  }
  const int64_t total = (int64_t)B * Hout * Wout;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  if (Hin != Hout) {
    const DevCoeffs& cv = dev_coeffs(Hin, Hout, filter);
    resample_v_norm_kernel<<<blocks, 256, 0, st>>>(vin, Hin, Wout, Hout, cv.bounds, cv.kk, cv.ksize, mean3[0], mean3[1],
                                                   mean3[2], std3[0], std3[1], std3[2], out, total);
  } else {
    resample_v_norm_kernel<<<blocks, 256, 0, st>>>(vin, Hin, Wout, Hout, nullptr, nullptr, 0, mean3[0], mean3[1], mean3[2],
                                                   std3[0], std3[1], std3[2], out, total);
  }
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
