// extern "C" unit-level entry points (mg_op_*): thin wrappers that let the parity tests drive single
// kernels through the C ABI with plain device pointers. Declared in include/mg_b200.h.
#include <string>

#include "kernels.h"
#include "mg_b200.h"

namespace mg {
thread_local std::string g_last_error;
int set_error(const Error& e) {
  g_last_error = e.what();
  return e.code;
}
int set_error(const std::exception& e) {
  g_last_error = e.what();
  return -100;
}
}  // namespace mg

#define MG_API_BEGIN try {
#define MG_API_END                                      \
  return 0;                                             \
  }                                                     \
  catch (const mg::Error& e) { return mg::set_error(e); } \
  catch (const std::exception& e) { return mg::set_error(e); }

extern "C" {

const char* mg_last_error(void) { return mg::g_last_error.c_str(); }

int mg_op_gemm(void* stream, int M, int N, int K, const float* a, const float* b, float* c, const float* bias,
               const float* residual, int act, int planes, int block_n, int ksplit, int swap_out) {
  MG_API_BEGIN
  using namespace mg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t ldk = (K + 7) / 8 * 8;
  bf16 *ah, *al, *bh, *bl;
  MG_CHECK_CUDA(cudaMallocAsync(&ah, sizeof(bf16) * M * ldk, st));
  MG_CHECK_CUDA(cudaMallocAsync(&al, sizeof(bf16) * M * ldk, st));
  MG_CHECK_CUDA(cudaMallocAsync(&bh, sizeof(bf16) * N * ldk, st));
  MG_CHECK_CUDA(cudaMallocAsync(&bl, sizeof(bf16) * N * ldk, st));
  launch_split(st, a, M, K, K, Planes{ah, al}, ldk);
  launch_split(st, b, N, K, K, Planes{bh, bl}, ldk);
  GemmOperand A, B;
  A.hi = ah; A.lo = planes == 2 ? al : nullptr; A.rows = M; A.ld = ldk;
  B.hi = bh; B.lo = planes == 2 ? bl : nullptr; B.rows = N; B.ld = ldk;
  GemmEpilogue ep;
  ep.out_f32 = c;
  if (swap_out) { ep.ld_r = 1; ep.ld_c = M; } else { ep.ld_r = N; ep.ld_c = 1; }
  ep.bias = bias;
  ep.act = act;
  if (ksplit > 1) {
    // accumulate into c: initialise it with the residual (or zero) first
    if (residual) MG_CHECK_CUDA(cudaMemcpyAsync(c, residual, sizeof(float) * M * N, cudaMemcpyDeviceToDevice, st));
    else MG_CHECK_CUDA(cudaMemsetAsync(c, 0, sizeof(float) * M * N, st));
    ep.atomic = 1;
  } else {
    ep.residual = residual;
  }
  launch_gemm(st, A, B, M, N, K, 1, 1, ksplit, ep, block_n);
  MG_CHECK_CUDA(cudaFreeAsync(ah, st));
  MG_CHECK_CUDA(cudaFreeAsync(al, st));
  MG_CHECK_CUDA(cudaFreeAsync(bh, st));
  MG_CHECK_CUDA(cudaFreeAsync(bl, st));
  MG_API_END
}

int mg_op_kv24_roundtrip(void* stream, int B, int H, int Mp, const float* kt, const float* v, float* out) {
  MG_API_BEGIN
  using namespace mg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MG_REQUIRE(B > 0 && H > 0 && Mp > 0 && Mp % 8 == 0, "kv24: Mp must be a positive multiple of 8");
  uint8_t* packed;
  MG_CHECK_CUDA(cudaMallocAsync(&packed, (size_t)B * H * 384 * Mp, st));
  launch_kv24_roundtrip(st, kt, v, B, H, Mp, packed, out);
  MG_CHECK_CUDA(cudaFreeAsync(packed, st));
  MG_API_END
}

int mg_pack_pixels(void* stream, int B, int Hin, int Win, const uint8_t* src, int Hout, int Wout, int filter,
                   const float* mean3_host, const float* std3_host, float* out) {
  MG_API_BEGIN
  using namespace mg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MG_REQUIRE(src && out && mean3_host && std3_host, "mg_pack_pixels: null pointer");
  uint8_t* tmp = nullptr;
  if (Win != Wout) MG_CHECK_CUDA(cudaMallocAsync(&tmp, (size_t)B * Hin * Wout * 3, st));
  launch_pack_pixels(st, B, Hin, Win, src, Hout, Wout, filter, mean3_host, std3_host, tmp, out);
  if (tmp) MG_CHECK_CUDA(cudaFreeAsync(tmp, st));
  MG_API_END
}

int mg_resample_coeffs(int in_size, int out_size, int filter, int32_t* ksize_out, int32_t* bounds_host, int32_t* kk_host,
                       int kk_capacity) {
  MG_API_BEGIN
  using namespace mg;
  std::vector<int> b, k;
  const int ksize = resample_coeffs(in_size, out_size, filter, b, k);
  if (ksize_out) *ksize_out = ksize;
  if (bounds_host)
    for (size_t i = 0; i < b.size(); ++i) bounds_host[i] = b[i];
  if (kk_host) {
    MG_REQUIRE((size_t)kk_capacity >= k.size(), "mg_resample_coeffs: kk_capacity too small (need out_size * ksize)");
    for (size_t i = 0; i < k.size(); ++i) kk_host[i] = k[i];
  }
  MG_API_END
}

// ---- batched detokeniser ------------------------------------------------------------------------------------
struct mg_detok {
  int vocab = 0;
  int* text_off = nullptr;
  uint8_t* text = nullptr;
  uint8_t* flags = nullptr;
  // scratch of the last measure call
  int *tok_off = nullptr, *tok_len = nullptr;
  int64_t *row_len = nullptr, *row_off = nullptr;
  int cap_B = 0, cap_T = 0;
};

int mg_detok_create(int vocab, const uint8_t* text_host, const int32_t* text_off_host, const uint8_t* flags_host,
                    mg_detok** out) {
  MG_API_BEGIN
  using namespace mg;
  MG_REQUIRE(vocab > 0 && text_off_host && flags_host && out, "mg_detok_create: bad arguments");
  mg_detok* d = new mg_detok();
  d->vocab = vocab;
  const size_t nbytes = (size_t)text_off_host[vocab];
  MG_CHECK_CUDA(cudaMalloc((void**)&d->text_off, sizeof(int) * (vocab + 1)));
  MG_CHECK_CUDA(cudaMalloc((void**)&d->text, std::max<size_t>(nbytes, 16)));
  MG_CHECK_CUDA(cudaMalloc((void**)&d->flags, vocab));
  MG_CHECK_CUDA(cudaMemcpy(d->text_off, text_off_host, sizeof(int) * (vocab + 1), cudaMemcpyHostToDevice));
  if (nbytes) MG_CHECK_CUDA(cudaMemcpy(d->text, text_host, nbytes, cudaMemcpyHostToDevice));
  MG_CHECK_CUDA(cudaMemcpy(d->flags, flags_host, vocab, cudaMemcpyHostToDevice));
  *out = d;
  MG_API_END
}

void mg_detok_destroy(mg_detok* d) {
  if (!d) return;
  cudaFree(d->text_off); cudaFree(d->text); cudaFree(d->flags);
  cudaFree(d->tok_off); cudaFree(d->tok_len); cudaFree(d->row_len); cudaFree(d->row_off);
  delete d;
}

int mg_detok_measure(mg_detok* d, void* stream, int B, int T, const int64_t* ids, const int32_t* lens,
                     int64_t* row_off_host) {
  MG_API_BEGIN
  using namespace mg;
  MG_REQUIRE(d && ids && B > 0 && T > 0 && row_off_host, "mg_detok_measure: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B > d->cap_B || T > d->cap_T) {
    cudaFree(d->tok_off); cudaFree(d->tok_len); cudaFree(d->row_len); cudaFree(d->row_off);
    d->cap_B = std::max(B, d->cap_B);
    d->cap_T = std::max(T, d->cap_T);
    MG_CHECK_CUDA(cudaMalloc((void**)&d->tok_off, sizeof(int) * (size_t)d->cap_B * d->cap_T));
    MG_CHECK_CUDA(cudaMalloc((void**)&d->tok_len, sizeof(int) * (size_t)d->cap_B * d->cap_T));
    MG_CHECK_CUDA(cudaMalloc((void**)&d->row_len, sizeof(int64_t) * d->cap_B));
    MG_CHECK_CUDA(cudaMalloc((void**)&d->row_off, sizeof(int64_t) * (d->cap_B + 1)));
  }
  launch_detok_measure(st, ids, B, T, lens, d->text_off, d->flags, d->vocab, d->tok_off, d->tok_len, d->row_len, d->row_off);
  MG_CHECK_CUDA(cudaMemcpyAsync(row_off_host, d->row_off, sizeof(int64_t) * (B + 1), cudaMemcpyDeviceToHost, st));
  MG_CHECK_CUDA(cudaStreamSynchronize(st));
  MG_API_END
}

int mg_detok_write(mg_detok* d, void* stream, int B, int T, const int64_t* ids, uint8_t* out) {
  MG_API_BEGIN
  using namespace mg;
  MG_REQUIRE(d && ids && out && B > 0 && T > 0 && B <= d->cap_B && T <= d->cap_T, "mg_detok_write: call mg_detok_measure first");
  launch_detok_write(static_cast<cudaStream_t>(stream), ids, B, T, d->text_off, d->text, d->vocab, d->tok_off, d->tok_len,
                     d->row_off, out);
  MG_API_END
}

}  // extern "C"
