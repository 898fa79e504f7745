// Internal (non-ABI) declarations shared by the .cu translation units of libmg_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <stdexcept>
#include <string>

namespace mg {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// error plumbing: internal code throws, the extern "C" layer catches and stores mg_last_error()
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define MG_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      throw mg::Error(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ +  \
                              ":" + std::to_string(__LINE__) + ")");                                     \
  } while (0)
#define MG_REQUIRE(cond, msg)                                                                            \
  do {                                                                                                   \
    if (!(cond)) throw mg::Error(-1, std::string(msg) + " [" #cond "] (" + __FILE__ + ":" +              \
                                         std::to_string(__LINE__) + ")");                                \
  } while (0)

// ------------------------------------------------------------------------------------------------
// "Split planes": an fp32 matrix stored as two bf16 matrices hi + lo (x ~= hi + lo, err ~2^-17).
// The tensor-core GEMM multiplies planes (hi*hi + hi*lo + lo*hi, fp32 accumulate) which restores
// near-fp32 accuracy from bf16 tcgen05.mma. lo == nullptr means plain bf16 (throughput policy).
struct Planes {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

// One GEMM operand: K-major matrix [rows, K] (row stride ld elements), optionally batched over the
// problem's two batch dimensions (use_b1/use_b2 = does this operand vary along that dimension).
struct GemmOperand {
  const bf16* hi = nullptr;
  const bf16* lo = nullptr;
  int64_t rows = 0;
  int64_t ld = 0;       // elements, multiple of 8
  int64_t bs1 = 0;      // element stride along batch dim 1 (multiple of 8), ignored if !use_b1
  int64_t bs2 = 0;
  bool use_b1 = false;
  bool use_b2 = false;
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU_ERF = 2 };

// Where and how the fp32 accumulator tile leaves the kernel. Element (m, n) of batch (b1, b2) goes to
//   off = b1*bs1 + b2*bs2 + row(m)*ld_r + n*ld_c,   row(m) = row_map ? row_map[m] : m
// Exactly one of {out_f32, out_hi(+out_lo)} is written. v = act(acc + bias) (+ residual[off]).
struct GemmEpilogue {
  float* out_f32 = nullptr;
  bf16* out_hi = nullptr;
  bf16* out_lo = nullptr;
  int64_t ld_r = 0, ld_c = 1;
  int64_t bs1 = 0, bs2 = 0;
  const float* bias = nullptr;
  int bias_on_rows = 0;          // bias indexed by m instead of n (swapped-operand GEMMs)
  const float* residual = nullptr;
  const int* row_map = nullptr;
  int act = ACT_NONE;
  int atomic = 0;                // red.global.add.f32 into out_f32 (split-K / residual accumulate)
};

// C[M,N] (+)= A[M,K] * B[N,K]^T on tcgen05 tensor cores. block_n in {32,64,128}; ksplit >= 1
// (ksplit > 1 requires ep.atomic). nb1/nb2: problem batch dims.
void launch_gemm(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int nb1, int nb2,
                 int ksplit, const GemmEpilogue& ep, int block_n);

// TMA descriptor (128-byte swizzle, box 64 x box_rows) of a K-major bf16 operand batched along bs1: coordinates
// (inner, row, batch, 0); rows past op.rows are zero-filled
void make_tmap_bf16(CUtensorMap* map, const bf16* ptr, const GemmOperand& op, int inner, int nb, int box_rows);

void make_tmap_f32_2d(CUtensorMap* map, const float* ptr, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows);

// fp32 [rows, cols] (row stride ld_in) -> planes [rows, ld_out], zero-filling cols..ld_out
void launch_split(cudaStream_t st, const float* in, int64_t rows, int64_t cols, int64_t ld_in, Planes out,
                  int64_t ld_out);

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may begin while its predecessor
// in the stream is still running and MUST call griddep_wait() before touching anything the predecessor produces
// (or still reads). Captured into CUDA graphs as programmatic dependency edges.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  static const bool pdl_on = !(getenv("MG_NO_PDL") && getenv("MG_NO_PDL")[0] == '1');  // A/B switch for profiling
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_on ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

}  // namespace mg
