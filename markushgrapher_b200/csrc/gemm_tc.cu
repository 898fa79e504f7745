// tcgen05 / TMA GEMM for sm_100a:  C[M,N] (+)= A[M,K] * B[N,K]^T, both operands K-major.
//
// * Operands are "split planes" (bf16 hi + bf16 lo, see mg_internal.h). With both planes present the
//   kernel issues three tcgen05.mma products per k-step (hi*hi + hi*lo + lo*hi) into one fp32 TMEM
//   accumulator, which recovers ~fp32 accuracy (error ~2^-16) from the bf16 tensor pipe. With only the
//   hi plane it is a plain bf16 GEMM.
// * Tiles: 128 (M) x BN (N) x 64 (K); TMA (cp.async.bulk.tensor.4d, 128-byte swizzle) stages A/B planes
//   into a 3-5 deep shared-memory ring guarded by full/empty mbarriers; one elected thread issues the MMAs
//   (UMMA 128 x BN x 16); the accumulator lives in TMEM and is read back with tcgen05.ld by four epilogue
//   warps that apply bias / activation / residual and write fp32, split planes, or atomically accumulate
//   (split-K and residual-stream accumulation for the skinny decode GEMMs).
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2-5 = epilogue.
//
// Replaces the cuBLAS calls under torch.nn.Linear / torch.matmul in the reference stack
// (transformers/models/udop/modeling_udop.py:465-468,590,610,617; models/swin/modeling_swin.py:403-405,424,447).
#include "mg_internal.h"
#include "ptx.cuh"

namespace mg {

struct alignas(64) GemmParams {
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int M, N, K;
  int nx, ny, nz;    // tile grid: n tiles, m tiles, ksplit * nb1 * nb2
  int nb1;           // z = ks + ksplit * (b1 + nb1 * b2)
  int ksplit, kb_per_split, num_kb;
  int a_use1, a_use2, b_use1, b_use2;
  int nplanes;       // 2 = hi+lo (3 products), 1 = hi only
  int vec_ok;        // output offsets are 4-element aligned -> vector path allowed
  GemmEpilogue ep;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 3 : (BN == 64 ? 4 : 5);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU_ERF) return v * 0.5f * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

template <int BN>
__global__ void __launch_bounds__(320, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  // Persistent: one CTA per SM walks the tile list (tile = blockIdx.x, += gridDim.x).  The shared-memory ring and its
  // barrier phases run on across tiles, and the fp32 accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so the
  // epilogue of tile i (TMEM -> registers -> global) overlaps the MMAs of tile i+1 and the per-tile prologue
  // (barrier init, TMEM allocation, tensor-map prefetch) is paid once per SM instead of once per tile.
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2], 256 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.nx * p.ny * p.nz;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.ta_hi);
    prefetch_tmap(&p.tb_hi);
    if (p.nplanes == 2) {
      prefetch_tmap(&p.ta_lo);
      prefetch_tmap(&p.tb_lo);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base0 = *tmem_slot;

  // tile id -> (n tile, m tile, k-split / batch indices); identical in all three roles
#define MG_TILE_DECODE(t)                                        \
  const int tx_ = (t) % p.nx, ty_ = ((t) / p.nx) % p.ny, tz_ = (t) / (p.nx * p.ny); \
  const int n0 = tx_ * BN, m0 = ty_ * BM;                      \
  const int ks = tz_ % p.ksplit, zb = tz_ / p.ksplit;          \
  const int b1 = zb % p.nb1, b2 = zb / p.nb1;                  \
  const int kb0 = ks * p.kb_per_split;                         \
  const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);

  {
  {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        const uint32_t tx = (p.nplanes == 2) ? C::STAGE_BYTES : (A_BYTES + C::B_BYTES);
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          MG_TILE_DECODE(tile)
          const int a1 = p.a_use1 ? b1 : 0, a2 = p.a_use2 ? b2 : 0;
          const int c1 = p.b_use1 ? b1 : 0, c2 = p.b_use2 ? b2 : 0;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + s * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[s], tx);
            tma_load_4d(st, &p.ta_hi, &full_bar[s], kb * BK, m0, a1, a2);
            tma_load_4d(st + 2 * A_BYTES, &p.tb_hi, &full_bar[s], kb * BK, n0, c1, c2);
            if (p.nplanes == 2) {
              tma_load_4d(st + A_BYTES, &p.ta_lo, &full_bar[s], kb * BK, m0, a1, a2);
              tma_load_4d(st + 2 * A_BYTES + C::B_BYTES, &p.tb_lo, &full_bar[s], kb * BK, n0, c1, c2);
            }
            if (++s == C::STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer
      // The whole warp runs the (warp-uniform) loop and one elected lane issues.  Under a divergent `if (lane == 0)`
      // the compiler cannot keep the descriptors in uniform registers and wraps every tcgen05.mma in an
      // ELECT / R2UR loop (~17 instructions, about as long as a 128x128x16 MMA runs): the issue loop, not the
      // tensor pipe, paced the kernel.  Now: ~2 uniform-datapath instructions per MMA.
      {
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
          MG_TILE_DECODE(tile)
          (void)n0; (void)m0; (void)b1; (void)b2;
          const int ab = it & 1;  // accumulator buffer
          mbar_wait(&tmem_empty_bar[ab], ((uint32_t)(it >> 1) & 1u) ^ 1u);  // its previous tile has been read out
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base0 + (uint32_t)(ab * BN);
          uint32_t acc = 0;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
            const uint64_t da_hi = make_sw128_kmajor_desc(sa);
            const uint64_t da_lo = make_sw128_kmajor_desc(sa + A_BYTES);
            const uint64_t db_hi = make_sw128_kmajor_desc(sa + 2 * A_BYTES);
            const uint64_t db_lo = make_sw128_kmajor_desc(sa + 2 * A_BYTES + C::B_BYTES);
            if (elect_one()) {
              if (p.nplanes == 2) {
                // small cross terms first, the dominant hi*hi product last
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, da_lo + 2 * k, db_hi + 2 * k, idesc, k == 0 ? acc : 1u);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, da_hi + 2 * k, db_lo + 2 * k, idesc, 1);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, da_hi + 2 * k, db_hi + 2 * k, idesc, 1);
              } else {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, da_hi + 2 * k, db_hi + 2 * k, idesc, k == 0 ? acc : 1u);
              }
              umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
            }
            __syncwarp();
            acc = 1;
            if (++s == C::STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
          if (elect_one()) umma_commit(&tmem_full_bar[ab]);  // accumulator complete
          __syncwarp();
        }
      }
    } else {
      // ------------------------------------------------------------------ epilogue (warps 2..9)
      // two warps per TMEM lane quadrant, each taking half of the tile's columns: with K = 64 (attention scores)
      // a tile is one k-block of MMAs and 64 KB of output, and the epilogue's throughput is the kernel's
      const GemmEpilogue& ep = p.ep;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      MG_TILE_DECODE(tile)
      (void)kb0; (void)kb1;
      const int ab = it & 1;
      const uint32_t tmem_base = tmem_base0 + (uint32_t)(ab * BN);
      mbar_wait(&tmem_full_bar[ab], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const int q = warp & 3;  // TMEM lane quadrant this warp may access
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      const int64_t row = (row_ok && ep.row_map) ? ep.row_map[m] : m;
      const int64_t off_b = (int64_t)b1 * ep.bs1 + (int64_t)b2 * ep.bs2;
      const float bias_m = (row_ok && ep.bias && ep.bias_on_rows) ? ep.bias[m] : 0.f;
      constexpr int CHALF = (BN >= 64) ? BN / 2 : BN;        // columns per warp of a quadrant pair
      const int c_lo = (BN >= 64) ? ((warp - 2) >> 2) * CHALF : 0;
      const int c_hi = (BN >= 64 || warp < 6) ? c_lo + CHALF : 0;  // BN = 32: the second warp of a pair has no columns
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores below
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int nvalid = min(32, p.N - (n0 + c0));
        if (ep.bias) {
          if (ep.bias_on_rows) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias_m;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) v[j] += __ldg(ep.bias + n0 + c0 + j);
          }
        }
        if (ep.act != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
        }
        const int64_t base = off_b + row * ep.ld_r + (int64_t)(n0 + c0) * ep.ld_c;
        if (ep.ld_c == 1 && nvalid == 32 && p.vec_ok) {
          // each thread owns 32 contiguous outputs of its row
          if (ep.out_f32) {
            float4* dst = reinterpret_cast<float4*>(ep.out_f32 + base);
            if (ep.atomic) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                atomicAdd(dst + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            } else {
              if (ep.residual) {
                const float4* res = reinterpret_cast<const float4*>(ep.residual + base);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 t = res[j];
                  v[4 * j] += t.x;
                  v[4 * j + 1] += t.y;
                  v[4 * j + 2] += t.z;
                  v[4 * j + 3] += t.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          } else {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              bf16 h0, l0, h1, l1;
              split_bf16(v[2 * j], h0, l0);
              split_bf16(v[2 * j + 1], h1, l1);
              hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            uint4* dh = reinterpret_cast<uint4*>(ep.out_hi + base);
#pragma unroll
            for (int j = 0; j < 4; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            if (ep.out_lo) {
              uint4* dl = reinterpret_cast<uint4*>(ep.out_lo + base);
#pragma unroll
              for (int j = 0; j < 4; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
        } else {
          // generic path: scalar, coalesced across lanes when ld_r == 1 (swapped-operand GEMMs)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              const int64_t o = base + (int64_t)j * ep.ld_c;
              float x = v[j];
              if (ep.out_f32) {
                if (ep.atomic) {
                  atomicAdd(ep.out_f32 + o, x);
                } else {
                  if (ep.residual) x += ep.residual[o];
                  ep.out_f32[o] = x;
                }
              } else {
                bf16 h, l;
                split_bf16(x, h, l);
                ep.out_hi[o] = h;
                if (ep.out_lo) ep.out_lo[o] = l;
              }
            }
          }
        }
      }
      // every tcgen05.ld of this tile has completed (wait::ld above): hand the accumulator back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[ab]);
      }
    }
  }
  }
#undef MG_TILE_DECODE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base0, 2 * BN);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    MG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    MG_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static void encode_operand(CUtensorMap* map, const bf16* ptr, const GemmOperand& op, int K, int nb1, int nb2,
                           int box_rows) {
  MG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand must be 16-byte aligned");
  MG_REQUIRE(op.ld % 8 == 0, "GEMM operand row stride must be a multiple of 8 elements");
  const cuuint64_t d1 = op.use_b1 ? nb1 : 1, d2 = op.use_b2 ? nb2 : 1;
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)op.rows, d1, d2};
  const cuuint64_t dflt = (cuuint64_t)op.ld * 2 * (cuuint64_t)(op.rows > 0 ? op.rows : 1);
  cuuint64_t s1 = op.use_b1 ? (cuuint64_t)op.bs1 * 2 : dflt;
  cuuint64_t s2 = op.use_b2 ? (cuuint64_t)op.bs2 * 2 : dflt;
  MG_REQUIRE(s1 % 16 == 0 && s2 % 16 == 0, "GEMM operand batch strides must be multiples of 8 elements");
  cuuint64_t strides[3] = {(cuuint64_t)op.ld * 2, s1, s2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(-3, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (K=" +
                        std::to_string(K) + " rows=" + std::to_string(op.rows) + " ld=" + std::to_string(op.ld) + ")");
}

// TMA descriptor of a K-major bf16 operand [nb][rows][inner] (128-byte swizzle, box = 64 x box_rows) for kernels outside
// this file (enc_flash.cu); coordinates are (inner, row, batch, 0)
void make_tmap_bf16(CUtensorMap* map, const bf16* ptr, const GemmOperand& op, int inner, int nb, int box_rows) {
  encode_operand(map, ptr, op, inner, nb, 1, box_rows);
}

// TMA descriptor of an fp32 matrix [rows][cols] (row stride ld elements), 128-byte swizzle, box box_cols x box_rows
// (box_cols * 4 must be 128 bytes): coordinates (col, row); rows past the matrix are zero-filled
void make_tmap_f32_2d(CUtensorMap* map, const float* ptr, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows) {
  MG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 4 == 0, "fp32 TMA operand must be 16-byte aligned");
  MG_REQUIRE(box_cols * 4 == 128 && box_rows <= 256, "fp32 TMA box must be 128 bytes wide, <= 256 rows");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled (fp32) failed with CUresult " + std::to_string((int)r));
}

template <int BN>
static void launch_bn(cudaStream_t st, const GemmParams& p, dim3 grid) {
  static bool attr_set = false;
  if (!attr_set) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    attr_set = true;
  }
  gemm_tc_kernel<BN><<<grid, 320, Cfg<BN>::SMEM, st>>>(p);
  MG_CHECK_CUDA(cudaGetLastError());
}

void launch_gemm(cudaStream_t st, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int nb1, int nb2,
                 int ksplit, const GemmEpilogue& ep, int block_n) {
  MG_REQUIRE(M > 0 && N > 0 && K > 0 && nb1 > 0 && nb2 > 0, "empty GEMM");
  MG_REQUIRE(block_n == 32 || block_n == 64 || block_n == 128, "block_n must be 32, 64 or 128");
  MG_REQUIRE((A.lo != nullptr) == (B.lo != nullptr), "both operands must have the same number of planes");
  MG_REQUIRE(ksplit >= 1 && (ksplit == 1 || (ep.atomic && ep.out_f32)), "split-K needs the atomic fp32 epilogue");
  MG_REQUIRE((ep.out_f32 != nullptr) != (ep.out_hi != nullptr), "exactly one output kind");
  MG_REQUIRE(A.rows >= M && B.rows >= N, "operand has fewer rows than the problem");
  GemmParams p;
  p.nplanes = A.lo ? 2 : 1;
  encode_operand(&p.ta_hi, A.hi, A, K, nb1, nb2, BM);
  encode_operand(&p.tb_hi, B.hi, B, K, nb1, nb2, block_n);
  if (p.nplanes == 2) {
    encode_operand(&p.ta_lo, A.lo, A, K, nb1, nb2, BM);
    encode_operand(&p.tb_lo, B.lo, B, K, nb1, nb2, block_n);
  } else {
    p.ta_lo = p.ta_hi;
    p.tb_lo = p.tb_hi;
  }
  p.M = M;
  p.N = N;
  p.K = K;
  p.nb1 = nb1;
  p.num_kb = (K + BK - 1) / BK;
  if (ksplit > p.num_kb) ksplit = p.num_kb;
  p.kb_per_split = (p.num_kb + ksplit - 1) / ksplit;
  ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.ksplit = ksplit;
  p.a_use1 = A.use_b1;
  p.a_use2 = A.use_b2;
  p.b_use1 = B.use_b1;
  p.b_use2 = B.use_b2;
  p.ep = ep;
  const void* optr = ep.out_f32 ? (const void*)ep.out_f32 : (const void*)ep.out_hi;
  const int align_el = ep.out_f32 ? 4 : 8;  // 16-byte vectors
  p.vec_ok = (ep.ld_c == 1) && (ep.ld_r % align_el == 0) && (ep.bs1 % align_el == 0) && (ep.bs2 % align_el == 0) &&
             ((reinterpret_cast<uintptr_t>(optr) & 15) == 0) &&
             (!ep.out_lo || (reinterpret_cast<uintptr_t>(ep.out_lo) & 15) == 0) &&
             (!ep.residual || (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
  p.nx = (N + block_n - 1) / block_n;
  p.ny = (M + BM - 1) / BM;
  p.nz = ksplit * nb1 * nb2;
  const int64_t tiles = (int64_t)p.nx * p.ny * p.nz;
  MG_REQUIRE(tiles < (int64_t)1 << 30, "GEMM tile count too large");
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    MG_CHECK_CUDA(cudaGetDevice(&dev));
    MG_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  dim3 grid((unsigned)std::min<int64_t>(tiles, n_sm));  // persistent: one CTA per SM walks the tile list
  if (block_n == 128)
    launch_bn<128>(st, p, grid);
  else if (block_n == 64)
    launch_bn<64>(st, p, grid);
  else
    launch_bn<32>(st, p, grid);
}

// ================================================================================================
// Skinny (decode-step) linear on the tensor cores:  out[b][n] (+)= post_b * sum_k pro(x)[b][k] * W[n][k]
// for R = 32*NG activation rows.  W rows sit on the UMMA M axis (128 features per CTA, TMA-staged split planes),
// the activation tile is BUILT IN-KERNEL by the four worker warps: they read the fp32 residual stream / hidden
// buffer, apply the fused prologue (RMSNorm weight, ReLU), split to bf16 hi/lo and store straight into the
// 128B-swizzled K-major layout the UMMA descriptor expects -- so no separate norm / activation / cast kernels
// exist in the decode step.  The RMSNorm row statistic is applied in the epilogue (rs[b] * acc is linear in the
// split-K partial sums), and is computed by the workers while TMA and MMA are in flight.  Split-K over
// blockIdx.y with fp32 atomics keeps every SM streaming a slice of W; accumulating straight into the residual
// stream gives the residual add; an optional zero duty clears a later GEMM's accumulation buffer.
// (UdopLayerNorm :333-355, UdopDenseActDense :359-381, UdopAttention q/k/v/o :465-468, lm head :1585-1590)
struct alignas(64) SkinnyParams {
  CUtensorMap tw_hi, tw_lo;
  const float* x;
  int ldx;
  float* out;
  int ld_out;
  int B, N, K;
  int kb_per_cta, num_kb, stages;
  int nplanes;
  int pro;  // 0 none, 1 rms (x*lnw staged, rs*scale in the epilogue), 2 relu
  const float* lnw;
  float eps, scale;
  float* zero_ptr;
  long long zero_n;
  int store;
  // optional fused partial argmax over this CTA's 128 features, per activation row (LM head -> greedy selection):
  // amax_val/amax_idx [B][gridDim.x]; ties resolve to the lowest feature index (torch.argmax semantics)
  float* amax_val;
  int* amax_idx;
  const float* rs_in;  // pro == 1: row scales rsqrt(mean(x^2) + eps) * scale computed by rms_rowscale_kernel (null: in-kernel)
};

// RMSNorm row scales of up to 128 activation rows, one warp per row -- the same arithmetic, in the same order, as the
// statistic warps of skinny_tc_kernel.  With more than 32 rows every CTA of the linear would recompute all of them (16
// dependent round trips to L2 for 128 rows: +10 us on a 12 us kernel), so the wide launches get them from here.
__global__ void __launch_bounds__(256) rms_rowscale_kernel(const float* __restrict__ x, int ldx, int B, int K, float eps,
                                                           float scale, float* __restrict__ rs) {
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  griddep_wait();
  if (r >= B) return;
  const int n4row = K >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)r * ldx);
  float4 qa[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qa[i] = xr[min(lane + 32 * i, n4row - 1)];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = qa[i];
    ss += (lane + 32 * i < n4row) ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f;
  }
  ss = warp_sum(ss);
  if (lane == 0) rs[r] = rsqrtf(ss / (float)K + eps) * scale;
}

template <int NG>
__global__ void __launch_bounds__(320, 1) skinny_tc_kernel(const __grid_constant__ SkinnyParams p) {
  constexpr int R = 32 * NG;             // activation rows = UMMA N
  constexpr int X_BYTES = R * BK * 2;    // one plane of the activation tile
  constexpr int STAGE = 2 * A_BYTES + 2 * X_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * STAGE);   // W tile landed (TMA tx)
  uint64_t* xrdy_bar = full_bar + S;                                    // activation tile staged (128 arrivals)
  uint64_t* empty_bar = xrdy_bar + S;                                   // stage consumed by the MMAs
  uint64_t* tmem_full_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_rs = reinterpret_cast<float*>(tmem_slot + 2);                // [R] row scale for the epilogue
  float* s_part = s_rs + R;                                             // [R][4] partial sums of squares

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int kb0 = blockIdx.y * p.kb_per_cta;
  const int kb1 = min(p.num_kb, kb0 + p.kb_per_cta);
  griddep_launch();  // the next kernel may start prefetching its own weights now

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tw_hi);
    if (p.nplanes == 2) prefetch_tmap(&p.tw_lo);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xrdy_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, R);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (p.nplanes == 2) ? 2 * A_BYTES : A_BYTES;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE;
        mbar_expect_tx(&full_bar[s], tx);
        tma_load_4d(st, &p.tw_hi, &full_bar[s], kb * BK, m0, 0, 0);
        if (p.nplanes == 2) tma_load_4d(st + A_BYTES, &p.tw_lo, &full_bar[s], kb * BK, m0, 0, 0);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform loop, one elected lane issues (see gemm_tc_kernel)
      constexpr uint32_t idesc = make_idesc_bf16(BM, R);
      int s = 0;
      uint32_t ph = 0, acc = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        mbar_wait(&xrdy_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE);
        const uint64_t da_hi = make_sw128_kmajor_desc(sa), da_lo = make_sw128_kmajor_desc(sa + A_BYTES);
        const uint64_t db_hi = make_sw128_kmajor_desc(sa + 2 * A_BYTES);
        const uint64_t db_lo = make_sw128_kmajor_desc(sa + 2 * A_BYTES + X_BYTES);
        if (elect_one()) {
          if (p.nplanes == 2) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, da_lo + 2 * k, db_hi + 2 * k, idesc, k == 0 ? acc : 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_lo + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_hi + 2 * k, idesc, 1);
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_hi + 2 * k, idesc, k == 0 ? acc : 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
        acc = 1;
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else {
  // warps 2..9 meet at ONE named barrier (row scales published by the statistic warps before the workers' epilogue);
  // both groups reach it through the same instruction, which is also what compute-sanitizer synccheck expects
  if (warp >= 6) {
    // ---------------------------------------------------------------- statistic warps (6..9, 128 threads)
    // RMSNorm row statistic over the FULL row, computed concurrently with the activation staging of the worker
    // warps (it used to follow it: +2.2 us on the critical path of every RMSNorm-fused linear)
    const int t = threadIdx.x - 192;
    const int wq = warp - 6;  // 0..3
    griddep_wait();
    if (p.pro == 1 && p.rs_in) {
      if (t < R) s_rs[t] = p.rs_in[min(t, p.B - 1)];
    } else if (p.pro == 1) {
      // K <= 1024 here (d_model).  Warp wq owns rows wq, wq + 4, ...: a lane holds 8 float4 of a row, two rows (16
      // loads) are in flight per round trip to L2, and a row's sum never leaves its warp -- no block-level combine.
      const int n4row = p.K >> 2;
#pragma unroll 1
      for (int r0 = wq; r0 < R; r0 += 8) {
        float4 qa[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4* xr = reinterpret_cast<const float4*>(p.x + (int64_t)min(r0 + 4 * u, p.B - 1) * p.ldx);
#pragma unroll
          for (int i = 0; i < 8; ++i) qa[u][i] = xr[min(lane + 32 * i, n4row - 1)];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = qa[u][i];
            ss += (lane + 32 * i < n4row) ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f;
          }
          ss = warp_sum(ss);
          if (lane == 0 && r0 + 4 * u < R) s_rs[r0 + 4 * u] = rsqrtf(ss / (float)p.K + p.eps) * p.scale;
        }
      }
    } else {
      if (t < R) s_rs[t] = p.scale;
    }
  } else {
    // ---------------------------------------------------------------- workers (warps 2..5, 128 threads)
    const int t = threadIdx.x - 64;
    griddep_wait();  // everything below reads / accumulates into buffers owned by the preceding kernels
    {  // activation tiles for every k-block of this CTA, G k-blocks per round trip to L2
      constexpr int NI = (R * 16) / 128;  // float4 items per thread per k-block
      constexpr int G = 4 / NG;           // k-blocks whose loads are issued together
      const int c4 = t & 15;              // 16 float4 per 64-wide row; identical for all items of a thread
      const int nkb = kb1 - kb0;
      int s = 0;
      uint32_t ph = 0;
      for (int g0 = 0; g0 < nkb; g0 += G) {
        float4 v[G][NI];
        float4 gw[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int kb = kb0 + min(g0 + u, nkb - 1);
          if (p.pro == 1) gw[u] = *reinterpret_cast<const float4*>(p.lnw + kb * BK + c4 * 4);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int r = min((t + i * 128) >> 4, p.B - 1);  // rows >= B are zeroed below
            v[u][i] = *reinterpret_cast<const float4*>(p.x + (int64_t)r * p.ldx + kb * BK + c4 * 4);
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
          if (g0 + u < nkb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* xs_hi = smem + s * STAGE + 2 * A_BYTES;
            uint8_t* xs_lo = xs_hi + X_BYTES;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              const int r = (t + i * 128) >> 4;
              float4 w = v[u][i];
              if (r >= p.B) w = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.pro == 1) {
                w.x *= gw[u].x; w.y *= gw[u].y; w.z *= gw[u].z; w.w *= gw[u].w;
              } else if (p.pro == 2) {
                w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
              }
              bf16 h0, l0, h1, l1, h2, l2, h3, l3;
              split_bf16(w.x, h0, l0); split_bf16(w.y, h1, l1); split_bf16(w.z, h2, l2); split_bf16(w.w, h3, l3);
              // 128B swizzle: 16-byte chunk index XOR (row % 8); this float4 covers half a chunk (8 bytes)
              const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)c4 >> 1) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)c4 & 1u) << 3);
              uint2 ph2, pl2;
              ph2.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              ph2.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
              pl2.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              pl2.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
              *reinterpret_cast<uint2*>(xs_hi + off) = ph2;
              if (p.nplanes == 2) *reinterpret_cast<uint2*>(xs_lo + off) = pl2;
            }
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&xrdy_bar[s]);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        }
      }
    }
    // zero duty (other GEMM's accumulation buffer), off the critical path
    if (p.zero_ptr) {
      const int64_t n4 = p.zero_n >> 2;
      const int64_t nthreads = (int64_t)gridDim.x * gridDim.y * 128;
      float4* z4 = reinterpret_cast<float4*>(p.zero_ptr);
      for (int64_t i = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 128 + t; i < n4; i += nthreads)
        z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncwarp();
  asm volatile("bar.sync 3, 256;" ::: "memory");  // row scales s_rs[] published by the statistic warps
  if (warp < 6) {
    const int t = threadIdx.x - 64;
    // ---------------------------------------------------------------- epilogue
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int n = m0 + q * 32 + lane;  // output feature of this thread (TMEM lane)
#pragma unroll 1
    for (int c0 = 0; c0 < R; c0 += 32) {
      uint32_t rr[32];
      __syncwarp();
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, rr);
      tmem_ld_wait();
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int b = c0 + j;
          if (b < p.B) {
            const float v = __uint_as_float(rr[j]) * s_rs[b];
            float* o = p.out + (int64_t)b * p.ld_out + n;  // lanes -> consecutive features: coalesced
            if (p.store) *o = v; else atomicAdd(o, v);
          }
        }
      }
      if (p.amax_val) {  // warp-level (value, index) max per activation row, then across the 4 epilogue warps
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float bv = (n < p.N) ? __uint_as_float(rr[j]) * s_rs[min(c0 + j, R - 1)] : -INFINITY;
          int bi = n;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          if (lane == 0) {
            s_part[(c0 + j) * 4 + q] = bv;                              // s_part is free after the statistics
            reinterpret_cast<int*>(s_rs + R * 5)[(c0 + j) * 4 + q] = bi;  // [R][4] ints after s_part
          }
        }
      }
    }
    if (p.amax_val) {
      __syncwarp();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t < p.B) {
        const int* s_pi = reinterpret_cast<const int*>(s_rs + R * 5);
        float bv = s_part[t * 4];
        int bi = s_pi[t * 4];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float ov = s_part[t * 4 + w];
          const int oi = s_pi[t * 4 + w];
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        p.amax_val[(int64_t)t * gridDim.x + blockIdx.x] = bv;
        p.amax_idx[(int64_t)t * gridDim.x + blockIdx.x] = bi;
      }
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, R);
}

template <int NG>
static void launch_skinny_ng(cudaStream_t st, const SkinnyParams& p, dim3 grid) {
  constexpr int STAGE = 2 * A_BYTES + 2 * (32 * NG * BK * 2);
  const int smem = p.stages * STAGE + 1024 + 256 + 32 * NG * 9 * 4 + 64;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    MG_CHECK_CUDA(cudaFuncSetAttribute(skinny_tc_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_smem = 200 * 1024;
  }
  launch_pdl(skinny_tc_kernel<NG>, grid, dim3(320), (size_t)smem, st, p);
}

void launch_skinny_tc(cudaStream_t st, int pro, const float* x, int ldx, Planes W, int64_t ldw, float* out, int ld_out,
                      int B, int N, int K, const float* lnw, float eps, float scale, float* zero_ptr, int64_t zero_n,
                      bool store, float* amax_val, int* amax_idx, float* rs_scratch) {
  MG_REQUIRE(amax_val == nullptr || (store && amax_idx != nullptr), "fused argmax needs the direct-store (no split-K) mode");
  MG_REQUIRE(K % BK == 0 && ldx % 4 == 0, "skinny linear: K must be a multiple of 64");
  MG_REQUIRE(pro != 1 || K <= 1024, "skinny linear: fused RMSNorm needs K <= 1024");
  MG_REQUIRE(B >= 1 && B <= 128, "skinny linear: 1 <= B <= 128 per launch");
  MG_REQUIRE(zero_n % 4 == 0, "zero duty must be a multiple of 4 floats");
  SkinnyParams p;
  GemmOperand A;
  A.hi = W.hi; A.lo = W.lo; A.rows = N; A.ld = ldw;
  p.nplanes = W.lo ? 2 : 1;
  encode_operand(&p.tw_hi, W.hi, A, K, 1, 1, BM);
  if (W.lo) encode_operand(&p.tw_lo, W.lo, A, K, 1, 1, BM); else p.tw_lo = p.tw_hi;
  p.x = x; p.ldx = ldx; p.out = out; p.ld_out = ld_out; p.B = B; p.N = N; p.K = K;
  p.num_kb = K / BK;
  const int tiles = (N + BM - 1) / BM;
  int ksplit = 1;
  static const int target_ctas = getenv("MG_SKINNY_CTAS") ? atoi(getenv("MG_SKINNY_CTAS")) : 148;
  if (!store) ksplit = std::min(p.num_kb, std::max(1, target_ctas / tiles));
  p.kb_per_cta = (p.num_kb + ksplit - 1) / ksplit;
  ksplit = (p.num_kb + p.kb_per_cta - 1) / p.kb_per_cta;
  static const int max_stages = getenv("MG_SKINNY_STAGES") ? atoi(getenv("MG_SKINNY_STAGES")) : 4;
  p.stages = std::min(max_stages, p.kb_per_cta);  // <= 84 KB: CTAs of consecutive (PDL-overlapped) kernels co-reside
  {  // the ring must fit the 200 KB this kernel may use: a stage grows with the activation rows (64 KB at 65..128 rows)
    const int ng = B <= 32 ? 1 : (B <= 64 ? 2 : 4);
    const int stage = 2 * A_BYTES + 2 * (32 * ng * BK * 2);
    const int fixed = 1024 + 256 + 32 * ng * 9 * 4 + 64;
    p.stages = std::max(1, std::min(p.stages, (200 * 1024 - fixed) / stage));
    // 65..128 rows: two 64 KB stages instead of three, so that two CTAs share an SM (the LM head's 260 tiles then run in
    // one wave instead of two; measured at batch 128: 6.35 -> 6.28 ms per step)
    if (ng == 4 && !(getenv("MG_SKINNY_STAGES"))) p.stages = std::min(p.stages, 2);
  }
  p.pro = pro; p.lnw = lnw; p.eps = eps; p.scale = scale;
  p.zero_ptr = zero_ptr; p.zero_n = zero_n; p.store = store ? 1 : 0;
  p.amax_val = amax_val; p.amax_idx = amax_idx;
  p.rs_in = nullptr;
  if (pro == 1 && B > 32 && rs_scratch) {  // wide launch: row scales once, not once per CTA
    launch_pdl(rms_rowscale_kernel, dim3((B + 7) / 8), dim3(256), (size_t)0, st, x, ldx, B, K, eps, scale, rs_scratch);
    p.rs_in = rs_scratch;
  }
  dim3 grid(tiles, ksplit);
  if (B <= 32) launch_skinny_ng<1>(st, p, grid);
  else if (B <= 64) launch_skinny_ng<2>(st, p, grid);
  else launch_skinny_ng<4>(st, p, grid);
}

// ------------------------------------------------------------------------------------------------ fp32 -> planes
__global__ void split_kernel(const float* __restrict__ in, int64_t rows, int64_t cols, int64_t ld_in, bf16* hi,
                             bf16* lo, int64_t ld_out) {
  const int64_t total = rows * ld_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_out, c = i % ld_out;
    const float x = c < cols ? in[r * ld_in + c] : 0.f;
    bf16 h, l;
    split_bf16(x, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

void launch_split(cudaStream_t st, const float* in, int64_t rows, int64_t cols, int64_t ld_in, Planes out,
                  int64_t ld_out) {
  const int64_t total = rows * ld_out;
  if (total == 0) return;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  split_kernel<<<blocks, 256, 0, st>>>(in, rows, cols, ld_in, out.hi, out.lo, ld_out);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
