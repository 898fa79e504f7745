// Fused encoder self-attention (SURVEY K11 + K13): one tcgen05 flash kernel per encoder layer.
//
//   ctx[b, i, h*64:(h+1)*64] = softmax_j( q_i . k_j  +  (bias_v + (bias_h + bias_1d))[h, i, j]  +  (1 - mask_j) * finfo.min ) . v_j
//
// (UdopAttention.forward, transformers/models/udop/modeling_udop.py:531-622: unscaled dot product, position bias added
// to the scores, fp32 softmax; RelativePositionBiasAggregated :973-989 sums the 1-D / horizontal / vertical bucketed
// tables in that order; UdopStack :1173-1190 adds the extended mask).  The (B, 16, S, S) score / probability tensors of
// the reference -- 2 x 2.4 GB per layer at batch 32 in the unfused round-1 path -- are never written:
//
//   * a work item is (image, head, 256 query rows): two 128-row Q tiles stay in shared memory (TMA, split-bf16 planes);
//   * warp 0 streams K blocks (64 keys) and V^T blocks through two TMA rings; warp 1 issues the tcgen05.mma:
//     S_w = Q_w . K^T (three plane products, fp32 in TMEM, double-buffered per Q tile) and O_w += P_w . V;
//   * two softmax warpgroups (one per Q tile, thread = query row) read S from TMEM, add the bias -- the horizontal /
//     vertical bucket pair of every (i, j) is a 16-bit code computed ONCE per forward and shared by all layers and
//     heads (enc_bias_code_kernel), the 1-D bias is a per-head table indexed by j - i in shared memory -- keep a
//     running row maximum (the accumulator in TMEM is rescaled lazily, only when the maximum grows by more than 2^11),
//     and write the unnormalised probabilities as split-bf16 planes in the 128-byte-swizzled K-major layout the P . V
//     MMA reads; softmax of block k overlaps the Q . K^T of block k + 1 and the P . V of block k - 1;
//   * the epilogue divides by the row sum and writes the context as planes for the output projection GEMM.
#include <stdio.h>

#include <algorithm>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

constexpr int FA_THREADS = 384;      // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-7 / 8-11 softmax of Q tile 0 / 1
constexpr int FA_QT = 128;           // rows per Q tile (UMMA M)
constexpr int FA_KB = 64;            // keys per block (UMMA N of S, K of P.V)
constexpr int FA_Q_BYTES = FA_QT * 64 * 2;   // one plane of a Q tile
constexpr int FA_K_BYTES = FA_KB * 64 * 2;   // one plane of a K / V^T block
constexpr int FA_P_BYTES = FA_QT * FA_KB * 2;
constexpr int FA_KST = 2, FA_VST = 3;
// shared memory map (offsets from a 1024-byte aligned base)
constexpr int FA_OFF_Q = 0;                                   // [2 tiles][hi | lo]
constexpr int FA_OFF_K = FA_OFF_Q + 4 * FA_Q_BYTES;           // [KST][hi | lo]
constexpr int FA_OFF_V = FA_OFF_K + FA_KST * 2 * FA_K_BYTES;  // [VST][hi | lo]
constexpr int FA_OFF_P = FA_OFF_V + FA_VST * 2 * FA_K_BYTES;  // [2 tiles][hi | lo]
constexpr int FA_OFF_L1 = FA_OFF_P + 4 * FA_P_BYTES;          // float [2 * FA_MAXS]: 1-D bias of this head by j - i
constexpr int FA_MAXS = 1664;                                 // max padded sequence (S <= 1536 + slack)
constexpr int FA_OFF_TH = FA_OFF_L1 + 2 * FA_MAXS * 4;        // float [33] horizontal table of this head, [32] = -inf (invisible key)
constexpr int FA_OFF_TV = FA_OFF_TH + 256;                    // float [32] vertical
constexpr int FA_OFF_BAR = FA_OFF_TV + 128;                   // mbarriers
constexpr int FA_NBAR = 2 + 2 * FA_KST + 2 * FA_VST + 8 + 2 + 2 + 2;
constexpr int FA_OFF_SLOT = FA_OFF_BAR + FA_NBAR * 8;
constexpr int FA_SMEM = FA_OFF_SLOT + 16 + 1024;
static_assert(FA_SMEM <= 227 * 1024, "shared memory budget");

struct alignas(64) FlashParams {
  CUtensorMap tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo;
  const uint16_t* code;   // [B][nqt * 128][nkb * 64], tiled [qt][kb][i][j]: low byte = bucket_h * 4 (128 = key invisible),
                          // high byte = bucket_v * 4 -- both are byte offsets into the per-head tables
  const float *tab1d, *tabh, *tabv;  // [buckets][H]
  const int* lut1d;       // |j - i| -> bucket (without the sign offset)
  int lut1d_n, half_buckets;
  bf16 *out_hi, *out_lo;  // ctx planes [B * Sp][D]
  int B, H, D, Sp, nkb, nq256, nqt;
  float rescale_gap;      // the running maximum's reference point moves only when the maximum grew by more than this
  int opt;                // bit 0: bias codes loaded with an L2 evict-first policy (they are re-read once per head and must
                          // not push the K / V tiles out of L2); bit 1: L2 prefetch of the code line two key blocks ahead;
                          // bit 2: the producer prefetches the K / V blocks two ahead of the ring into L2
};

// bounded mbarrier wait: a protocol error becomes a trap (launch failure) instead of a hung GPU
__device__ __noinline__ void fa_wait(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 22)) {
      printf("enc_flash_attn_kernel: barrier %u parity %u never completed (block %d thread %d)\n", bar, parity,
             (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void fa_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fa_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fa_tma_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(0)
      : "memory");
}
__device__ __forceinline__ void fa_tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(0)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 2^x, one MUFU (rel. error 2^-22; exp2f() adds a denormal-range rescale = 3 more instructions per score)
__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte read-only load with an L2 eviction-priority hint (L1 allocation as usual: a thread reads its 128-byte code
// line with eight of these)
__device__ __forceinline__ uint4 fa_ldg_hint(const uint4* p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(FA_THREADS, 1) enc_flash_attn_kernel(const __grid_constant__ FlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sm_a = smem_u32(sm);
  float* const s_l1 = reinterpret_cast<float*>(sm + FA_OFF_L1);
  float* const s_th = reinterpret_cast<float*>(sm + FA_OFF_TH);
  float* const s_tv = reinterpret_cast<float*>(sm + FA_OFF_TV);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(sm + FA_OFF_SLOT);
  // barriers
  const uint32_t bar0 = sm_a + FA_OFF_BAR;
  const uint32_t b_qfull = bar0, b_qempty = bar0 + 8;
  const uint32_t b_kfull = bar0 + 16, b_kempty = b_kfull + 8 * FA_KST;
  const uint32_t b_vfull = b_kempty + 8 * FA_KST, b_vempty = b_vfull + 8 * FA_VST;
  const uint32_t b_sfull = b_vempty + 8 * FA_VST;  // [w][buf]
  const uint32_t b_sempty = b_sfull + 32;          // [w][buf]
  const uint32_t b_pready = b_sempty + 32;         // [w]
  const uint32_t b_pempty = b_pready + 16;         // [w]
  const uint32_t b_ofree = b_pempty + 16;          // [w]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // every CTA takes one contiguous run of items: consecutive items share (image, head), i.e. K / V stay in L2 and the
  // per-head bias tables are rebuilt only when the head changes
  const int n_total = p.B * p.H * p.nq256;
  const int per_cta = (n_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int item0 = (int)blockIdx.x * per_cta;
  const int n_items = min(n_total, item0 + per_cta);

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tq_hi); prefetch_tmap(&p.tq_lo); prefetch_tmap(&p.tk_hi);
    prefetch_tmap(&p.tk_lo); prefetch_tmap(&p.tv_hi); prefetch_tmap(&p.tv_lo);
    uint64_t* b = reinterpret_cast<uint64_t*>(sm + FA_OFF_BAR);
    mbar_init(b + 0, 1);  // q_full
    mbar_init(b + 1, 1);  // q_empty
    for (int i = 0; i < 2 * FA_KST + 2 * FA_VST; ++i) mbar_init(b + 2 + i, 1);
    uint64_t* c = b + 2 + 2 * FA_KST + 2 * FA_VST;
    for (int i = 0; i < 4; ++i) mbar_init(c + i, 1);        // s_full
    for (int i = 0; i < 4; ++i) mbar_init(c + 4 + i, 128);  // s_empty
    for (int i = 0; i < 2; ++i) mbar_init(c + 8 + i, 128);  // p_ready
    for (int i = 0; i < 2; ++i) mbar_init(c + 10 + i, 1);   // p_empty
    for (int i = 0; i < 2; ++i) mbar_init(c + 12 + i, 128); // o_free
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S of tile w, buffer u at w * 128 + u * 64; O of tile w at 256 + w * 64
  const int nkb = p.nkb;

  if (warp == 0) {
    // ================================================================================================ TMA producer
    if (lane == 0) {
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, it_n = 0;
      for (int item = item0; item < n_items; ++item, ++it_n) {
        const int q256 = item % p.nq256, bh = item / p.nq256;
        const int h = bh % p.H, b = bh / p.H;
        fa_wait(b_qempty, (it_n & 1) ^ 1);  // the previous item's Q . K^T products have all completed
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_qfull), "r"(4 * FA_Q_BYTES) : "memory");
        for (int w = 0; w < 2; ++w) {
          fa_tma_3d(sm_a + FA_OFF_Q + (2 * w) * FA_Q_BYTES, &p.tq_hi, b_qfull, h * 64, q256 * 256 + w * 128, b);
          fa_tma_3d(sm_a + FA_OFF_Q + (2 * w + 1) * FA_Q_BYTES, &p.tq_lo, b_qfull, h * 64, q256 * 256 + w * 128, b);
        }
        for (int kb = 0; kb < nkb; ++kb) {
          if ((p.opt & 4) && kb + 2 < nkb) {  // L2 prefetch of the K / V blocks two ahead of the ring
            fa_tma_prefetch_3d(&p.tk_hi, p.D + h * 64, (kb + 2) * FA_KB, b);
            fa_tma_prefetch_3d(&p.tk_lo, p.D + h * 64, (kb + 2) * FA_KB, b);
            fa_tma_prefetch_3d(&p.tv_hi, (kb + 2) * FA_KB, h * 64, b);
            fa_tma_prefetch_3d(&p.tv_lo, (kb + 2) * FA_KB, h * 64, b);
          }
          fa_wait(b_kempty + 8 * ks, kph ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_kfull + 8 * ks), "r"(2 * FA_K_BYTES) : "memory");
          fa_tma_3d(sm_a + FA_OFF_K + (2 * ks) * FA_K_BYTES, &p.tk_hi, b_kfull + 8 * ks, p.D + h * 64, kb * FA_KB, b);
          fa_tma_3d(sm_a + FA_OFF_K + (2 * ks + 1) * FA_K_BYTES, &p.tk_lo, b_kfull + 8 * ks, p.D + h * 64, kb * FA_KB, b);
          if (++ks == FA_KST) { ks = 0; kph ^= 1; }
          fa_wait(b_vempty + 8 * vs, vph ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_vfull + 8 * vs), "r"(2 * FA_K_BYTES) : "memory");
          fa_tma_3d(sm_a + FA_OFF_V + (2 * vs) * FA_K_BYTES, &p.tv_hi, b_vfull + 8 * vs, kb * FA_KB, h * 64, b);
          fa_tma_3d(sm_a + FA_OFF_V + (2 * vs + 1) * FA_K_BYTES, &p.tv_lo, b_vfull + 8 * vs, kb * FA_KB, h * 64, b);
          if (++vs == FA_VST) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================================================ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(FA_QT, FA_KB);
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0, it_n = 0;
    uint32_t n_s[2] = {0, 0};   // S tiles issued per Q tile (buffer = n & 1)
    uint32_t n_p[2] = {0, 0};   // P . V products issued per Q tile
    // S_w(buffer) = Q_w . K^T for the K block in ring stage ks
    auto issue_qk = [&](int w) {
      const uint32_t u = n_s[w] & 1u;
      fa_wait(b_sempty + 8 * (2 * w + u), ((n_s[w] >> 1) & 1u) ^ 1u);  // softmax has read this buffer's previous tile
      tc_fence_after();
      const uint32_t qa = sm_a + FA_OFF_Q + (2 * w) * FA_Q_BYTES, ka = sm_a + FA_OFF_K + (2 * ks) * FA_K_BYTES;
      const uint64_t dq_hi = make_sw128_kmajor_desc(qa), dq_lo = make_sw128_kmajor_desc(qa + FA_Q_BYTES);
      const uint64_t dk_hi = make_sw128_kmajor_desc(ka), dk_lo = make_sw128_kmajor_desc(ka + FA_K_BYTES);
      const uint32_t acc = tmem + (uint32_t)(w * 128) + u * 64;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(acc, dq_lo + 2 * k, dk_hi + 2 * k, idesc, k == 0 ? 0u : 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(acc, dq_hi + 2 * k, dk_lo + 2 * k, idesc, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(acc, dq_hi + 2 * k, dk_hi + 2 * k, idesc, 1);
        fa_commit(b_sfull + 8 * (2 * w + u));
      }
      __syncwarp();
      ++n_s[w];
    };
    for (int item = item0; item < n_items; ++item, ++it_n) {
      fa_wait(b_qfull, it_n & 1);
      // prologue: S(0) of both tiles
      fa_wait(b_kfull + 8 * ks, kph);
      tc_fence_after();
      issue_qk(0);
      issue_qk(1);
      if (elect_one()) fa_commit(b_kempty + 8 * ks);
      __syncwarp();
      if (++ks == FA_KST) { ks = 0; kph ^= 1; }
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb + 1 < nkb) {  // S(kb + 1) while the softmax warps work on S(kb)
          fa_wait(b_kfull + 8 * ks, kph);
          tc_fence_after();
          issue_qk(0);
          issue_qk(1);
          if (elect_one()) fa_commit(b_kempty + 8 * ks);
          __syncwarp();
          if (++ks == FA_KST) { ks = 0; kph ^= 1; }
        } else {
          if (elect_one()) fa_commit(b_qempty);  // every Q . K^T of this item has been issued
          __syncwarp();
        }
        fa_wait(b_vfull + 8 * vs, vph);
        const uint32_t va = sm_a + FA_OFF_V + (2 * vs) * FA_K_BYTES;
        const uint64_t dv_hi = make_sw128_kmajor_desc(va), dv_lo = make_sw128_kmajor_desc(va + FA_K_BYTES);
        for (int w = 0; w < 2; ++w) {
          if (kb == 0 && it_n > 0) fa_wait(b_ofree + 8 * w, (it_n - 1) & 1);  // the previous item's O has been read out
          fa_wait(b_pready + 8 * w, n_p[w] & 1);
          tc_fence_after();
          const uint32_t pa = sm_a + FA_OFF_P + (2 * w) * FA_P_BYTES;
          const uint64_t dp_hi = make_sw128_kmajor_desc(pa), dp_lo = make_sw128_kmajor_desc(pa + FA_P_BYTES);
          const uint32_t acc = tmem + 256u + (uint32_t)(w * 64);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(acc, dp_lo + 2 * k, dv_hi + 2 * k, idesc, (kb == 0 && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(acc, dp_hi + 2 * k, dv_lo + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(acc, dp_hi + 2 * k, dv_hi + 2 * k, idesc, 1);
            fa_commit(b_pempty + 8 * w);
          }
          __syncwarp();
          ++n_p[w];
        }
        if (elect_one()) fa_commit(b_vempty + 8 * vs);
        __syncwarp();
        if (++vs == FA_VST) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================================================ softmax
    const int w = (warp - 4) >> 2;              // Q tile of this warpgroup
    const int t = threadIdx.x - 128 - w * 128;  // row inside the tile = TMEM lane
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t n_blk = 0, n_pe = 0, it_n = 0;
    constexpr float LOG2E = 1.4426950408889634f;
    // (measured: streaming the codes with L1::no_allocate + an L2 evict-first policy is SLOWER -- encoder 109.9 ->
    // 123.6 ms per batch: a thread reads its 128-byte code line with eight 16-byte loads, and without the L1 allocation
    // each of them goes to L2)
    uint64_t pol_first;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    const bool code_first = (p.opt & 1) != 0, code_pf = (p.opt & 2) != 0;
    uint8_t* const p_hi = sm + FA_OFF_P + (2 * w) * FA_P_BYTES + t * 128;
    uint8_t* const p_lo = p_hi + FA_P_BYTES;
    int h_tab = -1;  // head whose tables are in shared memory
    for (int item = item0; item < n_items; ++item, ++it_n) {
      const int q256 = item % p.nq256, bh = item / p.nq256;
      const int h = bh % p.H, b = bh / p.H;
      const int i = q256 * 256 + w * 128 + t;  // query row
      // per-head tables: both warpgroups rebuild them when the head changes (all 256 softmax threads, barrier 1)
      if (h != h_tab) {
      h_tab = h;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone is done with the previous head's tables
      {
        const int tt = threadIdx.x - 128;
        if (tt < 32) s_th[tt] = p.tabh[tt * p.H + h];
        else if (tt < 64) s_tv[tt - 32] = p.tabv[(tt - 32) * p.H + h];
        else if (tt == 64) s_th[32] = -INFINITY;  // invisible keys index this entry: their bias, hence their score, is -inf
        for (int dd = tt; dd < 2 * p.Sp + 64; dd += 256) {  // index dd = (j - i) + Sp (finite for the padded keys too)
          int rel = dd - p.Sp;
          const int o1 = rel > 0 ? p.half_buckets : 0;
          rel = rel < 0 ? -rel : rel;
          rel = min(rel, p.lut1d_n - 1);
          s_l1[dd] = p.tab1d[(o1 + p.lut1d[rel]) * p.H + h];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      // a warp whose 32 rows all lie past the sequence (the last item of an (image, head) is padded to 256 rows) only
      // keeps the barrier protocol going: it leaves the issue slots of its scheduler to the warps with real rows
      if (q256 * 256 + w * 128 + quad * 32 >= p.Sp) {
        for (int kb = 0; kb < nkb; ++kb, ++n_blk) {
          const uint32_t u = n_blk & 1u;
          fa_wait(b_sfull + 8 * (2 * w + u), (n_blk >> 1) & 1u);
          tc_fence_before();
          fa_arrive(b_sempty + 8 * (2 * w + u));
          if (kb > 0) {
            fa_wait(b_pempty + 8 * w, n_pe & 1);
            ++n_pe;
          }
          fa_arrive(b_pready + 8 * w);
        }
        fa_wait(b_pempty + 8 * w, n_pe & 1);
        ++n_pe;
        fa_arrive(b_ofree + 8 * w);
        continue;
      }
      const uint16_t* code_row = p.code + ((size_t)b * p.nqt + (size_t)(i >> 7)) * (size_t)nkb * (128 * 64) + (size_t)(i & 127) * 64;
      uint4 cd[8];  // this row's 64 codes of the current block
      {
        const uint4* src = reinterpret_cast<const uint4*>(code_row);
        if (code_first) {
#pragma unroll
          for (int v = 0; v < 8; ++v) cd[v] = fa_ldg_hint(src + v, pol_first);
        } else {
#pragma unroll
          for (int v = 0; v < 8; ++v) cd[v] = __ldg(src + v);
        }
        if (code_pf && nkb > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(code_row + (size_t)(128 * 64)));
      }
      float m_ref = -INFINITY, sum = 0.f;
      for (int kb = 0; kb < nkb; ++kb, ++n_blk) {
        const uint32_t u = n_blk & 1u;
        fa_wait(b_sfull + 8 * (2 * w + u), (n_blk >> 1) & 1u);
        tc_fence_after();
        float s[64];
        {
          uint32_t r0[32], r1[32];
          const uint32_t ta = tmem + lane_off + (uint32_t)(w * 128) + u * 64;
          tmem_ld_32x32(ta, r0);
          tmem_ld_32x32(ta + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) { s[c] = __uint_as_float(r0[c]); s[32 + c] = __uint_as_float(r1[c]); }
        }
        tc_fence_before();
        fa_arrive(b_sempty + 8 * (2 * w + u));
        // ---- bias + mask: s = qk + (tv + (th + t1d)); invisible keys drop out (exp(finfo.min - max) == 0 in fp32)
        const float* l1 = s_l1 + (kb * FA_KB - min(i, p.Sp - 1) + p.Sp);  // rows past the sequence: any finite bias
        float bm = -INFINITY;
        const uint32_t* cw = reinterpret_cast<const uint32_t*>(cd);
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2) {
          const uint32_t two = cw[c2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * c2 + e;
            const uint32_t cc = e ? (two >> 16) : (two & 0xffffu);
            (void)cc;
            // the two code bytes ARE the table offsets (one PRMT each); an invisible key reads th[32] = -inf
            const uint32_t oh = __byte_perm(two, 0u, e ? 0x4442 : 0x4440), ov = __byte_perm(two, 0u, e ? 0x4443 : 0x4441);
            const float bias = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(s_tv) + ov) +
                               (*reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(s_th) + oh) + l1[c]);
            const float v = s[c] + bias;
            s[c] = v;
            bm = fmaxf(bm, v);
          }
        }
        // next block's codes: the loads fly during the exponentials below
        if (kb + 1 < nkb) {
          const uint4* src = reinterpret_cast<const uint4*>(code_row + (size_t)(kb + 1) * (128 * 64));
          if (code_first) {
#pragma unroll
            for (int v = 0; v < 8; ++v) cd[v] = fa_ldg_hint(src + v, pol_first);
          } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) cd[v] = __ldg(src + v);
          }
          if (code_pf && kb + 3 < nkb) asm volatile("prefetch.global.L2 [%0];" ::"l"(code_row + (size_t)(kb + 3) * (128 * 64)));
        }
        // ---- running maximum, lazily updated: the reference point only moves when the maximum grew by > 2^11 (the
        // probabilities of a block then reach at most 2048: no overflow, and the accumulator is rarely rescaled)
        float scale_o = 1.f;
        bool moved = false;
        if (bm > m_ref + p.rescale_gap || m_ref == -INFINITY) {
          if (bm != -INFINITY) {
            scale_o = (m_ref == -INFINITY) ? 0.f : fa_ex2((m_ref - bm) * LOG2E);
            sum *= scale_o;
            m_ref = bm;
            moved = true;
          }
        }
        const float mneg = (m_ref == -INFINITY) ? 0.f : -m_ref * LOG2E;
        float bsum = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          s[c] = fa_ex2(fmaf(s[c], LOG2E, mneg));  // -inf -> 0
          bsum += s[c];
        }
        sum += bsum;
        // ---- P smem is free once P.V of the previous block has completed (then O is quiescent too)
        if (kb > 0) {
          fa_wait(b_pempty + 8 * w, n_pe & 1);
          ++n_pe;
          tc_fence_after();
          if (__any_sync(0xffffffffu, moved)) {  // rescale this warp's 32 accumulator rows
            const uint32_t oa = tmem + lane_off + 256u + (uint32_t)(w * 64);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t r[32];
              tmem_ld_32x32(oa + half * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * scale_o);
              tmem_st_32x32(oa + half * 32, r);
            }
            tmem_st_wait();
          }
        }
        // ---- split planes, 128-byte swizzle (16-byte chunk index XOR (row & 7))
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = s[ch * 8 + 2 * e], c = s[ch * 8 + 2 * e + 1];
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
            const float2 f2 = __bfloat1622float2(h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - f2.x, c - f2.y);
            hw[e] = *reinterpret_cast<const uint32_t*>(&h2);
            lw[e] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          const uint32_t off = (uint32_t)((ch ^ (t & 7)) << 4);
          *reinterpret_cast<uint4*>(p_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(p_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async();
        tc_fence_before();
        fa_arrive(b_pready + 8 * w);
      }
      // ---- epilogue: O / sum -> ctx planes
      fa_wait(b_pempty + 8 * w, n_pe & 1);
      ++n_pe;
      tc_fence_after();
      const float inv = 1.f / sum;
      const bool row_ok = i < p.Sp;
      bf16* oh = p.out_hi + ((size_t)b * p.Sp + (size_t)(row_ok ? i : 0)) * p.D + h * 64;
      bf16* ol = p.out_lo ? p.out_lo + ((size_t)b * p.Sp + (size_t)(row_ok ? i : 0)) * p.D + h * 64 : nullptr;
      const uint32_t oa = tmem + lane_off + 256u + (uint32_t)(w * 64);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32(oa + half * 32, r);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          fa_arrive(b_ofree + 8 * w);
        }
        if (row_ok) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(r[ch * 8 + 2 * e]) * inv, c = __uint_as_float(r[ch * 8 + 2 * e + 1]) * inv;
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
              const float2 f2 = __bfloat1622float2(h2);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - f2.x, c - f2.y);
              hw[e] = *reinterpret_cast<const uint32_t*>(&h2);
              lw[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            *reinterpret_cast<uint4*>(oh + half * 32 + ch * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            if (ol) *reinterpret_cast<uint4*>(ol + half * 32 + ch * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ bias codes
// Horizontal / vertical relative-position buckets (RelativePositionBiasHorizontal / Vertical :935-970,
// get_relative_position :887-895, bucket :466-512) + key visibility, ONCE per forward, shared by all layers / heads:
//   pos = (b0 + b2) / 2 (float64);  rel = ((pos_j - pos_i) * 100).long();  bucket = (rel > 0) * 16 + lut[min(|rel|, cap)]
// code: low byte = bucket_h * 4, or 128 if key j is invisible; high byte = bucket_v * 4 (byte offsets into the per-head
// tables), tiled [b][i / 128][j / 64][i % 128][j % 64] so that the 64 codes a softmax thread needs for one key block are
// one 128-byte line.
__global__ void enc_bias_code_kernel(const double* __restrict__ bbox_ext, const int* __restrict__ mask, int Sp, int nqt,
                                     int nkb, const int* __restrict__ lut_hv, int lut_n, int half_buckets,
                                     double scaling, uint16_t* __restrict__ code) {
  const int i = blockIdx.x, b = blockIdx.y;  // i < nqt * 128
  const bool i_ok = i < Sp;
  const double* bi = bbox_ext + ((int64_t)b * Sp + (i_ok ? i : 0)) * 4;
  const double xi = (bi[0] + bi[2]) / 2.0, yi = (bi[1] + bi[3]) / 2.0;
  uint16_t* out = code + ((size_t)b * nqt + (size_t)(i >> 7)) * (size_t)nkb * (128 * 64) + (size_t)(i & 127) * 64;
  for (int j = threadIdx.x; j < nkb * 64; j += blockDim.x) {
    uint16_t c = 0;
    if (j < Sp && i_ok) {
      const double* bj = bbox_ext + ((int64_t)b * Sp + j) * 4;
      const double xj = (bj[0] + bj[2]) / 2.0, yj = (bj[1] + bj[3]) / 2.0;
      long long rx = (long long)((xj - xi) * scaling);
      long long ry = (long long)((yj - yi) * scaling);
      const int ox = rx > 0 ? half_buckets : 0, oy = ry > 0 ? half_buckets : 0;
      rx = rx < 0 ? -rx : rx;
      ry = ry < 0 ? -ry : ry;
      if (rx > lut_n - 1) rx = lut_n - 1;
      if (ry > lut_n - 1) ry = lut_n - 1;
      const int vis = mask[(int64_t)b * Sp + j] != 0;
      c = (uint16_t)((vis ? ((ox + lut_hv[rx]) << 2) : 128) | ((oy + lut_hv[ry]) << 10));
    } else if (j < Sp) {
      c = (uint16_t)(mask[(int64_t)b * Sp + j] ? 0 : 128);  // rows past the sequence: finite scores, never stored
    } else {
      c = 128;                                               // keys past the sequence: invisible
    }
    out[(size_t)(j >> 6) * (128 * 64) + (j & 63)] = c;
  }
}

size_t enc_bias_code_bytes(int B, int Sp) {
  const int nqt = (Sp + 255) / 256 * 2, nkb = (Sp + FA_KB - 1) / FA_KB;
  return (size_t)B * nqt * 128 * (size_t)nkb * 64 * sizeof(uint16_t);
}

void launch_enc_bias_code(cudaStream_t st, const double* bbox_ext, const int* mask, int B, int Sp, const int* lut_hv,
                          int lut_n, int half_buckets, uint16_t* code) {
  MG_REQUIRE(half_buckets * 2 <= 32, "encoder flash attention: at most 32 relative-position buckets");
  const int nqt = (Sp + 255) / 256 * 2, nkb = (Sp + FA_KB - 1) / FA_KB;
  dim3 grid(nqt * 128, B);
  enc_bias_code_kernel<<<grid, 256, 0, st>>>(bbox_ext, mask, Sp, nqt, nkb, lut_hv, lut_n, half_buckets, 100.0, code);
  MG_CHECK_CUDA(cudaGetLastError());
}

void launch_enc_flash_attn(cudaStream_t st, Planes qk, Planes vt, const uint16_t* code, const float* tab1d,
                           const float* tabh, const float* tabv, const int* lut1d, int lut1d_n, int half_buckets,
                           int nbuckets, int B, int H, int D, int Sp, Planes ctx) {
  MG_REQUIRE(D == H * 64, "encoder flash attention: head_dim must be 64");
  MG_REQUIRE(Sp % 8 == 0 && Sp <= FA_MAXS - 32, "encoder flash attention: sequence must be a multiple of 8, <= 1632");
  MG_REQUIRE(nbuckets <= 32 && qk.lo && vt.lo && ctx.hi, "encoder flash attention: split planes and <= 32 buckets");
  FlashParams p;
  GemmOperand q;  // [B][Sp rows][2D cols]: Q at columns [0, D), K at [D, 2D)
  q.rows = Sp; q.ld = 2 * D; q.use_b1 = true; q.bs1 = (int64_t)Sp * 2 * D;
  make_tmap_bf16(&p.tq_hi, qk.hi, q, 2 * D, B, FA_QT);
  make_tmap_bf16(&p.tq_lo, qk.lo, q, 2 * D, B, FA_QT);
  make_tmap_bf16(&p.tk_hi, qk.hi, q, 2 * D, B, FA_KB);
  make_tmap_bf16(&p.tk_lo, qk.lo, q, 2 * D, B, FA_KB);
  GemmOperand v;  // [B][D rows][Sp cols]
  v.rows = D; v.ld = Sp; v.use_b1 = true; v.bs1 = (int64_t)D * Sp;
  make_tmap_bf16(&p.tv_hi, vt.hi, v, Sp, B, 64);
  make_tmap_bf16(&p.tv_lo, vt.lo, v, Sp, B, 64);
  p.code = code; p.tab1d = tab1d; p.tabh = tabh; p.tabv = tabv; p.lut1d = lut1d; p.lut1d_n = lut1d_n;
  p.half_buckets = half_buckets;
  p.out_hi = ctx.hi; p.out_lo = ctx.lo;
  p.B = B; p.H = H; p.D = D; p.Sp = Sp;
  p.nkb = (Sp + FA_KB - 1) / FA_KB;
  p.nq256 = (Sp + 255) / 256;
  p.nqt = p.nq256 * 2;
  // MG_FLASH_GAP=<natural-log units>: test hook, 0 rescales the TMEM accumulator on every growth of the row maximum
  static const float gap = getenv("MG_FLASH_GAP") ? (float)atof(getenv("MG_FLASH_GAP")) : 11.f * 0.6931471805599453f;
  p.rescale_gap = gap;
  // A/B switch, see FlashParams::opt.  Encoder ms per batch of 32 on one box (gpurun_out/r2k_enc_opt*.log): 0: 113.4,
  // 1: 112.3, 2: 112.0, 3: 110.9, 4: 112.2, 5: 112.2, 7: 110.7 -> evict-first codes + code prefetch by default
  static const int opt = getenv("MG_FLASH_OPT") ? atoi(getenv("MG_FLASH_OPT")) : 3;
  p.opt = opt;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    MG_CHECK_CUDA(cudaGetDevice(&dev));
    MG_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    MG_CHECK_CUDA(cudaFuncSetAttribute(enc_flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
  }
  const int items = B * H * p.nq256;
  const int per_cta = (items + n_sm - 1) / n_sm;
  enc_flash_attn_kernel<<<(items + per_cta - 1) / per_cta, FA_THREADS, FA_SMEM, st>>>(p);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
