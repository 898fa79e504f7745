// The whole decoder step (all layers + LM head) of the greedy decode loop as ONE persistent cooperative kernel.
//
// Why: a decode step is a strictly sequential chain of ~8 small dependent operations per layer.  As separate
// kernels each one pays launch + drain + cold-start latency (~8 us against ~1-2 us of HBM time for its
// weights) and HBM sits idle in between.  Here one CTA per SM stays resident for the whole step:
//
//   * warp 0 (one lane) is a PRODUCER that walks the step's static load program -- weight tiles of every linear,
//     the self-attention K/V blocks, the cross-attention K/V blocks, in the exact order the CTA consumes them --
//     and streams it through a 5 x 40 KB shared-memory ring with cp.async.bulk + mbarrier expect_tx.  It never
//     waits for a grid barrier: everything it reads is constant during the step, so the weights of the NEXT
//     operation are already in shared memory while the other SMs are still finishing the current one and HBM
//     never drains between dependent operations.
//   * warp 1 (one lane) issues tcgen05.mma for the linears: weights on the UMMA M axis (128 features per tile,
//     pre-swizzled split-bf16 planes straight from the ring), the <=32 activation rows on the N axis, fp32
//     accumulator in TMEM, three products per k-step (lo*hi + hi*lo + hi*hi) like gemm_tc.cu.
//   * warps 2-9 are CONSUMERS: they stage the activation tiles (fused RMSNorm / ReLU, fp32 -> split planes in the
//     128B-swizzled K-major layout), run the TMEM epilogues (split-K partial sums accumulate with red.global.add
//     into the residual stream / next buffer), and run the two attention phases out of the ring.
//   * operations are separated by a software grid barrier (one arrive + acquire-spin per CTA) in which only the
//     consumer warps take part.
//
// Arithmetic follows transformers/models/udop/modeling_udop.py exactly like the unfused kernels in decode.cu /
// gemm_tc.cu (UdopLayerNorm :333-355, UdopAttention :431-622 incl. compute_bias :514-529, UdopLayerFF :412-427,
// decoder UdopStack :1146-1256, lm head :1585-1590); the selection stays in greedy_select_kernel.
#include <algorithm>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

constexpr int MK_THREADS = 320;
constexpr int MK_NST = 5;
constexpr int MK_STAGE = 40960;
constexpr int MK_WTILE = 32768;  // one (tile, k-block): [hi 128x64 bf16 swizzled][lo ...]
constexpr int MK_XOFF = 32768;   // activation tile inside a stage: hi [32][64] bf16 (4 KB) then lo (4 KB)
constexpr int MK_XPLANE = 4096;
constexpr int MK_R = 32;         // activation rows (UMMA N)
constexpr int MK_MAXSC = 2048;   // max keys of one attention row (cross: Mp, self: max_length)
constexpr int MK_SELF_KB = 4;    // 32-key blocks per self-K chunk (32 KB)
constexpr int MK_SELF_VR = 128;  // keys per self-V chunk (32 KB)
constexpr int MK_CROSS_VR = MK_STAGE / 256;  // keys per cross-V chunk (fp32: 64 floats per key)

// fine-grained in-kernel stamps (profiling builds only: -DMK_FINE); slot layout in tools/mega_phase_profile.py
#ifdef MK_FINE
#define MK_STAMP(ptr, i) do { if (ptr) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); (ptr)[i] = _t; } } while (0)
#else
#define MK_STAMP(ptr, i) do { } while (0)
#endif

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// bounded spin: a hang becomes a trap (-> launch failure) instead of a dead GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  } while (!done);
}

struct RingPos {
  int s = 0;
  uint32_t ph = 0;
  uint32_t xmask = 0;  // per-stage parity of the "activation tile staged" barrier (flips on linear uses only)
  __device__ __forceinline__ void adv() {
    if (++s == MK_NST) { s = 0; ph ^= 1; }
  }
  __device__ __forceinline__ void adv_n(int n) {
    const int t = s + n;
    ph ^= (uint32_t)((t / MK_NST) & 1);
    s = t % MK_NST;
  }
};

__device__ __forceinline__ int mega_lin_of_phase(int ph) { return ph == 0 ? 0 : (ph < 4 ? ph - 1 : ph - 2); }  // 0,2,3,5,6,7 -> 0..5

__device__ __forceinline__ int items_of_cta(int total, int g, int G) { return g < total ? (total - g + G - 1) / G : 0; }

struct MegaShared {
  uint8_t* ring;
  float *sc, *sq, *snew, *sred, *s_rs, *s_part, *s_b;
  int* s_pi;
  uint64_t *full, *empty, *xrdy, *tmem_full, *tmem_empty;
  uint32_t* tmem_slot;
};

constexpr int MK_SMEM_MISC = MK_MAXSC * 4 + 64 * 4 + 64 * 4 + 16 * 64 * 4 + MK_R * 4 + MK_R * 4 * 4 + MK_R * 4 * 4 + 64 +
                             (3 * MK_NST + 2) * 8 + 16;
constexpr int MK_SMEM = MK_NST * MK_STAGE + MK_SMEM_MISC + 1024;

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(MK_THREADS, 1) decode_step_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  MegaShared S;
  // pointer arithmetic (not an integer round trip) keeps the shared address space: LDS/STS instead of generic LD/ST
  S.ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  S.sc = reinterpret_cast<float*>(S.ring + MK_NST * MK_STAGE);
  S.sq = S.sc + MK_MAXSC;
  S.snew = S.sq + 64;
  S.sred = S.snew + 64;
  S.s_rs = S.sred + 16 * 64;
  S.s_part = S.s_rs + MK_R;
  S.s_pi = reinterpret_cast<int*>(S.s_part + MK_R * 4);
  S.s_b = reinterpret_cast<float*>(S.s_pi + MK_R * 4);
  S.full = reinterpret_cast<uint64_t*>(S.s_b + 16);
  S.empty = S.full + MK_NST;
  S.xrdy = S.empty + MK_NST;
  S.tmem_full = S.xrdy + MK_NST;
  S.tmem_empty = S.tmem_full + 1;
  S.tmem_slot = reinterpret_cast<uint32_t*>(S.tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, G = gridDim.x;
  const int step = *p.step_ptr;  // tokens already in the self-attention caches
  const int H = p.H, D = p.D, B = p.B, Mp = p.Mp;
  const int n_attn = B * H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MK_NST; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
      mbar_init(&S.xrdy[s], 128);
    }
    mbar_init(S.tmem_full, 1);
    mbar_init(S.tmem_empty, 128);
    *reinterpret_cast<volatile int*>(S.s_b + 8) = 0;  // phase the consumers have entered (prefetch gate)
    fence_mbar_init();
    if (g == 0) p.bar_ctr[(step + 1) & 1] = 0u;  // the other parity's counter is idle during this launch
  }
  if (warp == 1) {
    tmem_alloc(S.tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *S.tmem_slot;

  // per-item chunk counts of the two attention phases (identical for every item of a phase)
  const int self_nblk = (step + 31) >> 5;
  const int self_nkc = (self_nblk + MK_SELF_KB - 1) / MK_SELF_KB;
  const int self_nvc = (step + MK_SELF_VR - 1) / MK_SELF_VR;
  const int cross_rk = min(64, MK_STAGE / (Mp * 4));
  const int cross_nkc = (64 + cross_rk - 1) / cross_rk;
  const int cross_nvc = (Mp + MK_CROSS_VR - 1) / MK_CROSS_VR;
  const int my_attn = items_of_cta(n_attn, g, G);
  const int Tb = p.Tp >> 5;  // 32-key blocks per (image, head) of the self K cache

  if (warp == 0) {
    // =============================================================================================== producer
    if (lane == 0) {
      RingPos r;
      int n_put = 0;
      const int max_inflight = p.max_inflight;
      auto put = [&](const void* src, uint32_t bytes) {
        // cap the bytes this SM has outstanding in the memory system: a deeper queue adds no bandwidth, only
        // latency for the latency-critical traffic (activation loads, split-K reductions, barrier flags)
        if (n_put >= max_inflight) {
          const int m = n_put - max_inflight;
          mbar_wait_wd(&S.full[m % MK_NST], (uint32_t)((m / MK_NST) & 1));
        }
        ++n_put;
        mbar_wait_wd(&S.empty[r.s], r.ph ^ 1);
        mbar_expect_tx(&S.full[r.s], bytes);
        bulk_load_1d(S.ring + (size_t)r.s * MK_STAGE, src, bytes, &S.full[r.s]);
        r.adv();
      };
      auto lin_loads = [&](const MegaLin W) {
        const int items = W.tiles * W.ksplit;
        for (int it = g; it < items; it += G) {
          const int tile = it / W.ksplit, ks = it - tile * W.ksplit;
          const int kb0 = ks * W.kb_per_item, kb1 = min(W.num_kb, kb0 + W.kb_per_item);
          for (int kb = kb0; kb < kb1; ++kb) put(W.w + ((size_t)tile * W.num_kb + kb) * MK_WTILE, MK_WTILE);
        }
      };
      // one call site per phase kind (the program is a loop, not straight-line code: the step must stay
      // resident in the instruction cache)
      for (int l = 0; l <= p.NL; ++l) {
        const MegaLayer& L = p.layers[min(l, p.NL - 1)];
        const int nph = l < p.NL ? 8 : 1;
        for (int ph = 0; ph < nph; ++ph) {
          if (p.gate && l < p.NL && (ph == 1 || ph == 4)) {
            // Prefetch gate: the K/V stream of an attention phase starts only when this CTA's consumers have
            // entered that phase.  A saturated memory system multiplies the latency of everything else (L2 loads
            // 0.2 -> 1.7 us, a grid barrier 1.2 -> 12 us, measured: tools/ubench/phase_bench.cu), so the bulk
            // stream must not run underneath the latency-bound linears that precede it.
            const int want = l * 8 + ph;
            volatile int* flag = reinterpret_cast<volatile int*>(S.s_b + 8);
            const long long t0 = clock64();
            while (*flag < want) {
              if (clock64() - t0 > 4000000000LL) __trap();
            }
          }
          if (l < p.NL && ph == 1) {
            for (int it = g; it < n_attn; it += G) {  // it = b * H + h
              const float* kb_ = L.skb + (size_t)it * Tb * 2048;
              for (int c = 0; c < self_nkc; ++c) {
                const int nb = min(MK_SELF_KB, self_nblk - c * MK_SELF_KB);
                put(kb_ + (size_t)c * MK_SELF_KB * 2048, (uint32_t)nb * 8192u);
              }
              const float* vb_ = L.svb + (size_t)it * p.Tp * 64;
              for (int c = 0; c < self_nvc; ++c) {
                const int rows = min(MK_SELF_VR, step - c * MK_SELF_VR);
                put(vb_ + (size_t)c * MK_SELF_VR * 64, (uint32_t)rows * 256u);
              }
            }
          } else if (l < p.NL && ph == 4) {
            for (int it = g; it < n_attn; it += G) {
              const float* ktb = L.ckt + (size_t)it * 64 * Mp;
              for (int c = 0; c < cross_nkc; ++c) {
                const int r0 = c * cross_rk, rows = min(cross_rk, 64 - r0);
                put(ktb + (size_t)r0 * Mp, (uint32_t)rows * Mp * 4u);
              }
              const float* vb_ = L.cv + (size_t)it * Mp * 64;
              for (int c = 0; c < cross_nvc; ++c) {
                const int m0 = c * MK_CROSS_VR, rows = min(MK_CROSS_VR, Mp - m0);
                put(vb_ + (size_t)m0 * 64, (uint32_t)rows * 256u);
              }
            }
          } else {
            lin_loads(l < p.NL ? L.lin[mega_lin_of_phase(ph)] : p.lm_head);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================================================== MMA issuer
    if (lane == 0) {
      RingPos r;
      uint32_t n_item = 0;
      constexpr uint32_t idesc = make_idesc_bf16(128, MK_R);
      int mma_phase = 0;
      auto lin_mma = [&](const MegaLin W) {
        const int items = W.tiles * W.ksplit;
#ifdef MK_FINE
        unsigned long long* fm = (p.prof && mma_phase >= 6 && mma_phase < 12) ? p.prof + ((size_t)g * 512 + 400 + (mma_phase - 6) * 4) * 2 : nullptr;
#endif
        ++mma_phase;
        for (int it = g; it < items; it += G) {
          const int tile = it / W.ksplit, ks = it - tile * W.ksplit;
          const int kb0 = ks * W.kb_per_item, kb1 = min(W.num_kb, kb0 + W.kb_per_item);
          if (n_item > 0) mbar_wait_wd(S.tmem_empty, (n_item - 1) & 1);  // epilogue of the previous item drained TMEM
          tc_fence_after();
          uint32_t acc = 0;
          for (int kb = kb0; kb < kb1; ++kb) {
            // activation tile first: once it is staged, the stage's previous occupant has been consumed, so the
            // parity wait on the weight barrier below cannot alias an older phase (this warp skips the attention
            // phases and may be far ahead of the ring)
            mbar_wait_wd(&S.xrdy[r.s], (r.xmask >> r.s) & 1u);
            r.xmask ^= 1u << r.s;
            if (kb == kb0) MK_STAMP(fm, 1);
            mbar_wait_wd(&S.full[r.s], r.ph);
            if (kb == kb0) MK_STAMP(fm, 2);
            tc_fence_after();
            const uint32_t sa = smem_u32(S.ring + (size_t)r.s * MK_STAGE);
            const uint64_t da_hi = make_sw128_kmajor_desc(sa), da_lo = make_sw128_kmajor_desc(sa + 16384);
            const uint64_t db_hi = make_sw128_kmajor_desc(sa + MK_XOFF);
            const uint64_t db_lo = make_sw128_kmajor_desc(sa + MK_XOFF + MK_XPLANE);
#pragma unroll
            for (int k = 0; k < 4; ++k) { umma_bf16(tmem_base, da_lo + 2 * k, db_hi + 2 * k, idesc, acc); acc = 1; }
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_lo + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_hi + 2 * k, idesc, 1);
            umma_commit(&S.empty[r.s]);
            r.adv();
          }
          umma_commit(S.tmem_full);
          MK_STAMP(fm, 3);
          ++n_item;
        }
      };
      for (int l = 0; l <= p.NL; ++l) {
        const MegaLayer& L = p.layers[min(l, p.NL - 1)];
        const int nph = l < p.NL ? 8 : 1;
        for (int ph = 0; ph < nph; ++ph) {
          if (l < p.NL && ph == 1) r.adv_n(my_attn * (self_nkc + self_nvc));
          else if (l < p.NL && ph == 4) r.adv_n(my_attn * (cross_nkc + cross_nvc));
          else lin_mma(l < p.NL ? L.lin[mega_lin_of_phase(ph)] : p.lm_head);
        }
      }
    }
  } else {
    // =============================================================================================== consumers
    const int ct = threadIdx.x - 64;  // 0..255
    const int cw = warp - 2;          // 0..7
    const bool is_worker = cw < 4;    // warps 2..5: activation staging + TMEM epilogue; warps 6..9: row statistics
    RingPos r;
    uint32_t n_item = 0;
    unsigned bar_target = 0;
    unsigned* bar_ctr = p.bar_ctr + (step & 1);

    int prof_i = 0;
    auto globaltimer = []() {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      return t;
    };
    auto grid_sync = [&]() {
      cons_sync();
      bar_target += (unsigned)G;
      if (ct == 0) {
        if (p.prof) p.prof[((size_t)g * 512 + prof_i) * 2] = globaltimer();
        __threadfence();
        atomicAdd(bar_ctr, 1u);
        unsigned v;
        const long long t0 = clock64();
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar_ctr) : "memory");
          if (v >= bar_target) break;
          if (clock64() - t0 > 4000000000LL) __trap();
        }
        if (p.prof) p.prof[((size_t)g * 512 + prof_i) * 2 + 1] = globaltimer();
        *reinterpret_cast<volatile int*>(S.s_b + 8) = prof_i + 1;  // consumers enter the next phase
      }
      ++prof_i;
      cons_sync();
    };

    // ---------------------------------------------------------------------------------------------- linear
    // out[b][n] (+)= rs[b] * sum_k pro(x)[b][k] * W[n][k];  pro: 0 none, 1 RMSNorm (x*lnw staged, rs in the epilogue),
    // 2 ReLU.  store: direct store (no split-K) + optional per-tile argmax partials.
    auto lin_phase = [&](const MegaLin W, int pro, const float* x, int ldx, float* out, int ld_out, const float* lnw,
                         float scale, float* zero_ptr, long long zero_n, bool store, float* amax_val, int* amax_idx) {
      const int items = W.tiles * W.ksplit;
      const bool have = g < items;
      if (!is_worker) {
        // ---- statistic warps (128 threads)
        const int t = ct - 128, wq = cw - 4;
        if (have) {
          if (pro == 1) {
            const int n4row = W.K >> 2;
            const int ca = min(t, n4row - 1), cb = min(t + 128, n4row - 1);
            const float wa = t < n4row ? 1.f : 0.f, wb = (t + 128) < n4row ? 1.f : 0.f;
            for (int r0 = 0; r0 < MK_R; r0 += 8) {
              float4 qa[8], qb[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float* xr = x + (int64_t)min(r0 + j, B - 1) * ldx;
                qa[j] = ldcg4(xr + 4 * ca);
                qb[j] = ldcg4(xr + 4 * cb);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float ss = wa * (qa[j].x * qa[j].x + qa[j].y * qa[j].y + qa[j].z * qa[j].z + qa[j].w * qa[j].w) +
                           wb * (qb[j].x * qb[j].x + qb[j].y * qb[j].y + qb[j].z * qb[j].z + qb[j].w * qb[j].w);
                ss = warp_sum(ss);
                if (lane == 0) S.s_part[(r0 + j) * 4 + wq] = ss;
              }
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (t < MK_R) {
              const float ss = (S.s_part[t * 4] + S.s_part[t * 4 + 1]) + (S.s_part[t * 4 + 2] + S.s_part[t * 4 + 3]);
              S.s_rs[t] = rsqrtf(ss / (float)W.K + p.eps) * scale;
            }
          } else if (t < MK_R) {
            S.s_rs[t] = scale;
          }
        }
#ifdef MK_FINE
        unsigned long long* fs = (p.prof && ct == 128 && prof_i >= 8 && prof_i < 16) ? p.prof + ((size_t)g * 512 + 256 + (prof_i - 8) * 16) * 2 : nullptr;
#endif
        MK_STAMP(fs, 10);
        cons_sync();  // row scales published (workers wait here before their first epilogue)
        if (zero_ptr) {  // zero duty, off the critical path (completes before this phase's grid barrier)
          const int64_t n4 = zero_n >> 2;
          float4* z4 = reinterpret_cast<float4*>(zero_ptr);
          for (int64_t i = (int64_t)g * 128 + t; i < n4; i += (int64_t)G * 128) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // keep this warp's view of the ring in step with the k-blocks the workers / MMA warp consume
        for (int it = g; it < items; it += G) {
          const int ks = it % W.ksplit;
          const int kb0 = ks * W.kb_per_item, kb1 = min(W.num_kb, kb0 + W.kb_per_item);
          r.adv_n(kb1 - kb0);
        }
        return;
      }
      // ---- workers (128 threads).  Compact rolled loops on purpose: this code runs once per phase and must stay
      // resident in the instruction cache (a fully unrolled version was ~10x slower, fetch-bound).
      const int t = ct;
      const int c4 = t & 15, r8 = t >> 4;
      const float relu_lo = (pro == 2) ? 0.f : -INFINITY;
      bool first = true;
#ifdef MK_FINE
      unsigned long long* fine = (p.prof && ct == 0 && prof_i >= 8 && prof_i < 16) ? p.prof + ((size_t)g * 512 + 256 + (prof_i - 8) * 16) * 2 : nullptr;
#endif
      MK_STAMP(fine, 0);
      for (int it = g; it < items; it += G) {
        const int tile = it / W.ksplit, ks = it - tile * W.ksplit;
        const int kb0 = ks * W.kb_per_item, kb1 = min(W.num_kb, kb0 + W.kb_per_item);
        float4 v[4], gw = make_float4(1.f, 1.f, 1.f, 1.f);
        {
          if (pro == 1) gw = *reinterpret_cast<const float4*>(lnw + kb0 * 64 + c4 * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = ldcg4(x + (int64_t)min(r8 + i * 8, B - 1) * ldx + kb0 * 64 + c4 * 4);
        }
#pragma unroll 1
        for (int kb = kb0; kb < kb1; ++kb) {
          float4 vn[4], gn = make_float4(1.f, 1.f, 1.f, 1.f);
          const int kn = min(kb + 1, kb1 - 1);  // prefetch the next k-block while this one is staged
          if (pro == 1) gn = *reinterpret_cast<const float4*>(lnw + kn * 64 + c4 * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) vn[i] = ldcg4(x + (int64_t)min(r8 + i * 8, B - 1) * ldx + kn * 64 + c4 * 4);
          mbar_wait_wd(&S.empty[r.s], r.ph ^ 1);
          if (kb == kb0 && v[3].w != 1.2345e-30f) MK_STAMP(fine, 1);
          uint8_t* xs_hi = S.ring + (size_t)r.s * MK_STAGE + MK_XOFF;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = r8 + i * 8;
            float4 w = v[i];
            if (rr >= B) w = make_float4(0.f, 0.f, 0.f, 0.f);
            w.x = fmaxf(w.x, relu_lo) * gw.x; w.y = fmaxf(w.y, relu_lo) * gw.y;
            w.z = fmaxf(w.z, relu_lo) * gw.z; w.w = fmaxf(w.w, relu_lo) * gw.w;
            bf16 h0, l0, h1, l1, h2, l2, h3, l3;
            split_bf16(w.x, h0, l0); split_bf16(w.y, h1, l1); split_bf16(w.z, h2, l2); split_bf16(w.w, h3, l3);
            // 128B swizzle: 16-byte chunk index XOR (row % 8); this float4 covers half a chunk (8 bytes)
            const uint32_t off = (uint32_t)rr * 128u + ((((uint32_t)c4 >> 1) ^ ((uint32_t)rr & 7u)) << 4) +
                                 (((uint32_t)c4 & 1u) << 3);
            uint2 ph2, pl2;
            ph2.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            ph2.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
            pl2.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            pl2.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
            *reinterpret_cast<uint2*>(xs_hi + off) = ph2;
            *reinterpret_cast<uint2*>(xs_hi + MK_XPLANE + off) = pl2;
          }
          fence_proxy_async();
          mbar_arrive(&S.xrdy[r.s]);
          r.adv();
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = vn[i];
          gw = gn;
        }
        MK_STAMP(fine, 2);
        if (first) {
          cons_sync();  // row scales from the statistic warps
          first = false;
        }
        MK_STAMP(fine, 3);
        // ---- epilogue of this item: TMEM -> registers in 8-column chunks (thread = output feature)
        mbar_wait_wd(S.tmem_full, n_item & 1);
        MK_STAMP(fine, 4);
        tc_fence_after();
        const int q = warp & 3;
        const int n = tile * 128 + q * 32 + lane;
        const bool n_ok = n < W.N;
        float* o = out + n;
#pragma unroll 1
        for (int c0 = 0; c0 < MK_R; c0 += 8) {
          uint32_t rr[8];
          __syncwarp();
#ifdef MK_FINE
          if (p.dbg & 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rr[j] = 0x3f800000u + c0;
          } else
#endif
          {
            tmem_ld_32x32_x8(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, rr);
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int b = c0 + j;
            const float val = __uint_as_float(rr[j]) * S.s_rs[b];
#ifdef MK_FINE
            if (p.dbg & 1) {
              if (val == 1.2345e-30f) o[0] = val;
            } else
#endif
            if (n_ok && b < B) {
              if (store) o[(int64_t)b * ld_out] = val; else atomicAdd(o + (int64_t)b * ld_out, val);
            }
            if (amax_val) {
              // per-row (max, first argmax) over the warp's 32 features: order-preserving integer key + redux
              const uint32_t u = __float_as_uint(n_ok ? val : -INFINITY);
              const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
              const uint32_t mk = __reduce_max_sync(0xffffffffu, key);
              const uint32_t bal = __ballot_sync(0xffffffffu, key == mk);
              if (lane == 0) {
                S.s_part[b * 4 + q] = __uint_as_float((mk & 0x80000000u) ? (mk & 0x7fffffffu) : ~mk);
                S.s_pi[b * 4 + q] = tile * 128 + q * 32 + (__ffs(bal) - 1);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(S.tmem_empty);
        MK_STAMP(fine, 5);
        if (amax_val) {
          asm volatile("bar.sync 3, 128;" ::: "memory");
          if (t < B) {
            float bv = S.s_part[t * 4];
            int bi = S.s_pi[t * 4];
#pragma unroll
            for (int w2 = 1; w2 < 4; ++w2) {
              const float ov = S.s_part[t * 4 + w2];
              const int oi = S.s_pi[t * 4 + w2];
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            amax_val[(int64_t)t * W.tiles + tile] = bv;
            amax_idx[(int64_t)t * W.tiles + tile] = bi;
          }
          asm volatile("bar.sync 3, 128;" ::: "memory");  // s_part reusable by the next item
        }
        ++n_item;
      }
      if (first) cons_sync();
    };

    // block-wide (256 consumer threads) max / sum through s_b
    auto block_max = [&](float v) {
      v = warp_max(v);
      if (lane == 0) S.s_b[cw] = v;
      cons_sync();
      float m = S.s_b[0];
#pragma unroll
      for (int w2 = 1; w2 < 8; ++w2) m = fmaxf(m, S.s_b[w2]);
      cons_sync();
      return m;
    };
    auto block_sum = [&](float v) {
      v = warp_sum(v);
      if (lane == 0) S.s_b[cw] = v;
      cons_sync();
      float m = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) m += S.s_b[w2];
      cons_sync();
      return m;
    };

    // ---------------------------------------------------------------------------------------------- self-attention
    // fused KV-cache append + single-query attention with the T5 unidirectional bucket bias (no 1/sqrt(d) scale).
    // K cache: [b][h][key/32][64 d][32 keys] (one contiguous block per (image, head), conflict-free thread = key),
    // V cache: [b][h][key][64 d].
    auto self_phase = [&](const MegaLayer& L) {
      const int r16 = ct >> 4, c16 = ct & 15;
      for (int it = g; it < n_attn; it += G) {
        const int b = it / H, h = it - b * H;
        float* kblk = L.skb + (size_t)it * Tb * 2048;
        float* vrow = L.svb + (size_t)it * p.Tp * 64;
        if (ct < 64) {
          const float* qp = p.qkv + (int64_t)b * 3 * D + h * 64 + ct;
          const float qv = __ldcg(qp), kn = __ldcg(qp + D), vn = __ldcg(qp + 2 * D);
          S.sq[ct] = qv;
          S.snew[ct] = vn;
          kblk[(size_t)(step >> 5) * 2048 + ct * 32 + (step & 31)] = kn;  // append
          vrow[(size_t)step * 64 + ct] = vn;
          S.sred[ct] = qv * kn;
        }
        cons_sync();
        // scores over the cached keys
        for (int c = 0; c < self_nkc; ++c) {
          mbar_wait_wd(&S.full[r.s], r.ph);
          const float* buf = reinterpret_cast<const float*>(S.ring + (size_t)r.s * MK_STAGE);
          const int nb = min(MK_SELF_KB, self_nblk - c * MK_SELF_KB);
          // warp -> (block = cw & 3, d half = cw >> 2): two partial sums per key, combined below
          const int blk = cw & 3, dh = cw >> 2;
          if (blk < nb) {
            const float* kp = buf + blk * 2048 + dh * 32 * 32 + lane;
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 32; ++d) a += S.sq[dh * 32 + d] * kp[d * 32];
            const int key = (c * MK_SELF_KB + blk) * 32 + lane;
            if (dh == 0) S.sc[key] = a; else S.sc[1024 + key] = a;  // sc[1024..]: partials of the second d half (Tp <= 1024)
          }
          cons_sync();
          if (ct == 0) mbar_arrive(&S.empty[r.s]);
          r.adv();
        }
        // bias, max
        float mx = -INFINITY;
        for (int j = ct; j <= step; j += 256) {
          float s;
          if (j < step) {
            s = S.sc[j] + S.sc[1024 + j];
          } else {
            s = 0.f;
            for (int d = 0; d < 64; ++d) s += S.sred[d];
          }
          s += p.dec_bias[p.lut[step - j] * H + h];
          S.sc[j] = s;
          mx = fmaxf(mx, s);
        }
        mx = block_max(mx);
        float sum = 0.f;
        for (int j = ct; j <= step; j += 256) {
          const float e = expf(S.sc[j] - mx);
          S.sc[j] = e;
          sum += e;
        }
        sum = block_sum(sum);  // also publishes sc[]
        const float inv = 1.f / sum;
        // P.V
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < self_nvc; ++c) {
          mbar_wait_wd(&S.full[r.s], r.ph);
          const float4* buf4 = reinterpret_cast<const float4*>(S.ring + (size_t)r.s * MK_STAGE);
          const int m0 = c * MK_SELF_VR, rows = min(MK_SELF_VR, step - m0);
#pragma unroll
          for (int j = 0; j < MK_SELF_VR / 16; ++j) {
            const int jj = r16 + 16 * j;
            if (jj < rows) {
              const float4 vv = buf4[jj * 16 + c16];
              const float pj = S.sc[m0 + jj];
              acc.x += pj * vv.x; acc.y += pj * vv.y; acc.z += pj * vv.z; acc.w += pj * vv.w;
            }
          }
          cons_sync();
          if (ct == 0) mbar_arrive(&S.empty[r.s]);
          r.adv();
        }
        if (r16 == 0) {
          const float pj = S.sc[step];
          acc.x += pj * S.snew[4 * c16]; acc.y += pj * S.snew[4 * c16 + 1];
          acc.z += pj * S.snew[4 * c16 + 2]; acc.w += pj * S.snew[4 * c16 + 3];
        }
        cons_sync();
        reinterpret_cast<float4*>(S.sred)[r16 * 16 + c16] = acc;
        cons_sync();
        if (ct < 64) {
          float o = 0.f;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) o += S.sred[rr * 64 + ct];
          p.ctx[(int64_t)b * D + h * 64 + ct] = o * inv;
        }
        cons_sync();  // sq / snew / sred / sc reusable
      }
    };

    // ---------------------------------------------------------------------------------------------- cross-attention
    // K^T [b][h][64][Mp], V [b][h][Mp][64] (fp32, contiguous per (image, head)); additive mask (1-mask)*finfo.min,
    // no positional bias, no scale.
    auto cross_phase = [&](const MegaLayer& L) {
      const int r16 = ct >> 4, c16 = ct & 15;
      for (int it = g; it < n_attn; it += G) {
        const int b = it / H, h = it - b * H;
        if (ct < 64) S.sq[ct] = __ldcg(p.q + (int64_t)b * D + h * 64 + ct);
        cons_sync();
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int c = 0; c < cross_nkc; ++c) {
          mbar_wait_wd(&S.full[r.s], r.ph);
          const float* buf = reinterpret_cast<const float*>(S.ring + (size_t)r.s * MK_STAGE);
          const int r0 = c * cross_rk, rows = min(cross_rk, 64 - r0);
          for (int rr = 0; rr < rows; ++rr) {
            const float qd = S.sq[r0 + rr];
            const float* row = buf + rr * Mp;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int m = ct + 256 * i;
              if (m < Mp) acc[i] += qd * row[m];
            }
          }
          cons_sync();
          if (ct == 0) mbar_arrive(&S.empty[r.s]);
          r.adv();
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = ct + 256 * i;
          if (m < Mp) {
            acc[i] += (p.mem_mask[(int64_t)b * Mp + m] ? 0.f : -3.4028234663852886e38f);
            mx = fmaxf(mx, acc[i]);
          }
        }
        mx = block_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = ct + 256 * i;
          if (m < Mp) {
            const float e = expf(acc[i] - mx);
            S.sc[m] = e;
            sum += e;
          }
        }
        sum = block_sum(sum);
        const float inv = 1.f / sum;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < cross_nvc; ++c) {
          mbar_wait_wd(&S.full[r.s], r.ph);
          const float4* buf4 = reinterpret_cast<const float4*>(S.ring + (size_t)r.s * MK_STAGE);
          const int m0 = c * MK_CROSS_VR, rows = min(MK_CROSS_VR, Mp - m0);
#pragma unroll
          for (int j = 0; j < MK_CROSS_VR / 16; ++j) {
            const int jj = r16 + 16 * j;
            if (jj < rows) {
              const float4 vv = buf4[jj * 16 + c16];
              const float pj = S.sc[m0 + jj];
              a4.x += pj * vv.x; a4.y += pj * vv.y; a4.z += pj * vv.z; a4.w += pj * vv.w;
            }
          }
          cons_sync();
          if (ct == 0) mbar_arrive(&S.empty[r.s]);
          r.adv();
        }
        reinterpret_cast<float4*>(S.sred)[r16 * 16 + c16] = a4;
        cons_sync();
        if (ct < 64) {
          float o = 0.f;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) o += S.sred[rr * 64 + ct];
          p.ctx[(int64_t)b * D + h * 64 + ct] = o * inv;
        }
        cons_sync();
      }
    };

    // ---------------------------------------------------------------------------------------------- the program
    const int64_t nB = B;
    for (int l = 0; l <= p.NL; ++l) {
      const MegaLayer& L = p.layers[min(l, p.NL - 1)];
      const int nph = l < p.NL ? 8 : 1;
      for (int ph = 0; ph < nph; ++ph) {
        if (l < p.NL && ph == 1) {
          self_phase(L);
        } else if (l < p.NL && ph == 4) {
          cross_phase(L);
        } else {
          // the linears of a layer: (input, prologue, output, which buffer this phase zeroes for a later one)
          int pro = 1, ldx = D, ld_out = D;
          const float* xin = p.x;
          float* out = p.x;
          const float* lnw = nullptr;
          float scale = 1.f;
          float* zp = nullptr;
          long long zn = 0;
          bool store = false;
          if (l == p.NL) {  // LM head: final RMSNorm * d_model^-0.5 fused, direct store + per-tile argmax
            lnw = p.final_ln; scale = p.logit_scale; out = p.logits; ld_out = p.ld_logits; store = true;
          } else if (ph == 0) {  // x -> qkv (RMSNorm ln1); zero: FF hidden buffer
            lnw = L.ln[0]; out = p.qkv; ld_out = 3 * D; zp = p.hbuf; zn = nB * p.DFF;
          } else if (ph == 2 || ph == 5) {  // ctx -> x (+=): attention output projections
            pro = 0; xin = p.ctx;
          } else if (ph == 3) {  // x -> q (RMSNorm ln2); zero: qkv
            lnw = L.ln[1]; out = p.q; zp = p.qkv; zn = nB * 3 * D;
          } else if (ph == 6) {  // x -> hidden (RMSNorm ln3); zero: q
            lnw = L.ln[2]; out = p.hbuf; ld_out = p.DFF; zp = p.q; zn = nB * D;
          } else {  // ph == 7: relu(hidden) -> x (+=)
            pro = 2; xin = p.hbuf; ldx = p.DFF;
          }
          lin_phase(l < p.NL ? L.lin[mega_lin_of_phase(ph)] : p.lm_head, pro, xin, ldx, out, ld_out, lnw, scale, zp, zn,
                    store, store ? p.part_val : nullptr, store ? p.part_idx : nullptr);
        }
        if (l < p.NL) grid_sync();
      }
    }
    if (p.prof && ct == 0) p.prof[((size_t)g * 512 + prof_i) * 2] = globaltimer();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32);
}

// ------------------------------------------------------------------------------------------------ weight tiling
// planes [N][ldk] (hi, lo) -> [tile][k-block][hi: 128 rows x 128 B, 16-byte chunk c of row r at c ^ (r & 7)][lo ...]
// i.e. exactly the shared-memory image TMA's 128-byte swizzle would produce, so one 32 KB bulk copy fills a stage.
__global__ void tile_weights_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, int N, int K, int64_t ldk,
                                    int tiles, int num_kb, uint8_t* __restrict__ out) {
  const int64_t total = (int64_t)tiles * num_kb * 2 * 128 * 8;  // 16-byte chunks
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const int plane = (int)((i >> 10) & 1);
    const int64_t tk = i >> 11;
    const int kb = (int)(tk % num_kb), tile = (int)(tk / num_kb);
    const int n = tile * 128 + r;
    const int k0 = kb * 64 + c * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const bf16* src = plane ? lo : hi;
    if (n < N && k0 < K && src) v = *reinterpret_cast<const uint4*>(src + (int64_t)n * ldk + k0);
    uint8_t* dst = out + tk * MK_WTILE + plane * 16384 + r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

MegaLin make_mega_lin(cudaStream_t st, Planes w, int N, int K, int64_t ldk, bool store, int n_ctas, uint8_t* dst) {
  MegaLin L;
  L.N = N;
  L.K = K;
  L.tiles = (N + 127) / 128;
  L.num_kb = K / 64;
  int ksplit = 1;
  if (!store) ksplit = std::min(L.num_kb, std::max(1, n_ctas / L.tiles));
  L.kb_per_item = (L.num_kb + ksplit - 1) / ksplit;
  L.ksplit = (L.num_kb + L.kb_per_item - 1) / L.kb_per_item;
  L.w = dst;
  const int64_t total = (int64_t)L.tiles * L.num_kb * 2 * 128 * 8;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  tile_weights_kernel<<<blocks, 256, 0, st>>>(w.hi, w.lo, N, K, ldk, L.tiles, L.num_kb, dst);
  MG_CHECK_CUDA(cudaGetLastError());
  return L;
}

size_t mega_lin_bytes(int N, int K) { return (size_t)((N + 127) / 128) * (K / 64) * MK_WTILE; }

int mega_max_ctas() {
  static int n = -1;
  if (n < 0) {
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    MG_CHECK_CUDA(cudaGetDevice(&dev));
    MG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MG_CHECK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    MG_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MK_SMEM));
    MG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel, MK_THREADS, MK_SMEM));
    n = (coop && per_sm >= 1) ? sms : 0;
  }
  return n;
}

void launch_decode_step(cudaStream_t st, const MegaParams& p, int n_ctas) {
  MG_REQUIRE(p.B >= 1 && p.B <= MK_R, "fused decode step: 1 <= B <= 32");
  MG_REQUIRE(p.Mp % 4 == 0 && p.Mp <= MK_MAXSC && p.Tp % 32 == 0 && p.Tp <= 1024, "fused decode step: Mp / max_length out of range");
  MG_REQUIRE(p.D == p.H * 64 && p.D <= 1024 && p.D % 64 == 0 && p.DFF % 64 == 0, "fused decode step: unsupported dims");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(MK_THREADS);
  cfg.dynamicSmemBytes = MK_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_step_kernel, p));
}

}  // namespace mg
