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
// Two measured facts shape the code (tools/ubench/phase_bench.cu, profiles/r1_mega_*.txt):
//   * the step executes every piece of code once per phase, so the kernel is INSTRUCTION-FETCH bound unless its
//     whole hot path fits the SM's instruction cache (~32 KB): a first, fully unrolled / multiply-inlined version
//     (314 KB of SASS) ran every phase ~10x slower than its memory traffic explains.  Hence one call site per phase
//     kind, rolled loops, and the shared waits / reductions as __noinline__ functions.
//   * a saturated memory system multiplies the latency of everything else (L2 load 0.2 -> 1.7 us, a grid barrier
//     1.2 -> 12 us), so an attention phase's K/V stream starts only when the CTA's consumers enter that phase.
//
// Arithmetic follows transformers/models/udop/modeling_udop.py exactly like the unfused kernels in decode.cu /
// gemm_tc.cu (UdopLayerNorm :333-355, UdopAttention :431-622 incl. compute_bias :514-529, UdopLayerFF :412-427,
// decoder UdopStack :1146-1256, lm head :1585-1590); the selection stays in greedy_select_kernel.
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace mg {

#ifdef MK_PF_LANE
// experiment build (MG_B200_CFLAGS=-DMK_PF_LANE): an 11th warp whose lane 0 is a dedicated L2-prefetch lane, see below
constexpr int MK_THREADS = 352;
#else
constexpr int MK_THREADS = 320;
#endif
#ifndef MK_NST_N
#define MK_NST_N 5
#endif
#ifndef MK_XUNROLL
#define MK_XUNROLL 2  // d-rows (K pass) / key rows (V pass) of the cross-attention inner loops in flight per thread
#endif
#define MK_PRAGMA_(x) _Pragma(#x)
#define MK_PRAGMA(x) MK_PRAGMA_(x)
#ifndef MK_OCC
#define MK_OCC 1  // experiment build: -DMK_OCC=2 -DMK_NST_N=2 lets two instances (two half-batch lanes) share every SM
#endif
constexpr int MK_NST = MK_NST_N;
constexpr int MK_STAGE = 40960;
constexpr int MK_WTILE = 32768;  // one (tile, k-block): [hi 128x64 bf16 swizzled][lo ...]
constexpr int MK_XOFF = 32768;   // activation tile inside a stage: hi [32][64] bf16 (4 KB) then lo (4 KB)
constexpr int MK_XPLANE = 4096;
constexpr int MK_R = 32;         // activation rows (UMMA N)
constexpr int MK_MAXSC = 2048;   // max keys of one attention row (cross: Mp, self: 2 x max_length)
constexpr int MK_SELF_KB = 4;    // 32-key blocks per self-K chunk (32 KB)
constexpr int MK_SELF_VR = 128;  // keys per self-V chunk (32 KB)
constexpr int MK_SELF_NG = 4;    // self-attention items a CTA processes concurrently (2 warps each)
constexpr int MK_CROSS_VR = 208;  // keys per cross-V chunk (kv24: 192 bytes per key, 39936 bytes per chunk)

// in-kernel globaltimer stamps, profiling builds only (MG_B200_CFLAGS=-DMK_FINE, MG_MEGA_PROF=<file>);
// slot layout in tools/mega_phase_profile.py
#ifdef MK_FINE
#define MK_STAMP(ptr, i) do { if (ptr) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); (ptr)[i] = _t; } } while (0)
#else
#define MK_STAMP(ptr, i) do { } while (0)
#endif

// cycle accounting of the cross-attention phase, profiling build only (MG_B200_CFLAGS=-DMK_XPROF, MG_MEGA_PROF=<file>,
// reader: tools/cross_phase_cycles.py): per CTA, layer NL/2 of the last step -- producer: cycles waiting for a free
// stage / issuing; consumers (thread 0): cycles waiting for data / in the arithmetic / in the barrier + release, and the
// per-item sections outside the chunk loops (head: query + mask loads, softmax, tail: output reduction)
#ifdef MK_XPROF
#define MK_XP(...) __VA_ARGS__
#else
#define MK_XP(...)
#endif

// CTA-local progress flags (producer -> consumers "loads issued" counter, consumers -> producer phase counter) are
// plain volatile shared words: single writer, monotonic, readers spin.  compute-sanitizer racecheck reports every
// spin read as a hazard (10^7 of them); the -DMK_RACECHECK build routes the flag accesses through shared-memory
// atomics, which racecheck treats as synchronising, so that any OTHER hazard becomes visible.
#ifdef MK_RACECHECK
#define MK_FLAG_LD(p) atomicAdd(const_cast<int*>(p), 0)
#define MK_FLAG_ST(p, v) atomicExch(const_cast<int*>(p), (v))
#else
#define MK_FLAG_LD(p) (*(p))
#define MK_FLAG_ST(p, v) (*(p) = (v))
#endif

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// mbarrier wait shared by every role (one copy in the instruction cache); bounded: a hang becomes a trap
// (-> launch failure) instead of a dead GPU
// watchdog diagnostics: host-mapped pinned words written just before the trap
__device__ int* g_mk_dbg = nullptr;
__device__ __noinline__ void mk_die(int code, uint32_t a, uint32_t b) {
  int* d = g_mk_dbg;
  if (d && atomicCAS(d, 0, code) == 0) {
    d[1] = blockIdx.x; d[2] = threadIdx.x; d[3] = (int)a; d[4] = (int)b;
    __threadfence_system();
  }
  __trap();
}
__device__ __noinline__ void mk_wait(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) mk_die(1, bar, parity);
  } while (!done);
}
__device__ __forceinline__ void mk_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// block-wide (256 consumer threads) max / sum through 8 shared floats
__device__ __noinline__ float mk_block_reduce(float v, float* s_b, int cw, int lane, int is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  if (lane == 0) s_b[cw] = v;
  cons_sync();
  float m = s_b[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = is_max ? fmaxf(m, s_b[w]) : m + s_b[w];
  cons_sync();
  return m;
}

struct RingPos {
  int s = 0;
  uint32_t ph = 0;
  uint32_t xmask = 0;  // per-stage parity of the "activation tile staged" barrier (flips on linear uses only)
  int n = 0;           // loads of this CTA's program before this position
  __device__ __forceinline__ void adv() {
    if (++s == MK_NST) { s = 0; ph ^= 1; }
    ++n;
  }
  __device__ __forceinline__ void adv_n(int k) {
    const int t = s + k;
    ph ^= (uint32_t)((t / MK_NST) & 1);
    s = t % MK_NST;
    n += k;
  }
};

// Phases of a decoder layer: 0 qkv (+ the x part of the cross query), 1 self-attention, 2 o (+ the ctx part of the
// cross query), 3 cross-attention, 4 co, 5 wi, 6 wo.  The cross query q = Wcq' (x + Wo ctx), Wcq' = Wcq diag(ln2), is
// linear in its two inputs, so it is accumulated as Wcq' x (extra rows of phase 0, same activation tile as q/k/v) plus
// (Wcq' Wo) ctx (extra rows of phase 2, same activation tile as o; the 1024^2 product is precomputed at finalize):
// the separate "cq" phase of the kernel chain -- a whole latency-bound linear and its grid barrier -- is gone.
// -DMK_FOLD_FF: the cross-attention output projection is folded into the FF input projection as well,
//   h = Wi' (x1 + Wco ctx) = Wi' x1 + (Wi' Wco) ctx,   Wi' = Wi diag(ln3),
// so phases 4 (co) and 5 (wi) become ONE three-segment linear -- rows [0, d): Wco on ctx -> dx; rows [d, d + dff):
// (Wi' Wco) on ctx -> hidden; rows [d + dff, d + 2 dff): Wi' on the raw residual x1 -> hidden -- and the layer has 6
// phases: 0 qkv|cq, 1 self, 2 o|cq, 3 cross, 4 co|wi, 5 wo.  x1 must stay intact while phase 4 reads it, so Wco ctx
// goes to a side buffer dx and the wo phase writes the NEXT residual x3 = (x1 + dx) + rs * Wwo relu(h) into the other
// of two residual buffers (k-slice 0 of every output tile carries x1 + dx); rs = rsqrt(mean((x1 + dx)^2) + eps) is
// computed by every CTA's statistic warps during the wo phase and applied in its epilogue.
#ifdef MK_FOLD_FF
constexpr int MK_NPH = 6;
#else
constexpr int MK_NPH = 7;
#endif
constexpr int MK_PH_SELF = 1, MK_PH_CROSS = 3;
__device__ __forceinline__ int mega_lin_of_phase(int ph) { return ph == 0 ? 0 : (ph == 2 ? 1 : ph - 2); }  // 0,2,4,5,6 -> 0..4
__device__ __forceinline__ int items_of_cta(int total, int g, int G) { return g < total ? (total - g + G - 1) / G : 0; }
// k-blocks of work item `it` of a linear
__device__ __forceinline__ int lin_item_kbs(const MegaLin& W, int it, int& tile, int& kb0) {
  tile = it / W.ksplit;
  kb0 = (it - tile * W.ksplit) * W.kb_per_item;
  return min(W.num_kb, kb0 + W.kb_per_item) - kb0;
}

// shared memory after the ring
constexpr int MK_OFF_SC = MK_NST * MK_STAGE;        // float [MK_MAXSC]
constexpr int MK_OFF_SQ = MK_OFF_SC + MK_MAXSC * 4;  // float [64]
constexpr int MK_OFF_SNEW = MK_OFF_SQ + 256;         // float [64]
constexpr int MK_OFF_SRED = MK_OFF_SNEW + 256;       // float [16*64]
constexpr int MK_OFF_RS = MK_OFF_SRED + 4096;        // float [32]
constexpr int MK_OFF_PART = MK_OFF_RS + 128;         // float [32*4]
constexpr int MK_OFF_PI = MK_OFF_PART + 512;         // int   [32*4]
constexpr int MK_OFF_SB = MK_OFF_PI + 512;           // float [8] + int phase flag at [8]
constexpr int MK_OFF_BAR = MK_OFF_SB + 64;           // full[NST] empty[NST] xrdy[NST] tmem_full tmem_empty
constexpr int MK_OFF_SLOT = MK_OFF_BAR + (3 * MK_NST + 2) * 8;
constexpr int MK_MAX_LAYERS = 24;
constexpr int MK_OFF_TAB = MK_OFF_SLOT + 16;        // MegaLayer [MK_MAX_LAYERS]: the per-layer table, copied once
constexpr int MK_SMEM = MK_OFF_TAB + MK_MAX_LAYERS * (int)sizeof(MegaLayer) + 1024;
static_assert(sizeof(MegaLayer) % 8 == 0 && MK_OFF_TAB % 8 == 0, "layer table is copied in 8-byte words");
static_assert(MK_SMEM <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(MK_THREADS, MK_OCC) decode_step_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic (not an integer round trip) keeps the shared address space: LDS/STS instead of generic LD/ST
  uint8_t* const ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* const s_sc = reinterpret_cast<float*>(ring + MK_OFF_SC);
  float* const s_q = reinterpret_cast<float*>(ring + MK_OFF_SQ);
  float* const s_red = reinterpret_cast<float*>(ring + MK_OFF_SRED);
  float* const s_rs = reinterpret_cast<float*>(ring + MK_OFF_RS);
  float* const s_part = reinterpret_cast<float*>(ring + MK_OFF_PART);
  int* const s_pi = reinterpret_cast<int*>(ring + MK_OFF_PI);
  float* const s_b = reinterpret_cast<float*>(ring + MK_OFF_SB);
  volatile int* const s_phase = reinterpret_cast<volatile int*>(ring + MK_OFF_SB + 32);
  volatile int* const s_issued = reinterpret_cast<volatile int*>(ring + MK_OFF_SB + 36);  // loads the producer has issued
  int* const s_rel = reinterpret_cast<int*>(ring + MK_OFF_SB + 40);  // [MK_NST] cross-attention: consumer warps done with a stage
  const uint32_t ring_a = smem_u32(ring);
  const uint32_t bar_full = ring_a + MK_OFF_BAR, bar_empty = bar_full + 8 * MK_NST, bar_xrdy = bar_empty + 8 * MK_NST;
  const uint32_t bar_tfull = bar_xrdy + 8 * MK_NST, bar_tempty = bar_tfull + 8;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(ring + MK_OFF_SLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, G = gridDim.x;
  const int step = *p.step_ptr;  // tokens already in the self-attention caches
  const int H = p.H, D = p.D, B = p.B, Mp = p.Mp, NL = p.NL;
  const int n_attn = B * H;

  // the layer table lives in shared memory: its fields are read on the critical path of every phase
  const MegaLayer* const s_layers = reinterpret_cast<const MegaLayer*>(ring + MK_OFF_TAB);
  {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(p.layers);
    uint64_t* dst = reinterpret_cast<uint64_t*>(ring + MK_OFF_TAB);
    for (int i = threadIdx.x; i < NL * (int)(sizeof(MegaLayer) / 8); i += MK_THREADS) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < MK_NST; ++s) {
      mbar_init(reinterpret_cast<uint64_t*>(ring + MK_OFF_BAR) + s, 1);
      mbar_init(reinterpret_cast<uint64_t*>(ring + MK_OFF_BAR) + MK_NST + s, 1);
      mbar_init(reinterpret_cast<uint64_t*>(ring + MK_OFF_BAR) + 2 * MK_NST + s, 128);
    }
    mbar_init(reinterpret_cast<uint64_t*>(ring + MK_OFF_BAR) + 3 * MK_NST, 1);
    mbar_init(reinterpret_cast<uint64_t*>(ring + MK_OFF_BAR) + 3 * MK_NST + 1, 256);
    *s_phase = 0;  // phase the consumers have entered (prefetch gate)
    *s_issued = 0;
    for (int i = 0; i < MK_NST; ++i) s_rel[i] = 0;
    fence_mbar_init();
    if (g == 0) p.bar_ctr[(step + 1) & 1] = 0u;  // the other parity's counter is idle during this launch
    if (g == 0) g_mk_dbg = p.dbg_host;
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // per-item chunk counts of the two attention phases (identical for every item of a phase)
  const int self_nblk = (step + 31) >> 5;
  const int self_nkc = (self_nblk + MK_SELF_KB - 1) / MK_SELF_KB;
  const int self_nvc = (step + MK_SELF_VR - 1) / MK_SELF_VR;
  // d-rows per K chunk: both planes' copies must be 16-byte sized and aligned -- any count if Mp is a multiple of 16 (the
  // compacted memory is padded to that), else an even one
  int cross_rk = (Mp & 15) == 0 ? min(64, MK_STAGE / (Mp * 3)) : min(64, (MK_STAGE / (Mp * 3)) & ~1);
  int cross_vr = MK_CROSS_VR;
  // experiment switches (MG_MEGA_CROSS_RK / MG_MEGA_CROSS_VR): smaller chunks = emptier ring stages
  if (p.cross_rk > 0) cross_rk = max(2, min(cross_rk, p.cross_rk & ~1));
  if (p.cross_vr > 0) cross_vr = max(16, min(cross_vr, p.cross_vr & ~7));
  const int cross_nkc = (64 + cross_rk - 1) / cross_rk;
  const int cross_nvc = (Mp + cross_vr - 1) / cross_vr;
  // attention items are spread over Ga <= G CTAs so that every one of them gets the same count (512 items over 148
  // CTAs would be 3 or 4 each and the phase would run at the pace of 4; over 128 CTAs it is exactly 4 each and HBM,
  // not the SM count, stays the limit)
#ifdef MK_SELF_ALL
  // experiment build (-DMK_SELF_ALL): self-attention items over ALL CTAs (3 or 4 each at 512 items / 148 CTAs); the
  // phase still ends with the 4-item CTAs, but they compete with fewer streams towards its end
  const int Ga = min(G, n_attn);
#else
  const int Ga = (n_attn + (n_attn + G - 1) / G - 1) / ((n_attn + G - 1) / G);
#endif
  const int my_attn = g < Ga ? items_of_cta(n_attn, g, Ga) : 0;
  const int attn_first = g < Ga ? g : n_attn;
  const int Tb = p.Tp >> 5;  // 32-key blocks per (image, head) of the self K cache
  // Cross-attention is balanced by BYTES over all CTAs: an item's K pass and V pass stream equal amounts, so the
  // 2 * n_attn half-works  K0 V0 K1 V1 ...  are cut into G equal ranges.  A range that ends with a K pass does it
  // FIRST and publishes the unnormalised probabilities (Mp floats + their sum) through global memory; the next CTA,
  // whose range starts with that item's V pass, does it LAST -- so the one cross-CTA dependency never waits.
  const int xh_total = 2 * n_attn, xh_per = (xh_total + G - 1) / G;
  const int xh_lo = min(g * xh_per, xh_total), xh_hi = min(xh_lo + xh_per, xh_total);
  const int x_lone_k = (xh_hi > xh_lo && ((xh_hi - 1) & 1) == 0) ? 1 : 0;  // range ends with a K pass
  const int x_lone_v = (xh_hi > xh_lo && (xh_lo & 1) == 1) ? 1 : 0;        // range starts with a V pass
  const int x_first = (xh_lo + 1) >> 1;                                     // whole items [x_first, x_first + x_whole)
  const int x_whole = max(((xh_hi - x_lone_k) >> 1) - x_first, 0);
  const int x_entries = x_lone_k + x_whole + x_lone_v;
  // entry e of this CTA's cross schedule -> (item, passes: 1 = K, 2 = V, 3 = both)
  auto cross_entry = [&](int e, int& item, int& passes) {
    if (x_lone_k && e == 0) { item = (xh_hi - 1) >> 1; passes = 1; }
    else if (e - x_lone_k < x_whole) { item = x_first + e - x_lone_k; passes = 3; }
    else { item = xh_lo >> 1; passes = 2; }
  };

  if (warp == 0) {
    // =============================================================================================== producer
    // Every phase is described as "n items, each item = up to two contiguous byte streams cut into chunks", so
    // one loop issues all loads of the step.
    if (lane == 0) {
      RingPos r;
      int n_put = 0;
      const int max_inflight = min(p.max_inflight, MK_NST);
      // everything streamed through the ring is read exactly once per step: evict-first keeps the small hot data
      // (activations, masks, norm weights, bias tables) resident in L2 underneath a ~9 GB/step stream
      uint64_t pol_stream;
      if (p.dbg & 4)
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
      else
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      // Paced L2 prefetch of the NEXT cross-attention phase (p.l2pf bytes per CTA and layer, 0 = off): while this
      // thread is blocked on a ring slot and its program position is in a phase selected by p.l2pf_mask (default: the
      // qkv / self-attention / o loads, i.e. the two or three phases right before the cross phase -- lines fetched
      // earlier do not survive the evict-first weight and self-K/V streams), it asks L2 for the next p.l2pf_piece
      // bytes of the range the CTA streams first in the coming cross phase, at most once per p.l2pf_gap cycles, so
      // the prefetch never forms a burst in front of the latency-critical activation loads of the linears (an
      // un-paced version issued on phase entry cost exactly the HBM time of the prefetched bytes).
      // The K/V blocks of consecutive items are contiguous in memory, so the CTA's half-works are plain byte ranges;
      // the cursor follows the order the CTA consumes them.  Measured (DESIGN.md decision 9): -0.75 % step time from
      // spinning on test_wait instead of the suspending try_wait in these phases (the self-attention phase's small
      // dependent chunks get their slots re-issued sooner), another -0.3 % from the prefetch; -1.6 % p50.
      const uint32_t pf_hw = (uint32_t)Mp * 192u;  // bytes of one half-work (K or V block of one item)
      const uint32_t pf_total = min((uint32_t)p.l2pf, (uint32_t)(xh_hi - xh_lo) * pf_hw);
      uint32_t pf_cur = 0;
      long long pf_last = 0;
      const uint8_t* pf_ckv = nullptr;
      bool pf_on = false;
      const bool poll_all = (p.l2pf_mask & 0x100) != 0;  // busy-poll (test_wait) for a free slot in every phase
      auto wait_slot = [&](uint32_t bar, uint32_t parity) {
        if (!(poll_all || (pf_on && pf_cur < pf_total))) { mk_wait(bar, parity); return; }
        uint32_t done, spins = 0;
        for (;;) {
          asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                       : "=r"(done) : "r"(bar), "r"(parity) : "memory");
          if (done) break;
          if (pf_on && pf_cur < pf_total) {
            const long long now = clock64();
            if (p.l2pf_piece > 0 && now - pf_last >= (long long)p.l2pf_gap) {  // piece 0: poll only (A/B of the wait style)
              pf_last = now;
              const uint32_t j = pf_cur / pf_hw, off = pf_cur - j * pf_hw;
              int hw;  // j-th half-work in consumption order -> index in memory order
              if (x_lone_k && j == 0) hw = xh_hi - 1;
              else if ((int)j - x_lone_k < 2 * x_whole) hw = 2 * x_first + (int)j - x_lone_k;
              else hw = xh_lo;
              const uint32_t n = min(min((uint32_t)p.l2pf_piece, pf_total - pf_cur), pf_hw - off);
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(pf_ckv + (size_t)hw * pf_hw + off)), "r"(n) : "memory");
              pf_cur += n;
            }
          }
          if (++spins > (1u << 26)) mk_die(1, bar, parity);
        }
      };
      MK_XP(long long xp_wait = 0; long long xp_issue = 0; long long xp_n = 0;)
      for (int l = 0; l <= NL; ++l) {
        const MegaLayer& L = s_layers[min(l, NL - 1)];
        const int nph = l < NL ? MK_NPH : 1;
        for (int ph = 0; ph < nph; ++ph) {
          MK_XP(if (p.prof && l == NL / 2 && ph == MK_PH_CROSS + 1) {
            unsigned long long* q = p.prof + ((size_t)g * 512 + 440) * 2;
            q[0] = xp_wait; q[1] = xp_issue; q[2] = xp_n;
          })
          const bool attn = l < NL && (ph == MK_PH_SELF || ph == MK_PH_CROSS);
          if (pf_total) {
            if (l < NL && ph == MK_PH_CROSS) { pf_cur = 0; pf_on = false; }
            else {
              const int lt = (l < NL && ph < MK_PH_CROSS) ? l : l + 1;  // layer of the next cross phase (wraps to the next step)
              pf_ckv = s_layers[lt < NL ? lt : 0].ckv;
              pf_on = l == NL || ((p.l2pf_mask >> ph) & 1);
            }
          }
          int n_items, div = 1, units = 1, upi = 1;  // item -> (major = it / div, minor = it % div)
          const uint8_t* base0;
          const uint8_t* base1 = nullptr;
          size_t sa0, sb0 = 0, sa1 = 0;              // stream strides (bytes) along major / minor
          uint32_t tot0, tot1 = 0, chunk0, chunk1 = MK_STAGE;
          uint32_t lo_off = 0;  // kv24 streams: byte distance from the 16-bit plane to the 8-bit plane (0 = plain stream)
          if (attn) {
            if (p.gate) {
              // prefetch gate: this phase's K/V stream starts when the CTA's consumers have entered the phase
              const int want = l * MK_NPH + ph;
              const long long t0 = clock64();
              while (MK_FLAG_LD(s_phase) < want)
                if (clock64() - t0 > 4000000000LL) mk_die(2, want, *s_phase);
            }
            n_items = n_attn;
            if (ph == MK_PH_SELF) {
              base0 = reinterpret_cast<const uint8_t*>(L.skb); sa0 = (size_t)Tb * 8192; tot0 = (uint32_t)self_nblk * 8192u;
              chunk0 = MK_SELF_KB * 8192;
              base1 = reinterpret_cast<const uint8_t*>(L.svb); sa1 = (size_t)p.Tp * 256; tot1 = (uint32_t)step * 256u;
              chunk1 = MK_SELF_VR * 256;
            } else {
              // kv24 block per (image, head): [K hi 2n][K lo n][V hi 2n][V lo n], n = 64 * Mp; tot / chunk count the
              // 16-bit plane, every chunk is followed in its stage by the matching half-sized piece of the 8-bit plane
              base0 = L.ckv; sa0 = (size_t)Mp * 384; tot0 = (uint32_t)Mp * 128u; chunk0 = (uint32_t)cross_rk * Mp * 2u;
              base1 = L.ckv + (size_t)Mp * 192; sa1 = sa0; tot1 = tot0; chunk1 = (uint32_t)cross_vr * 128u;
              lo_off = (uint32_t)Mp * 128u;
            }
          } else {
            const MegaLin& W = l < NL ? L.lin[mega_lin_of_phase(ph)] : p.lm_head;
            n_items = W.tiles * W.ksplit; div = W.ksplit; units = W.num_kb; upi = W.kb_per_item;
            base0 = W.w; sa0 = (size_t)W.num_kb * MK_WTILE; sb0 = (size_t)W.kb_per_item * MK_WTILE;
            tot0 = 0; chunk0 = MK_WTILE;
          }
          // items are issued in bundles: the self-attention phase processes MK_SELF_NG items concurrently, so their
          // chunks are interleaved (chunk c of every item of the bundle, then chunk c+1, ...); elsewhere bundle = 1
          const int stride = attn ? Ga : G;
          const int bundle = (attn && ph == MK_PH_SELF) ? MK_SELF_NG : 1;
          const bool is_cross = attn && ph == MK_PH_CROSS;
          for (int e = 0, it0 = attn ? attn_first : g; is_cross ? e < x_entries : it0 < n_items; ++e, it0 += stride * bundle) {
            int passes = 3;
            if (is_cross) cross_entry(e, it0, passes);
            for (int sidx = 0; sidx < 2; ++sidx) {
              if (sidx && !base1) break;
              if (!((passes >> sidx) & 1)) continue;
              const uint32_t chunk = sidx ? chunk1 : chunk0;
              const int minor0 = it0 % div;
              const uint32_t tot = sidx ? tot1 : (attn ? tot0 : (uint32_t)min(upi, units - minor0 * upi) * MK_WTILE);
              for (uint32_t off = 0; off < tot; off += chunk) {
                const uint32_t bytes = min(chunk, tot - off);
                for (int gi = 0; gi < bundle; ++gi) {
                  const int it = it0 + gi * stride;
                  if (it >= n_items) break;
                  const int major = it / div, minor = it - major * div;
                  const uint8_t* const src0 = sidx ? base1 + major * sa1 : base0 + major * sa0 + minor * sb0;
                  // cap the loads this SM keeps in flight (a deeper queue adds latency, not bandwidth)
                  MK_XP(const long long xt0 = clock64();)
                  if (n_put >= max_inflight) {
                    const int m = n_put - max_inflight;
                    mk_wait(bar_full + 8 * (m % MK_NST), (uint32_t)((m / MK_NST) & 1));
                  }
                  ++n_put;
                  wait_slot(bar_empty + 8 * r.s, r.ph ^ 1);
                  MK_XP(const long long xt1 = clock64();)
                  const uint32_t fb = bar_full + 8 * r.s;
                  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(lo_off ? bytes + (bytes >> 1) : bytes) : "memory");
                  asm volatile(
                      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                          ring_a + (uint32_t)r.s * MK_STAGE),
                      "l"(reinterpret_cast<uint64_t>(src0 + off)), "r"(bytes), "r"(fb), "l"(pol_stream)
                      : "memory");
                  if (lo_off)
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                            ring_a + (uint32_t)r.s * MK_STAGE + bytes),
                        "l"(reinterpret_cast<uint64_t>(src0 + lo_off + (off >> 1))), "r"(bytes >> 1), "r"(fb), "l"(pol_stream)
                        : "memory");
                  r.adv();
                  __threadfence_block();
                  MK_FLAG_ST(s_issued, n_put);
                  MK_XP(if (is_cross && l == NL / 2) { xp_wait += xt1 - xt0; xp_issue += clock64() - xt1; ++xp_n; })
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================================================== MMA issuer
    // The whole warp runs the (warp-uniform) control flow and one elected lane issues: with a divergent
    // `if (lane == 0)` the compiler cannot keep descriptors in uniform registers and wraps every tcgen05.mma in an
    // ELECT / R2UR loop (~17 instructions per MMA on the critical path of every linear).
    {
      RingPos r;
      uint32_t n_item = 0;
      constexpr uint32_t idesc = make_idesc_bf16(128, MK_R);
#ifdef MK_FINE
      int mma_phase = 0;
#endif
      for (int l = 0; l <= NL; ++l) {
        const MegaLayer& L = s_layers[min(l, NL - 1)];
        const int nph = l < NL ? MK_NPH : 1;
        for (int ph = 0; ph < nph; ++ph) {
          if (l < NL && ph == MK_PH_SELF) { r.adv_n(my_attn * (self_nkc + self_nvc)); continue; }
          if (l < NL && ph == MK_PH_CROSS) { r.adv_n((x_whole + x_lone_k) * cross_nkc + (x_whole + x_lone_v) * cross_nvc); continue; }
          const MegaLin& W = l < NL ? L.lin[mega_lin_of_phase(ph)] : p.lm_head;
          const int items = W.tiles * W.ksplit;
#ifdef MK_FINE
          unsigned long long* fm = (p.prof && mma_phase >= 5 && mma_phase < 10) ? p.prof + ((size_t)g * 512 + 400 + (mma_phase - 5) * 4) * 2 : nullptr;
          ++mma_phase;
#endif
          for (int it = g; it < items; it += G) {
            int tile, kb0;
            const int nkb = lin_item_kbs(W, it, tile, kb0);
            if (n_item > 0) mk_wait(bar_tempty, (n_item - 1) & 1);  // epilogue of the previous item drained TMEM
            tc_fence_after();
            uint32_t acc = 0;
            for (int kb = 0; kb < nkb; ++kb) {
              // activation tile first: once it is staged, the stage's previous occupant has been consumed, so the
              // parity wait on the weight barrier below cannot alias an older phase (this warp skips the attention
              // phases and may be far ahead of the ring)
              mk_wait(bar_xrdy + 8 * r.s, (r.xmask >> r.s) & 1u);
              r.xmask ^= 1u << r.s;
              if (kb == 0) MK_STAMP(fm, 1);
              if (kb == nkb - 1) MK_STAMP(fm, 4);
              mk_wait(bar_full + 8 * r.s, r.ph);
              if (kb == 0) MK_STAMP(fm, 2);
              if (kb == nkb - 1) MK_STAMP(fm, 5);
              tc_fence_after();
              const uint32_t sa = ring_a + (uint32_t)r.s * MK_STAGE;
              const uint64_t da_hi = make_sw128_kmajor_desc(sa), da_lo = make_sw128_kmajor_desc(sa + 16384);
              const uint64_t db_hi = make_sw128_kmajor_desc(sa + MK_XOFF);
              const uint64_t db_lo = make_sw128_kmajor_desc(sa + MK_XOFF + MK_XPLANE);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da_lo + 2 * k, db_hi + 2 * k, idesc, k == 0 ? acc : 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_lo + 2 * k, idesc, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da_hi + 2 * k, db_hi + 2 * k, idesc, 1);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty + 8 * r.s) : "memory");
              }
              __syncwarp();
              acc = 1;
              r.adv();
            }
            if (elect_one())
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull) : "memory");
            __syncwarp();
            MK_STAMP(fm, 3);
            ++n_item;
          }
        }
      }
    }
#ifdef MK_PF_LANE
  } else if (warp == 10) {
    // =============================================================================================== L2 prefetch lane
    // Next-round experiment: the paced prefetch of decision 9 lives in the producer's slot wait and so
    // realises only ~11 MB per layer.  This lane follows the consumers' phase counter instead: from the moment they
    // enter layer l (phase 0) until they enter its cross phase it issues p.l2pf bytes of the CTA's coming cross range
    // (consumption order, like the producer's cursor) as p.l2pf_piece-byte cp.async.bulk.prefetch.L2, one per
    // p.l2pf_gap cycles, i.e. an even, volume-controlled stream under the qkv / self / o phases.  Build this variant
    // next to the default library and A/B it with MG_B200_LIB (tools/ab_env.py); set MG_MEGA_L2PF_MASK=0 to switch the
    // producer's own prefetch off for a clean comparison.
    if (lane == 0 && p.l2pf > 0 && p.l2pf_piece > 0) {
      const uint32_t pf_hw = (uint32_t)Mp * 192u;
      const uint32_t pf_total = min((uint32_t)p.l2pf, (uint32_t)(xh_hi - xh_lo) * pf_hw);
      for (int l = 0; l < NL && pf_total; ++l) {
        const uint8_t* ckv = s_layers[l].ckv;
        const int ph_begin = l * MK_NPH, ph_cross = l * MK_NPH + MK_PH_CROSS;
        long long t0 = clock64();
        while (*s_phase < ph_begin) {
          __nanosleep(200);
          if (clock64() - t0 > 4000000000LL) mk_die(6, (uint32_t)ph_begin, (uint32_t)*s_phase);
        }
        uint32_t cur = 0;
        long long last = clock64() - p.l2pf_gap;
        while (cur < pf_total && *s_phase < ph_cross) {
          const long long now = clock64();
          if (now - last < (long long)p.l2pf_gap) continue;
          last = now;
          const uint32_t j = cur / pf_hw, off = cur - j * pf_hw;
          int hw;
          if (x_lone_k && j == 0) hw = xh_hi - 1;
          else if ((int)j - x_lone_k < 2 * x_whole) hw = 2 * x_first + (int)j - x_lone_k;
          else hw = xh_lo;
          const uint32_t n = min(min((uint32_t)p.l2pf_piece, pf_total - cur), pf_hw - off);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(ckv + (size_t)hw * pf_hw + off)), "r"(n) : "memory");
          cur += n;
        }
      }
    }
#endif
  } else {
    // =============================================================================================== consumers
    const int ct = threadIdx.x - 64;  // 0..255
    const int cw = warp - 2;          // 0..7
    const bool is_worker = cw < 4;    // warps 2..5: activation staging + TMEM epilogue; warps 6..9: row statistics
    const int r16 = ct >> 4, c16 = ct & 15;
    RingPos r;
    uint32_t n_item = 0;
    unsigned bar_target = 0;
    unsigned* const bar_ctr = p.bar_ctr + (step & 1);
    int phase_i = 0;

    for (int l = 0; l <= NL; ++l) {
      const MegaLayer& L = s_layers[min(l, NL - 1)];
      const int nph = l < NL ? MK_NPH : 1;
#ifdef MK_FOLD_FF
      float* const xin = (l & 1) ? p.xalt : p.x;   // residual stream read by layer l (the LM head reads X[NL & 1])
      float* const xout = (l & 1) ? p.x : p.xalt;  // written by its wo phase
#else
      float* const xin = p.x;
#endif
      for (int ph = 0; ph < nph; ++ph) {
        if (l < NL && ph == MK_PH_SELF) {
          // ---------------------------------------------------------------------------------- self-attention
          // fused KV-cache append + single-query attention with the T5 unidirectional bucket bias (no 1/sqrt(d)).
          // K cache [b][h][key/32][64 d][32 keys] (contiguous per (image, head), conflict-free thread = key),
          // V cache [b][h][key][64 d].  An item is small and its chain of dependent steps long, so the CTA works on
          // MK_SELF_NG items at once: one pair of warps per item, synchronised by its own named barrier; the
          // producer interleaves the items' chunks in the ring (chunk c of item gi sits c * ng + gi loads ahead).
          const int gi = cw >> 1, t = ct & 63, w2 = cw & 1;
          float* const sc_g = s_sc + gi * 512;                      // probabilities of this group's item (Tp <= 512)
          float* const q_g = s_red + gi * 192;                      // q [64] | new v [64] | scratch [64]
          float* const vn_g = q_g + 64;
          float* const red_g = q_g + 128;
          const int nch = self_nkc + self_nvc;
          auto gsync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(4 + gi) : "memory"); };
          MK_XP(long long xs_wait = 0, xs_kmath = 0, xs_vmath = 0, xs_head = 0, xs_soft = 0, xs_tail = 0, xs_sync = 0;
                const long long xs_begin = clock64(); long long xs_t = xs_begin;)
          for (int k0 = 0; k0 < my_attn; k0 += MK_SELF_NG) {
            const int ng = min(MK_SELF_NG, my_attn - k0);
            if (gi < ng) {
              const int it = attn_first + (k0 + gi) * Ga;
              const int b = it / H, h = it - b * H;
              {
                const float* qp = p.qkv + (int64_t)b * 3 * D + h * 64 + t;
                const float rsb = __ldcg(p.rs + b);  // RMSNorm row scale deferred from the qkv projection
                const float qv = __ldcg(qp) * rsb, kn = __ldcg(qp + D) * rsb, vn = __ldcg(qp + 2 * D) * rsb;
                q_g[t] = qv;
                vn_g[t] = vn;
                red_g[t] = qv * kn;
                L.skb[(size_t)it * Tb * 2048 + (size_t)(step >> 5) * 2048 + t * 32 + (step & 31)] = kn;  // append
                L.svb[((size_t)it * p.Tp + step) * 64 + t] = vn;
              }
              float bias8[8];  // T5 bucket bias of this thread's keys j = t + 64 i (two dependent global loads each)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = t + 64 * i;
                bias8[i] = j <= step ? p.dec_bias[p.lut[step - j] * H + h] : 0.f;
              }
              gsync();
              MK_XP(xs_head += clock64() - xs_t;)
              // scores over the cached keys stay in registers: the thread that scores key j also owns it below
              float s8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) s8[i] = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c < self_nkc) {
                  const int pos = r.s + c * ng + gi;
                  const int st = pos % MK_NST;
                  const uint32_t par = r.ph ^ (uint32_t)((pos / MK_NST) & 1);
                  // The groups drift apart and bulk copies land out of order, so this stage's PREVIOUS load (another
                  // group's) may still be in flight, and a parity wait two phases ahead would alias.  The producer
                  // issues a load only after the stage's previous occupant has landed and been consumed, so first
                  // wait until this load has been issued, then for it to land.
                  MK_XP(const long long w0 = clock64();)
                  {
                    const int seq = r.n + c * ng + gi;
                    uint32_t spins = 0;
                    while (MK_FLAG_LD(s_issued) <= seq)
                      if (++spins > (1u << 26)) mk_die(4, (uint32_t)seq, (uint32_t)*s_issued);
                  }
                  mk_wait(bar_full + 8 * st, par);
                  MK_XP(const long long w1 = clock64(); xs_wait += w1 - w0;)
                  const float* buf = reinterpret_cast<const float*>(ring + (size_t)st * MK_STAGE);
                  const int nb = min(MK_SELF_KB, self_nblk - c * MK_SELF_KB);
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const int blk = w2 + 2 * u;
                    if (blk < nb) {
                      const float* kp = buf + blk * 2048 + lane;
                      float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
                      for (int d = 0; d < 64; d += 2) {
                        a0 += q_g[d] * kp[d * 32];
                        a1 += q_g[d + 1] * kp[d * 32 + 32];
                      }
                      s8[c * 2 + u] = a0 + a1;
                    }
                  }
                  MK_XP(const long long w2 = clock64(); xs_kmath += w2 - w1;)
                  gsync();
                  if (t == 0) mk_arrive(bar_empty + 8 * st);
                  MK_XP(xs_sync += clock64() - w2;)
                }
              }
              MK_XP(xs_t = clock64();)
              float snew = 0.f;  // score of the token being appended
#pragma unroll 8
              for (int d = 0; d < 64; ++d) snew += red_g[d];
              float mx = -INFINITY;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = t + 64 * i;
                s8[i] = j < step ? s8[i] + bias8[i] : (j == step ? snew + bias8[i] : -INFINITY);
                mx = fmaxf(mx, s8[i]);
              }
              mx = warp_max(mx);
              if (lane == 0) s_b[gi * 2 + w2] = mx;
              gsync();
              mx = fmaxf(s_b[gi * 2], s_b[gi * 2 + 1]);
              float sum = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = t + 64 * i;
                if (j <= step) {
                  const float e = expf(s8[i] - mx);
                  sc_g[j] = e;
                  sum += e;
                }
              }
              sum = warp_sum(sum);
              gsync();  // both warps have read the maxima
              if (lane == 0) s_b[gi * 2 + w2] = sum;
              gsync();  // probabilities and partial sums visible
              sum = s_b[gi * 2] + s_b[gi * 2 + 1];
              // P.V: thread (r4 = t / 16, c = t % 16) -> float4 column c over keys j == r4 (mod 4)
              const int r4 = t >> 4, cc = t & 15;
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              MK_XP(xs_soft += clock64() - xs_t;)
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c < self_nvc) {
                  const int pos = r.s + (self_nkc + c) * ng + gi;
                  const int st = pos % MK_NST;
                  const uint32_t par = r.ph ^ (uint32_t)((pos / MK_NST) & 1);
                  MK_XP(const long long w0 = clock64();)
                  {
                    const int seq = r.n + (self_nkc + c) * ng + gi;
                    uint32_t spins = 0;
                    while (MK_FLAG_LD(s_issued) <= seq)
                      if (++spins > (1u << 26)) mk_die(4, (uint32_t)seq, (uint32_t)*s_issued);
                  }
                  mk_wait(bar_full + 8 * st, par);
                  MK_XP(const long long w1 = clock64(); xs_wait += w1 - w0;)
                  const float4* buf4 = reinterpret_cast<const float4*>(ring + (size_t)st * MK_STAGE);
                  const int m0 = c * MK_SELF_VR, rows = min(MK_SELF_VR, step - m0);
#pragma unroll 4
                  for (int jj = r4; jj < rows; jj += 4) {
                    const float4 vv = buf4[jj * 16 + cc];
                    const float pj = sc_g[m0 + jj];
                    acc.x += pj * vv.x; acc.y += pj * vv.y; acc.z += pj * vv.z; acc.w += pj * vv.w;
                  }
                  MK_XP(const long long w2 = clock64(); xs_vmath += w2 - w1;)
                  gsync();
                  if (t == 0) mk_arrive(bar_empty + 8 * st);
                  MK_XP(xs_sync += clock64() - w2;)
                }
              }
              MK_XP(xs_t = clock64();)
              // combine the four key residues: lanes l and l+16 inside a warp, then the two warps through shared memory
              acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
              acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
              if (w2 == 1 && lane < 16) reinterpret_cast<float4*>(red_g)[lane] = acc;
              gsync();
              if (w2 == 0 && lane < 16) {
                const float4 o2 = reinterpret_cast<const float4*>(red_g)[lane];
                const float4 vn4 = reinterpret_cast<const float4*>(vn_g)[lane];
                const float pn = sc_g[step], inv = 1.f / sum;
                float4 o;
                o.x = (acc.x + o2.x + pn * vn4.x) * inv; o.y = (acc.y + o2.y + pn * vn4.y) * inv;
                o.z = (acc.z + o2.z + pn * vn4.z) * inv; o.w = (acc.w + o2.w + pn * vn4.w) * inv;
                *reinterpret_cast<float4*>(p.ctx + (int64_t)b * D + h * 64 + 4 * lane) = o;
              }
              gsync();  // this group's shared scratch is reusable
              MK_XP(xs_tail += clock64() - xs_t; xs_t = clock64();)
            }
            r.adv_n(nch * ng);
          }
          MK_XP(if (p.prof && gi == 0 && t == 0 && l == NL / 2) {
            unsigned long long* q = p.prof + ((size_t)g * 512 + 464) * 2;
            q[0] = xs_wait; q[1] = xs_kmath; q[2] = xs_vmath; q[3] = xs_sync; q[4] = xs_head; q[5] = xs_soft; q[6] = xs_tail;
            q[7] = clock64() - xs_begin;
          })
        } else if (l < NL && ph == MK_PH_CROSS) {
          // ---------------------------------------------------------------------------------- cross-attention
          // kv24 K^T / V blocks (decode.cu: 16-bit + 8-bit planes, 3 bytes per element); additive mask
          // (1-mask)*finfo.min, no positional bias, no scale.  Scores: thread = quad of adjacent keys.
          const unsigned x_epoch = (unsigned)(step * NL + l + 1);  // unique per (step, layer): the flags need no reset
          // A chunk's stage is handed back after a 256-thread barrier.  (-DMK_WARPREL, measured experiment: the LAST of the
          // eight consumer warps to finish a chunk hands it back through a shared-memory counter, so the warps may drift
          // apart inside a pass -- 2.132 instead of 2.103 ms per step: the counter's atomics and the lost lockstep cost
          // more than the barrier.)
          auto cross_release = [&]() {
#ifdef MK_WARPREL
            __syncwarp();
            if (lane == 0 && atomicAdd(s_rel + r.s, 1) == 7) {
              s_rel[r.s] = 0;  // the stage cannot be refilled and finished by eight warps again before this arrive
              mk_arrive(bar_empty + 8 * r.s);
            }
#else
            cons_sync();
            if (ct == 0) mk_arrive(bar_empty + 8 * r.s);
#endif
          };
          int rs_b = -1;    // image whose row scale rsb holds: consecutive entries are mostly heads of one image
          float rsb = 0.f;
          MK_XP(long long xc_wait = 0, xc_math = 0, xc_sync = 0, xc_head = 0, xc_soft = 0, xc_tail = 0, xc_n = 0;
                const long long xc_begin = clock64(); long long xc_t = xc_begin;)
          for (int e = 0; e < x_entries; ++e) {
            int it, passes;
            cross_entry(e, it, passes);
            MK_XP(xc_t = clock64();)
            const int b = it / H, h = it - b * H;
            float* const xp = p.xp + (size_t)it * (Mp + 4);  // published probabilities of a split item
            float* const s_r2 = s_part + (e & 1) * 32;       // [0, 16) reduction pairs, [16, 24) partial sums; alternates per entry
            if (passes & 1) {
              // ---- K pass: scores of this thread's key pairs, mask, softmax numerators into s_sc
              // q is the UNSCALED cross query (accumulated by phases 0 and 2); its RMSNorm row scale
              // rsqrt(mean(x[b]^2) + eps) -- x is complete only since the barrier before this phase -- multiplies the
              // scores after the K stream: each thread fetches 4 of the D <= 1024 residual values now, the block
              // reduces them once the K pass is over, so the loads' latency hides under the stream.
              if (ct < 64) s_q[ct] = __ldcg(p.q + (int64_t)b * D + h * 64 + ct);
              float4 xr = make_float4(0.f, 0.f, 0.f, 0.f);
              const bool new_b = b != rs_b;  // CTA-uniform
              if (new_b && 4 * ct < D) xr = ldcg4(xin + (int64_t)b * D + 4 * ct);
              int mk8[8];  // this thread's mask bits, fetched now so their latency hides under the K pass
#pragma unroll
              for (int i = 0; i < 8; ++i) mk8[i] = p.mem_mask[(int64_t)b * Mp + min(4 * ct + (i & 3) + 1024 * (i >> 2), Mp - 1)];
              const bool kq0 = 128 * cw < Mp, kq1 = 1024 + 128 * cw < Mp;  // this warp's two key quads-of-32 exist (warp-uniform)
              cons_sync();
              float acc[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] = 0.f;
              MK_XP(xc_head += clock64() - xc_t;)
              for (int c = 0; c < cross_nkc; ++c) {
                MK_XP(const long long w0 = clock64();)
                mk_wait(bar_full + 8 * r.s, r.ph);
                MK_XP(const long long w1 = clock64(); xc_wait += w1 - w0; ++xc_n;)
                const uint8_t* buf = ring + (size_t)r.s * MK_STAGE;
                const int r0 = c * cross_rk, rows = min(cross_rk, 64 - r0);
                const int rows_math = (p.dbg & 8) ? 0 : rows;  // MG_MEGA_DBG bit 3: stream only, no K-pass arithmetic (wrong ids; timing A/B)
                // thread = 4 adjacent keys (+ the 4 keys 1024 further on): one 8-byte load of the 16-bit plane and one
                // 4-byte load of the 8-bit plane per quad -- the pass is bound by shared-memory wavefronts (the bulk
                // copies' writes included), and this layout needs 3 per 128 keys and d-row where key PAIRS with 2-byte
                // loads of the 8-bit plane needed 4 and ran every warp over all 2048 key slots.  Warps whose keys lie
                // beyond Mp skip the loads (warp-uniform); threads past the row's end inside an active warp read the
                // next row / plane (inside this CTA's shared memory), their sums are never used.
                const uint2* hrow = reinterpret_cast<const uint2*>(buf) + ct;
                const uint32_t* lrow = reinterpret_cast<const uint32_t*>(buf + (size_t)rows * Mp * 2) + ct;
                const int pitch = Mp >> 2;
                // (the loop exists once per warp-uniform case so that its body is branch-free: loads first, then arithmetic)
                auto k_rows = [&](auto both) {
MK_PRAGMA(unroll MK_XUNROLL)
                  for (int rr = 0; rr < rows_math; ++rr) {
                    const float qd = s_q[r0 + rr];
                    const uint2 h0 = hrow[0];
                    const uint32_t l0 = lrow[0];
                    uint2 h1 = make_uint2(0u, 0u);
                    uint32_t l1 = 0u;
                    if constexpr (decltype(both)::value) { h1 = hrow[256]; l1 = lrow[256]; }
                    hrow += pitch;
                    lrow += pitch;
                    const uint32_t la = __byte_perm(l0, 0u, 0x4240), lb = __byte_perm(l0, 0u, 0x4341);
                    acc[0] += qd * __uint_as_float(__byte_perm(h0.x, la, 0x1045));
                    acc[1] += qd * __uint_as_float(__byte_perm(h0.x, lb, 0x3245));
                    acc[2] += qd * __uint_as_float(__byte_perm(h0.y, la, 0x1065));
                    acc[3] += qd * __uint_as_float(__byte_perm(h0.y, lb, 0x3265));
                    if constexpr (decltype(both)::value) {
                      const uint32_t lc = __byte_perm(l1, 0u, 0x4240), ld = __byte_perm(l1, 0u, 0x4341);
                      acc[4] += qd * __uint_as_float(__byte_perm(h1.x, lc, 0x1045));
                      acc[5] += qd * __uint_as_float(__byte_perm(h1.x, ld, 0x3245));
                      acc[6] += qd * __uint_as_float(__byte_perm(h1.y, lc, 0x1065));
                      acc[7] += qd * __uint_as_float(__byte_perm(h1.y, ld, 0x3265));
                    }
                  }
                };
                if (kq1) k_rows(std::true_type{});
                else if (kq0) k_rows(std::false_type{});
                MK_XP(const long long w2 = clock64(); xc_math += w2 - w1;)
                cross_release();
                r.adv();
                MK_XP(xc_sync += clock64() - w2;)
              }
              MK_XP(xc_t = clock64();)
              // ONE reduction round for both block-wide values: the sum of squares of the residual row (new image only)
              // and the largest unmasked raw score.  rsb > 0 and rounding is monotonic, so max_i fl(acc_i * rsb) =
              // fl(max_i acc_i * rsb); a masked key's value acc * rsb + finfo.min is exactly finfo.min.  The softmax
              // denominator is only needed after the V pass: the warps' partial sums wait in shared memory until then.
              float ssq = (xr.x * xr.x + xr.y * xr.y) + (xr.z * xr.z + xr.w * xr.w);
              float mraw = -INFINITY;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = 4 * ct + (i & 3) + 1024 * (i >> 2);
                if (m < Mp && mk8[i]) mraw = fmaxf(mraw, acc[i]);
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
                mraw = fmaxf(mraw, __shfl_xor_sync(0xffffffffu, mraw, o));
              }
              if (lane == 0) { s_r2[cw] = ssq; s_r2[8 + cw] = mraw; }
              cons_sync();
              if (new_b) {
                float t = s_r2[0];
#pragma unroll
                for (int w = 1; w < 8; ++w) t += s_r2[w];
                rsb = rsqrtf(t / (float)D + p.eps);
                rs_b = b;
              }
              float mx = s_r2[8];
#pragma unroll
              for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_r2[8 + w]);
              // every key masked: all values equal finfo.min (uniform attention, like the reference's additive mask)
              mx = (mx == -INFINITY) ? -3.4028234663852886e38f : mx * rsb;
              float psum = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = 4 * ct + (i & 3) + 1024 * (i >> 2);
                if (m < Mp) {
                  const float ev = expf((acc[i] * rsb + (mk8[i] ? 0.f : -3.4028234663852886e38f)) - mx);
                  s_sc[m] = ev;
                  if (passes == 1) xp[m] = ev;  // the V pass of this item runs on the next CTA
                  psum += ev;
                }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
              if (lane == 0) s_r2[16 + cw] = psum;
              cons_sync();  // probabilities and the warps' partial sums visible
              if (passes == 1) {
                if (ct == 0) {
                  float t = s_r2[16];
#pragma unroll
                  for (int w = 1; w < 8; ++w) t += s_r2[16 + w];
                  xp[Mp] = t;
                  __threadfence();  // cumulative: the other threads' stores were ordered before it by the barrier above
                  atomicExch(p.xflag + it, x_epoch);
                }
                continue;
              }
            } else {
              // ---- V pass of an item whose K pass ran on the previous CTA (as ITS first work: long done)
              if (ct == 0) {
                unsigned v;
                const long long t0 = clock64();
                for (;;) {
                  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.xflag + it) : "memory");
                  if (v == x_epoch) break;
                  if (clock64() - t0 > 4000000000LL) mk_die(5, v, x_epoch);
                }
              }
              cons_sync();
              for (int m = ct; m < Mp; m += 256) s_sc[m] = __ldcg(xp + m);
              if (ct < 8) s_r2[16 + ct] = ct == 0 ? __ldcg(xp + Mp) : 0.f;
              cons_sync();
            }
            // ---- V pass: thread = 8 adjacent d-columns of the keys  jj == ct / 8 (mod 32)  -- one 16-byte load of the
            // 16-bit plane, one 8-byte load of the 8-bit plane and one probability per key row and thread
            float av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = 0.f;
            const int r32 = ct >> 3, c8 = ct & 7;
            MK_XP(xc_soft += clock64() - xc_t;)
            for (int c = 0; c < cross_nvc; ++c) {
              MK_XP(const long long w0 = clock64();)
              mk_wait(bar_full + 8 * r.s, r.ph);
              MK_XP(const long long w1 = clock64(); xc_wait += w1 - w0; ++xc_n;)
              const uint8_t* buf = ring + (size_t)r.s * MK_STAGE;
              const int m0 = c * cross_vr, rows = min(cross_vr, Mp - m0);
              const uint8_t* hp = buf + c8 * 16;
              const uint8_t* lp = buf + (size_t)rows * 128 + c8 * 8;
MK_PRAGMA(unroll MK_XUNROLL)
              for (int jj = (p.dbg & 16) ? rows : r32; jj < rows; jj += 32) {  // MG_MEGA_DBG bit 4: stream only, no V-pass arithmetic
                const uint4 hh = *reinterpret_cast<const uint4*>(hp + (size_t)jj * 128);
                const uint2 ll = *reinterpret_cast<const uint2*>(lp + (size_t)jj * 64);
                const float pj = s_sc[m0 + jj];
                const uint32_t la = __byte_perm(ll.x, 0u, 0x4240), lb = __byte_perm(ll.x, 0u, 0x4341);
                const uint32_t lc = __byte_perm(ll.y, 0u, 0x4240), ld = __byte_perm(ll.y, 0u, 0x4341);
                av[0] += pj * __uint_as_float(__byte_perm(hh.x, la, 0x1045));
                av[1] += pj * __uint_as_float(__byte_perm(hh.x, lb, 0x3245));
                av[2] += pj * __uint_as_float(__byte_perm(hh.y, la, 0x1065));
                av[3] += pj * __uint_as_float(__byte_perm(hh.y, lb, 0x3265));
                av[4] += pj * __uint_as_float(__byte_perm(hh.z, lc, 0x1045));
                av[5] += pj * __uint_as_float(__byte_perm(hh.z, ld, 0x3245));
                av[6] += pj * __uint_as_float(__byte_perm(hh.w, lc, 0x1065));
                av[7] += pj * __uint_as_float(__byte_perm(hh.w, ld, 0x3265));
              }
              MK_XP(const long long w2 = clock64(); xc_math += w2 - w1;)
              cross_release();
              r.adv();
              MK_XP(xc_sync += clock64() - w2;)
            }
            MK_XP(xc_t = clock64();)
            // the four key residues of a warp through shuffles, the eight warps through shared memory
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              av[i] += __shfl_xor_sync(0xffffffffu, av[i], 8);
              av[i] += __shfl_xor_sync(0xffffffffu, av[i], 16);
            }
            if (lane < 8) {
              float4* dst = reinterpret_cast<float4*>(s_red + cw * 64 + lane * 8);
              dst[0] = make_float4(av[0], av[1], av[2], av[3]);
              dst[1] = make_float4(av[4], av[5], av[6], av[7]);
            }
            cons_sync();
            if (ct < 64) {
              float o = 0.f;
#pragma unroll
              for (int w = 0; w < 8; ++w) o += s_red[w * 64 + ct];
              float t = s_r2[16];
#pragma unroll
              for (int w = 1; w < 8; ++w) t += s_r2[16 + w];
              p.ctx[(int64_t)b * D + h * 64 + ct] = o / t;
            }
            // no trailing barrier: s_red is next written after the next entry's V pass (barriers in between), s_q and
            // s_sc after its first barrier, and the reduction scratch alternates per entry
            MK_XP(xc_tail += clock64() - xc_t;)
          }
          MK_XP(if (p.prof && ct == 0 && l == NL / 2) {
            unsigned long long* q = p.prof + ((size_t)g * 512 + 448) * 2;
            q[0] = xc_wait; q[1] = xc_math; q[2] = xc_sync; q[3] = xc_head; q[4] = xc_soft; q[5] = xc_tail; q[6] = xc_n;
            q[7] = clock64() - xc_begin;
          })
        } else {
          // ---------------------------------------------------------------------------------- linear
          // out[b][n] (+)= sum_k pro(x)[b][k] * W[n][k];  pro: 0 none, 1 RMSNorm weight (x*lnw staged; the row scale
          // rs[b] = rsqrt(mean(x^2)+eps) is NOT applied here: it is computed off the critical path by one warp per
          // image, published through p.rs, and applied by the consumer of this phase's output -- q/k/v when they are
          // loaded, the FF hidden activations as relu(rs*h) = rs*relu(h); the cross query's row scale is computed by
          // the cross-attention phase itself, under its K stream), 2 ReLU * rs.
          // LM head only: rs * d_model^-0.5 applied in the epilogue (the logits are the final output).
          // (input, prologue, output, which buffer this phase zeroes for a later one):
          int pro = 1, ldx = D, ld_out = D, ld_out2 = D, rs_slot = -1;
          const float* x = xin;
          float* out = xin;
          float* out2 = p.q;  // rows >= W.n_split of a two-output linear (phases 0 and 2) go to the cross query
          const float* lnw = nullptr;
          float* zero_ptr = nullptr;
          int64_t zero_n = 0;
          const bool store = l == NL;
          bool wo_fold = false;  // MK_FOLD_FF wo phase: epilogue scales by rs[b] and k-slice 0 adds the residual x1 + dx
          if (store) {  // LM head: final RMSNorm * d_model^-0.5 fused, direct store + per-tile argmax
            lnw = p.final_ln; out = p.logits; ld_out = p.ld_logits;
          } else if (ph == 0) {  // raw x -> qkv | q (ln1 / ln2 are folded into the weight rows); zero: FF hidden buffer
            pro = 0; out = p.qkv; ld_out = 3 * D; zero_ptr = p.hbuf; zero_n = (int64_t)B * p.DFF; rs_slot = 0;
          } else if (ph == 2) {  // self-attention ctx -> x (+=) | q (+=); zero: qkv (consumed by phase 1) [+ dx behind it]
            pro = 0; x = p.ctx; zero_ptr = p.qkv;
#ifdef MK_FOLD_FF
            zero_n = (int64_t)B * 4 * D;
#else
            zero_n = (int64_t)B * 3 * D;
#endif
          }
#ifdef MK_FOLD_FF
          else if (ph == 4) {  // ctx -> dx | hidden, raw x1 -> hidden (ln3 folded into the columns); zero: the next residual
            pro = 0; x = p.ctx; out = p.dx; out2 = p.hbuf; ld_out2 = p.DFF; zero_ptr = xout; zero_n = (int64_t)B * D;
          } else {  // ph == 5: relu(hidden) -> x3 = (x1 + dx) + rs * (Wwo relu(h)); zero: q (consumed by phase 3)
            pro = 2; x = p.hbuf; ldx = p.DFF; out = xout; wo_fold = true; zero_ptr = p.q; zero_n = (int64_t)B * D;
          }
#else
          else if (ph == 4) {  // cross-attention ctx -> x (+=)
            pro = 0; x = p.ctx;
          } else if (ph == 5) {  // x -> hidden (RMSNorm ln3); zero: q (consumed by phase 3)
            lnw = L.ln[2]; out = p.hbuf; ld_out = p.DFF; zero_ptr = p.q; zero_n = (int64_t)B * D; rs_slot = 2;
          } else {  // ph == 6: rs * relu(hidden) -> x (+=)
            pro = 2; x = p.hbuf; ldx = p.DFF;
          }
#endif
          const bool all_rs = store || wo_fold;  // every CTA needs the scale of all rows in its epilogue
          const MegaLin& W = l < NL ? L.lin[mega_lin_of_phase(ph)] : p.lm_head;
          const int items = W.tiles * W.ksplit;
#ifdef MK_FINE
          unsigned long long* fine = (p.prof && ct == 0 && phase_i >= MK_NPH && phase_i < 2 * MK_NPH) ? p.prof + ((size_t)g * 512 + 256 + (phase_i - MK_NPH) * 16) * 2 : nullptr;
#endif
          MK_STAMP(fine, 0);
          if (!is_worker) {
            const int t = ct - 128, wq = cw - 4;
            if (all_rs) {
              // LM head / folded wo: every CTA needs the scale of all 32 rows in its epilogue (warp wq owns rows
              // 8 wq .. 8 wq + 7); the folded wo normalises x1 + dx (the residual after the cross-attention block)
              if (g < items) {
                const int n4 = D >> 2;
                const float post = store ? p.logit_scale : 1.f;
#pragma unroll 1
                for (int rp = 0; rp < 8; rp += 2) {
                  float4 qa[2][8];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const int64_t ro = (int64_t)min(wq * 8 + rp + u, B - 1) * D;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const int c = 4 * min(lane + 32 * i, n4 - 1);
                      qa[u][i] = ldcg4(xin + ro + c);
#ifdef MK_FOLD_FF
                      if (wo_fold) {
                        const float4 e = ldcg4(p.dx + ro + c);
                        qa[u][i].x += e.x; qa[u][i].y += e.y; qa[u][i].z += e.z; qa[u][i].w += e.w;
                      }
#endif
                    }
                  }
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    float ss = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const float4 v = qa[u][i];
                      ss += (lane + 32 * i < n4) ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f;
                    }
                    ss = warp_sum(ss);
                    if (lane == 0) s_rs[wq * 8 + rp + u] = rsqrtf(ss / (float)D + p.eps) * post;
                  }
                }
              }
            }
            if (!store) {
              if (rs_slot >= 0 && wq == 0 && g < B) {
                // RMSNorm row scale of image g, consumed one phase later (by other CTAs) through p.rs
                const int n4 = W.K >> 2;
                const float* xr = x + (int64_t)g * ldx;
                float4 qa[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) qa[i] = ldcg4(xr + 4 * min(lane + 32 * i, n4 - 1));
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  ss += (lane + 32 * i < n4) ? (qa[i].x * qa[i].x + qa[i].y * qa[i].y) + (qa[i].z * qa[i].z + qa[i].w * qa[i].w) : 0.f;
                ss = warp_sum(ss);
                if (lane == 0) p.rs[rs_slot * MK_R + g] = rsqrtf(ss / (float)W.K + p.eps);
              }
              if (zero_ptr) {  // zero duty while the workers stage
                float4* z4 = reinterpret_cast<float4*>(zero_ptr);
                for (int64_t i = (int64_t)g * 128 + t; i < (zero_n >> 2); i += (int64_t)G * 128) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
          const int c4 = ct & 15, r8 = (ct & 127) >> 4;
          const float relu_lo = (pro == 2) ? 0.f : -INFINITY;
          bool first = true;
          for (int it = g; it < items; it += G) {
            int tile, kb0;
            const int nkb = lin_item_kbs(W, it, tile, kb0);
            if (is_worker) {
              // ---- stage the activation tiles of this item's k-blocks (128 threads)
#ifdef MK_FOLD_FF
              // three-segment linear (phase 4): the rows >= n_split2 (Wi') multiply the raw residual, the others ctx
              const float* xk = ((W.n_split2 && tile * 128 >= W.n_split2) ? xin : x) + kb0 * 64 + c4 * 4;
#else
              const float* xk = x + kb0 * 64 + c4 * 4;
#endif
              float4 v[4], gw = make_float4(1.f, 1.f, 1.f, 1.f);
              if (pro == 1) gw = *reinterpret_cast<const float4*>(lnw + kb0 * 64 + c4 * 4);
              float rsr[4] = {1.f, 1.f, 1.f, 1.f};
#ifndef MK_FOLD_FF
              if (pro == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rsr[i] = __ldcg(p.rs + 2 * MK_R + min(r8 + i * 8, B - 1));
              }
#endif
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = ldcg4(xk + (int64_t)min(r8 + i * 8, B - 1) * ldx);
#ifdef MK_XPF2
              // experiment build: the activation tiles travel TWO k-blocks ahead of the staging (one more L2 round trip hidden)
              float4 v1[4], g1 = make_float4(1.f, 1.f, 1.f, 1.f);
              {
                const int k1 = min(1, nkb - 1) * 64;
                if (pro == 1) g1 = *reinterpret_cast<const float4*>(lnw + kb0 * 64 + k1 + c4 * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) v1[i] = ldcg4(xk + k1 + (int64_t)min(r8 + i * 8, B - 1) * ldx);
              }
#endif
#pragma unroll 1
              for (int kb = 0; kb < nkb; ++kb) {
                float4 vn[4], gn = make_float4(1.f, 1.f, 1.f, 1.f);
#ifdef MK_XPF2
                const int kn = min(kb + 2, nkb - 1) * 64;
#else
                const int kn = min(kb + 1, nkb - 1) * 64;  // prefetch the next k-block while this one is staged
#endif
                if (pro == 1) gn = *reinterpret_cast<const float4*>(lnw + kb0 * 64 + kn + c4 * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) vn[i] = ldcg4(xk + kn + (int64_t)min(r8 + i * 8, B - 1) * ldx);
                mk_wait(bar_empty + 8 * r.s, r.ph ^ 1);
                if (kb == 0 && v[3].w != 1.2345e-30f) MK_STAMP(fine, 1);
                uint8_t* xs_hi = ring + (size_t)r.s * MK_STAGE + MK_XOFF;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = r8 + i * 8;
                  float4 w = v[i];
                  if (rr >= B) w = make_float4(0.f, 0.f, 0.f, 0.f);
                  const float gx = gw.x * rsr[i], gy = gw.y * rsr[i], gz = gw.z * rsr[i], gq = gw.w * rsr[i];
                  w.x = fmaxf(w.x, relu_lo) * gx; w.y = fmaxf(w.y, relu_lo) * gy;
                  w.z = fmaxf(w.z, relu_lo) * gz; w.w = fmaxf(w.w, relu_lo) * gq;
                  // fp32 -> bf16 hi + bf16 lo (x ~= hi + lo), two elements per conversion
                  const __nv_bfloat162 h01 = __floats2bfloat162_rn(w.x, w.y), h23 = __floats2bfloat162_rn(w.z, w.w);
                  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                  const __nv_bfloat162 l01 = __floats2bfloat162_rn(w.x - f01.x, w.y - f01.y);
                  const __nv_bfloat162 l23 = __floats2bfloat162_rn(w.z - f23.x, w.w - f23.y);
                  // 128B swizzle: 16-byte chunk index XOR (row % 8); this float4 covers half a chunk (8 bytes)
                  const uint32_t off = (uint32_t)rr * 128u + ((((uint32_t)c4 >> 1) ^ ((uint32_t)rr & 7u)) << 4) +
                                       (((uint32_t)c4 & 1u) << 3);
                  uint2 ph2, pl2;
                  ph2.x = *reinterpret_cast<const uint32_t*>(&h01); ph2.y = *reinterpret_cast<const uint32_t*>(&h23);
                  pl2.x = *reinterpret_cast<const uint32_t*>(&l01); pl2.y = *reinterpret_cast<const uint32_t*>(&l23);
                  *reinterpret_cast<uint2*>(xs_hi + off) = ph2;
                  *reinterpret_cast<uint2*>(xs_hi + MK_XPLANE + off) = pl2;
                }
                fence_proxy_async();
                mk_arrive(bar_xrdy + 8 * r.s);
                r.adv();
#ifdef MK_XPF2
#pragma unroll
                for (int i = 0; i < 4; ++i) { v[i] = v1[i]; v1[i] = vn[i]; }
                gw = g1;
                g1 = gn;
#else
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = vn[i];
                gw = gn;
#endif
              }
            } else {
              r.adv_n(nkb);  // keep this warp's view of the ring in step with the workers / the MMA warp
            }
            MK_STAMP(fine, 2);
#ifdef MK_FOLD_FF
            // folded wo: k-slice 0 of every output tile carries the residual x1 + dx; fetch this thread's 16 values now,
            // the loads complete under the MMAs (thread = output feature n, columns = images c_lo .. c_lo + 15)
            const bool add_res = wo_fold && (it % W.ksplit) == 0;
            float res[16];
            if (add_res) {
              const int nn = min(tile * 128 + (warp & 3) * 32 + lane, W.N - 1);
              const int cl = (cw >> 2) * 16;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int64_t o = (int64_t)min(cl + j, B - 1) * D + nn;
                res[j] = __ldcg(xin + o) + __ldcg(p.dx + o);
              }
            }
#endif
            if (all_rs && first) cons_sync();  // LM head / folded wo: row scales from the statistic warps
            first = false;
            // ---- epilogue: TMEM -> registers (thread = output feature, column = image)
            mk_wait(bar_tfull, n_item & 1);
            MK_STAMP(fine, 4);
            tc_fence_after();
            const int q = warp & 3;
            int n = tile * 128 + q * 32 + lane, n_lim = W.N, ldo = ld_out;
            float* o_base = out;
            if (W.n_split) {  // multi-output linear, CTA-uniform branch (the splits are multiples of the tile height)
              if (W.n_split2 && tile * 128 >= W.n_split2) { n -= W.n_split2; n_lim = W.N - W.n_split2; o_base = out2; ldo = ld_out2; }
              else if (tile * 128 >= W.n_split) { n -= W.n_split; n_lim = (W.n_split2 ? W.n_split2 : W.N) - W.n_split; o_base = out2; ldo = ld_out2; }
              else n_lim = W.n_split;
            }
            const bool n_ok = n < n_lim;
            if (!store) {
              // split-K partial sums: red.global.add into the next buffer / the residual stream.  All 8 consumer
              // warps take part: two warps per TMEM lane quadrant, 16 of the 32 image columns each (this loop runs
              // with one warp per scheduler, so its length in instructions is what the phase waits for).
              const int c_lo = (cw >> 2) * 16;
              const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c_lo;
              float* op = o_base + (int64_t)c_lo * ldo + n;
#pragma unroll 1
              for (int c = 0; c < 2; ++c) {
                uint32_t rr[8];
                tmem_ld_32x32_x8(taddr + c * 8, rr);
                tmem_ld_wait();
#ifdef MK_FOLD_FF
                if (wo_fold) {  // CTA-uniform
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    float v = __uint_as_float(rr[j]) * s_rs[c_lo + c * 8 + j];
                    if (add_res) v += (c == 0) ? res[j] : res[8 + j];
                    rr[j] = __float_as_uint(v);
                  }
                }
#endif
                const int nv = B - (c_lo + c * 8);  // image rows left (warp-uniform)
                if (n_ok) {
                  if (nv >= 8) {
                    float* pj = op;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      atomicAdd(pj, __uint_as_float(rr[j]));
                      pj += ldo;
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (j < nv) atomicAdd(op + (int64_t)j * ldo, __uint_as_float(rr[j]));
                  }
                }
                op += (int64_t)8 * ldo;
              }
              tc_fence_before();
              mk_arrive(bar_tempty);
            } else {
              // LM head (once per step, worker warps only): scaled direct store + per-row (max, first argmax)
              if (is_worker) {
                float* o = out + n;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                for (int c0 = 0; c0 < MK_R; c0 += 8) {
                  uint32_t rr[8];
                  tmem_ld_32x32_x8(taddr + c0, rr);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const int b = c0 + j;
                    const float val = __uint_as_float(rr[j]) * s_rs[b];
                    if (n_ok && b < B) o[(int64_t)b * ld_out] = val;
                    // order-preserving integer key + redux.max; the lowest lane holding the max is the first argmax
                    const uint32_t u = __float_as_uint(n_ok ? val : -INFINITY);
                    const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                    const uint32_t mk = __reduce_max_sync(0xffffffffu, key);
                    const uint32_t bal = __ballot_sync(0xffffffffu, key == mk);
                    if (lane == 0) {
                      s_part[b * 4 + q] = __uint_as_float((mk & 0x80000000u) ? (mk & 0x7fffffffu) : ~mk);
                      s_pi[b * 4 + q] = tile * 128 + q * 32 + (__ffs(bal) - 1);
                    }
                  }
                }
              }
              tc_fence_before();
              mk_arrive(bar_tempty);
              if (is_worker) {
                asm volatile("bar.sync 3, 128;" ::: "memory");
                if (ct < B) {
                  float bv = s_part[ct * 4];
                  int bi = s_pi[ct * 4];
#pragma unroll
                  for (int w2 = 1; w2 < 4; ++w2) {
                    const float ov = s_part[ct * 4 + w2];
                    const int oi = s_pi[ct * 4 + w2];
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                  }
                  p.part_val[(int64_t)ct * W.tiles + tile] = bv;
                  p.part_idx[(int64_t)ct * W.tiles + tile] = bi;
                }
                asm volatile("bar.sync 3, 128;" ::: "memory");  // s_part reusable by the next item
              }
            }
            MK_STAMP(fine, 5);
            ++n_item;
          }
          if (all_rs && first) cons_sync();
        }
        if (l == NL) break;
        // ------------------------------------------------------------------------------------ grid barrier
        cons_sync();
        bar_target += (unsigned)G;
#ifdef MK_FLAGBAR
        // experiment build (-DMK_FLAGBAR): no atomics on one word -- every CTA publishes its arrival with a plain store to
        // its own flag, and one warp per CTA polls all flags with five coalesced loads per round
        if (cw == 0) {
          const unsigned epoch = (unsigned)step * (unsigned)(NL * MK_NPH) + (unsigned)phase_i + 1u;
          volatile unsigned* const flags = p.bar_ctr + 4;
          if (lane == 0) {
            __threadfence();
            flags[g] = epoch;
          }
          const long long t0 = clock64();
          for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int idx = lane + 32 * i;
              if (idx < G) ok = ok && (int)(flags[idx] - epoch) >= 0;
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - t0 > 4000000000LL) mk_die(3, epoch, (uint32_t)g);
          }
          __threadfence();
          if (lane == 0) MK_FLAG_ST(s_phase, phase_i + 1);
        }
        if (false) {
#else
        if (ct == 0) {
#endif
#ifdef MK_FINE
          unsigned long long* ps = p.prof ? p.prof + ((size_t)g * 512 + phase_i) * 2 : nullptr;
#endif
          MK_STAMP(ps, 0);
          __threadfence();
          atomicAdd(bar_ctr, 1u);
          unsigned v;
          const long long t0 = clock64();
          for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar_ctr) : "memory");
            if (v >= bar_target) break;
            if (clock64() - t0 > 4000000000LL) mk_die(3, v, bar_target);
          }
          MK_STAMP(ps, 1);
          MK_FLAG_ST(s_phase, phase_i + 1);  // consumers enter the next phase (releases the producer's prefetch gate)
        }
        ++phase_i;
        cons_sync();
      }
    }
#ifdef MK_FINE
    if (p.prof && ct == 0) MK_STAMP(p.prof + ((size_t)g * 512 + phase_i) * 2, 0);
#endif
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

bool mega_fold_ff() {
#ifdef MK_FOLD_FF
  return true;
#else
  return false;
#endif
}

MegaLin make_mega_lin(cudaStream_t st, Planes w, int N, int K, int64_t ldk, bool store, int n_ctas, uint8_t* dst, int n_split,
                      int n_split2) {
  MG_REQUIRE(n_split % 128 == 0 && n_split < N, "fused decode step: output split must be a multiple of the 128-row tile");
  MG_REQUIRE(n_split2 % 128 == 0 && n_split2 < N && (n_split2 == 0 || n_split2 > n_split), "fused decode step: bad second split");
  MegaLin L;
  L.N = N;
  L.n_split = n_split;
  L.n_split2 = n_split2;
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

// ------------------------------------------------------------------------------------------------ folded cross query
// finalize-time, once per layer: gain folding and the Wcq' Wo product of the two-output linears (phases 0 and 2)
__global__ void scale_cols_kernel(const float* __restrict__ src, const float* __restrict__ gain, int64_t total, int K,
                                  float* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * gain[i % K];
}
void launch_scale_cols(cudaStream_t st, const float* src, const float* gain, int rows, int K, float* dst) {
  const int64_t total = (int64_t)rows * K;
  scale_cols_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(src, gain, total, K, dst);
  MG_CHECK_CUDA(cudaGetLastError());
}
// P[n][j] = sum_k A[n][k] * Bm[k][j], row-major, fp64 accumulation (the result is rounded to fp32 once); 32x32 tiles
__global__ void fold_product_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int N, int K, int J,
                                    float* __restrict__ P) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n = blockIdx.y * 32 + ty, j = blockIdx.x * 32 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    sa[ty][tx] = (n < N && k0 + tx < K) ? A[(int64_t)n * K + k0 + tx] : 0.f;
    sb[ty][tx] = (k0 + ty < K && j < J) ? Bm[(int64_t)(k0 + ty) * J + j] : 0.f;
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc += (double)sa[ty][k] * (double)sb[k][tx];
    __syncthreads();
  }
  if (n < N && j < J) P[(int64_t)n * J + j] = (float)acc;
}
void launch_fold_product(cudaStream_t st, const float* A, const float* Bm, int N, int K, int J, float* P) {
  fold_product_kernel<<<dim3((J + 31) / 32, (N + 31) / 32), dim3(32, 32), 0, st>>>(A, Bm, N, K, J, P);
  MG_CHECK_CUDA(cudaGetLastError());
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
    // experiment switch (tools/ab_two_lanes.py): run the step on fewer CTAs
    if (n > 0 && getenv("MG_MEGA_CTAS")) n = std::max(8, std::min(n, atoi(getenv("MG_MEGA_CTAS"))));
  }
  return n;
}

void launch_decode_step(cudaStream_t st, const MegaParams& p, int n_ctas) {
  MG_REQUIRE(p.B >= 1 && p.B <= MK_R, "fused decode step: 1 <= B <= 32");
  MG_REQUIRE(p.Mp % 8 == 0 && p.Mp <= MK_MAXSC && p.Tp % 32 == 0 && p.Tp <= 512, "fused decode step: Mp / max_length out of range");
  MG_REQUIRE(p.D == p.H * 64 && p.D <= 1024 && p.D % 128 == 0 && p.DFF % 64 == 0, "fused decode step: unsupported dims");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(MK_THREADS);
  cfg.dynamicSmemBytes = MK_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  // MG_MEGA_NOCOOP=1 (experiments with two concurrent instances): plain launch; co-residency then rests on
  // 1 CTA per SM and instances that together fill at most the chip
  static const bool nocoop = getenv("MG_MEGA_NOCOOP") && getenv("MG_MEGA_NOCOOP")[0] == '1';
  cfg.numAttrs = nocoop ? 0 : 1;
  MG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_step_kernel, p));
}

}  // namespace mg
