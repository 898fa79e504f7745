// Host orchestration of the hot path and the model-level C ABI (mg_create ... mg_generate).
// The layer loops below are the B200 replacement of
//   * the fork's Markushgrapher encoder forward (Swin branch + projector + UdopStack encoder + fusion concat;
//     stock restatement transformers/models/udop/modeling_udop.py:1064-1256, models/swin/modeling_swin.py:534-913),
//   * GenerationMixin.generate / _sample (transformers/generation/utils.py:2131, 2658-2842) for greedy decode.
#include <cuda.h>  // types of the green-context driver API (entry points come from cudaGetDriverEntryPoint: no libcuda link)
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.h"
#include "mg_b200.h"

namespace mg {

extern thread_local std::string g_last_error;
int set_error(const Error& e);
int set_error(const std::exception& e);

static inline int64_t rup(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// NCCL is resolved at run time (dlopen of the libnccl.so.2 already loaded by torch, else the system one) so the
// library has no link-time dependency on it; only the multi-GPU entry points touch it.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static bool loaded = false;
  if (!loaded) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    MG_REQUIRE(h != nullptr, std::string("cannot load libnccl.so.2: ") + dlerror());
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    MG_REQUIRE(api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy, "libnccl lacks a required symbol");
    loaded = true;
  }
  return api;
}
#define MG_CHECK_NCCL(expr)                                                                                  \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess)                                                                                   \
      throw mg::Error(-5, std::string(#expr) + " failed: " +                                                 \
                              (nccl_api().GetErrorString ? nccl_api().GetErrorString(_r) : "nccl error"));   \
  } while (0)

// ------------------------------------------------------------------------------------------------ memory
// Chunked bump allocator: deterministic allocation sequences reuse the same addresses after reset().
struct Arena {
  struct Chunk {
    char* base;
    size_t cap, used;
  };
  std::vector<Chunk> chunks;
  size_t chunk_bytes = (size_t)1 << 30;
  size_t total = 0;
  void* alloc(size_t bytes) {
    bytes = (size_t)rup((int64_t)std::max<size_t>(bytes, 16), 1024);
    for (auto& c : chunks) {
      if (c.cap - c.used >= bytes) {
        void* p = c.base + c.used;
        c.used += bytes;
        return p;
      }
    }
    Chunk c;
    c.cap = std::max(bytes, chunk_bytes);
    MG_CHECK_CUDA(cudaMalloc((void**)&c.base, c.cap));
    // a chunk starts out zeroed (once, at allocation): padding rows / columns that no kernel writes then hold the same
    // bytes whether the pages are fresh from the driver or recycled from an earlier allocation of this process
    // (MG_ARENA_NOZERO: off for compute-sanitizer initcheck runs; MG_ARENA_POISON: 0xFF = NaN patterns instead, to
    // flush out kernels whose results depend on bytes nobody wrote)
    if (getenv("MG_ARENA_POISON")) MG_CHECK_CUDA(cudaMemset(c.base, 0xff, c.cap));
    else if (!getenv("MG_ARENA_NOZERO")) MG_CHECK_CUDA(cudaMemset(c.base, 0, c.cap));
    c.used = bytes;
    total += c.cap;
    chunks.push_back(c);
    return c.base;
  }
  template <typename T>
  T* get(int64_t n) {
    return reinterpret_cast<T*>(alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1)));
  }
  void reset() {
    for (auto& c : chunks) c.used = 0;
  }
  void release() {
    for (auto& c : chunks) cudaFree(c.base);
    chunks.clear();
    total = 0;
  }
};

struct RawWeight {
  const float* ptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// a Linear layer repacked as split planes [N, ldk] (+ optional fp32 bias [N])
struct LinearW {
  Planes w;
  int N = 0, K = 0;
  int64_t ldk = 0;
  const float* bias = nullptr;
};

struct EncLayer {
  float *ln1, *ln2;
  LinearW qk, v, o, wi, wo;
};
struct DecLayer {
  float *ln1, *ln2, *ln3;
  LinearW qkv, o, cq, co, wi, wo;
  LinearW ck, cv;
};
struct SwBlock {
  float *ln1w, *ln1b, *ln2w, *ln2b, *table;
  LinearW qkv, proj, fc1, fc2;
};
struct SwStage {
  int C, heads, res;
  std::vector<SwBlock> blocks;
  bool has_down = false;
  float *dnw = nullptr, *dnb = nullptr;
  LinearW down;
};

}  // namespace mg

using namespace mg;

struct mg_model {
  mg_config cfg;
  bool finalized = false;
  std::unordered_map<std::string, RawWeight> raw;
  std::vector<void*> owned;
  bool split2 = true;  // hi+lo planes

  // weights
  float* shared = nullptr;
  LinearW patch_embed;
  float *cell_x = nullptr, *cell_y = nullptr, *tab1d = nullptr, *tabh = nullptr, *tabv = nullptr;
  std::vector<EncLayer> enc;
  float* enc_final_ln = nullptr;
  std::vector<DecLayer> dec;
  float *dec_bias = nullptr, *dec_final_ln = nullptr;
  LinearW lm_head;
  LinearW sw_patch;
  float *sw_emb_w = nullptr, *sw_emb_b = nullptr, *sw_ln_w = nullptr, *sw_ln_b = nullptr;
  std::vector<SwStage> sw;
  LinearW proj1, proj2;
  int *lut_enc1d = nullptr, *lut_hv = nullptr, *lut_dec = nullptr;
  int lut_enc1d_n = 0, lut_hv_n = 0, lut_dec_n = 0;

  // derived
  int NP = 0, np_side = 0, n_sw = 0, sw_dim = 0;

  // run state
  Arena persist, scratch;
  int cur_B = 0, cur_Lt = 0, cur_S = 0, cur_Sp = 0, cur_M = 0, cur_Mp = 0;
  float* mem = nullptr;   // [B, Mp, d]
  Planes mem_pl;          // planes of mem
  int* mem_mask = nullptr;  // [B, Mp]
  // decoder-side view of the memory (valid rows only, padded to the batch maximum dMp; ops.cu compact_*): what the cross
  // K/V projection and every decode kernel read.  Equal to mem / mem_pl / mem_mask / cur_Mp when MG_COMPACT=0.
  float* dmem = nullptr;
  Planes dmem_pl;
  int* dmask = nullptr;
  int dMp = 0, dM = 0;  // padded length, longest valid length
  int64_t launches = 0;
  float last_encode_ms = 0.f, last_decode_ms = 0.f;
  float last_loop_ms = 0.f;   // the decode-step loop alone (cross-KV projection excluded), CUDA events on the stream
  int last_loop_steps = 0;
  int last_fused = 0;         // 1 if the last greedy generate used the fused persistent decode-step kernel
  cudaEvent_t ev_loop[2] = {nullptr, nullptr};
  float last_step_p50_ms = 0.f, last_step_p99_ms = 0.f;  // per-step latencies from the %globaltimer stamp of every step
  int64_t last_launches = 0;
  int* pinned_flag = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t own_stream = nullptr;
  cudaStream_t aux_stream = nullptr;  // second micro-batch lane of the greedy decode loop
  cudaEvent_t lane_ev[3] = {nullptr, nullptr, nullptr};
  int prof_bn = 0;
  // multi-GPU (image-batch sharding): one NCCL all-gather of the step's token ids per decode step
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  // token exchange over NVLink peer memory (kernels.h PeerExchange): this rank's buffer, the peers' buffers mapped with
  // cudaIpc, the device array of all bases; px_ok only if EVERY rank could map every peer (else: NCCL all-gather per step)
  int* px_local = nullptr;
  int** px_peers_dev = nullptr;
  std::vector<void*> px_opened;
  int px_bcap = 512;
  bool px_ok = false;
  unsigned px_calls = 0;
  int64_t* dist_all_ids = nullptr;  // (world*B, max_length) on every rank, set per call
  int* dist_chk = nullptr;          // [2 + 2*world] shard-shape handshake of mg_generate_dist
  // buffers of the last generate call (valid until the next encode/generate), for mg_profile_cross_attn
  std::vector<float*> prof_ckt, prof_cv;
  std::vector<uint8_t*> prof_ckv;
  bool prof_kv24 = false;
  float* prof_q = nullptr;
  float* prof_ctx = nullptr;
  // fused persistent decode step (decode_mega.cu): pre-swizzled weight tiles + per-layer table
  int mega_ctas = 0;
  std::vector<MegaLayer> mega_layers;  // weight fields filled at finalize, cache pointers per generate call
  MegaLin mega_lm;
  MegaLayer* mega_layers_dev = nullptr;
  unsigned* mega_bar = nullptr;

  ~mg_model() {
    for (void* p : owned) cudaFree(p);
    persist.release();
    scratch.release();
    if (pinned_flag) cudaFreeHost(pinned_flag);
    for (auto& hs : stage) {
      if (hs.buf) cudaFree(hs.buf);
      if (hs.ready) cudaEventDestroy(hs.ready);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (auto& a : ahead) {
      a.persist.release();
      a.scratch.release();
      if (a.done) cudaEventDestroy(a.done);
      if (a.begin) cudaEventDestroy(a.begin);
    }
    if (part.order) cudaEventDestroy(part.order);
    if (part.s_small) cudaStreamDestroy(part.s_small);
    if (part.s_big) cudaStreamDestroy(part.s_big);
    // (the two green contexts live until the primary context goes away)
    if (own_stream) cudaStreamDestroy(own_stream);
    if (aux_stream) cudaStreamDestroy(aux_stream);
    for (auto& e : lane_ev)
      if (e) cudaEventDestroy(e);
    for (void* q : px_opened) cudaIpcCloseMemHandle(q);
    if (comm) nccl_api().CommDestroy(comm);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    for (auto& e : ev_loop)
      if (e) cudaEventDestroy(e);
  }

  // ---------------------------------------------------------------------------------------------- helpers
  template <typename T>
  T* own(int64_t n) {
    void* p;
    MG_CHECK_CUDA(cudaMalloc(&p, sizeof(T) * (size_t)std::max<int64_t>(n, 1)));
    owned.push_back(p);
    return reinterpret_cast<T*>(p);
  }
  const RawWeight& need(const std::string& name) {
    auto it = raw.find(name);
    if (it == raw.end()) throw Error(-4, "missing weight: " + name);
    return it->second;
  }
  float* copy_f32(cudaStream_t st, const std::string& name, int64_t expect_numel) {
    const RawWeight& w = need(name);
    MG_REQUIRE(w.numel() == expect_numel, "unexpected size for weight " + name);
    float* p = own<float>(expect_numel);
    MG_CHECK_CUDA(cudaMemcpyAsync(p, w.ptr, sizeof(float) * expect_numel, cudaMemcpyDeviceToDevice, st));
    return p;
  }
  Planes new_planes_owned(int64_t n) {
    Planes p;
    p.hi = own<bf16>(n);
    p.lo = split2 ? own<bf16>(n) : nullptr;
    return p;
  }
  // Linear from one or more [n_i, K] weight matrices stacked along the output dim (+ stacked biases)
  LinearW make_linear(cudaStream_t st, const std::vector<std::string>& wnames, const std::vector<std::string>& bnames,
                      int K) {
    LinearW L;
    L.K = K;
    L.ldk = rup(K, 8);
    int N = 0;
    for (auto& n : wnames) {
      const RawWeight& w = need(n);
      MG_REQUIRE(w.numel() % K == 0, "weight " + n + " is not [*, K]");
      N += (int)(w.numel() / K);
    }
    L.N = N;
    L.w = new_planes_owned((int64_t)N * L.ldk);
    int row = 0;
    for (auto& n : wnames) {
      const RawWeight& w = need(n);
      const int rows = (int)(w.numel() / K);
      Planes dst{L.w.hi + (int64_t)row * L.ldk, L.w.lo ? L.w.lo + (int64_t)row * L.ldk : nullptr};
      launch_split(st, w.ptr, rows, K, K, dst, L.ldk);
      row += rows;
    }
    if (!bnames.empty()) {
      float* b = own<float>(N);
      int off = 0;
      for (auto& n : bnames) {
        const RawWeight& w = need(n);
        MG_CHECK_CUDA(cudaMemcpyAsync(b + off, w.ptr, sizeof(float) * w.numel(), cudaMemcpyDeviceToDevice, st));
        off += (int)w.numel();
      }
      MG_REQUIRE(off == N, "bias size mismatch for " + wnames[0]);
      L.bias = b;
    }
    return L;
  }
  int* upload_lut(cudaStream_t st, const std::vector<int32_t>& v) {
    int* p = own<int>((int64_t)v.size());
    MG_CHECK_CUDA(cudaMemcpyAsync(p, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));  // v is a temporary
    return p;
  }

  // host-entry staging (mg_generate_host / mg_prefetch_host): two device slots owned by the model, so that the
  // inputs of batch i+1 can travel host->device on the copy stream while batch i is decoding
  struct HostStage {
    void* buf = nullptr;
    size_t cap = 0;
    const void *k_ids = nullptr, *k_box = nullptr, *k_px = nullptr, *k_mask = nullptr;  // host pointers staged here
    int B = 0, Lt = 0;
    bool valid = false;     // holds a staged batch that no generate call has consumed yet
    uint64_t seq = 0;       // staging order (the older slot is recycled first)
    cudaEvent_t ready = nullptr;
    int64_t *d_ids = nullptr, *d_mask = nullptr, *d_out = nullptr;
    float *d_box = nullptr, *d_px = nullptr;
  };
  HostStage stage[2];
  uint64_t stage_seq = 0;
  cudaStream_t copy_stream = nullptr;
  // slot to stage the next batch into: one that holds no unconsumed batch, else the older one.  Calls are made from one
  // thread and mg_generate_host returns only when its batch is done, so a slot without a pending batch is idle.
  int pick_stage_slot() const {
    if (!stage[0].valid) return 0;
    if (!stage[1].valid) return 1;
    return stage[0].seq <= stage[1].seq ? 0 : 1;
  }
  // copies one batch of host inputs into slot `s` on stream `cs` (max_length sizes the ids-out buffer)
  void stage_inputs(HostStage& s, cudaStream_t cs, int B, int Lt, const int64_t* ids, const float* bbox, const float* px,
                    const int64_t* mask, int out_cols) {
    for (auto& a : ahead)  // a run-ahead encoding computed from this slot's previous contents is stale now
      if (a.valid && s.buf && a.k_ids == s.d_ids && a.k_px == s.d_px) a.valid = false;
    const size_t n_ids = (size_t)B * Lt, n_px = (size_t)B * 3 * cfg.image_size * cfg.image_size;
    const size_t need = rup(n_ids * 8, 256) * 2 + rup(n_ids * 16, 256) + rup(n_px * 4, 256) + rup((size_t)B * out_cols * 8, 256);
    if (need > s.cap) {
      if (s.buf) MG_CHECK_CUDA(cudaFree(s.buf));
      s.buf = nullptr;
      s.cap = 0;
      MG_CHECK_CUDA(cudaMalloc(&s.buf, need));
      s.cap = need;
    }
    if (!s.ready) MG_CHECK_CUDA(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
    char* p = static_cast<char*>(s.buf);
    s.d_ids = reinterpret_cast<int64_t*>(p); p += rup(n_ids * 8, 256);
    s.d_mask = reinterpret_cast<int64_t*>(p); p += rup(n_ids * 8, 256);
    s.d_box = reinterpret_cast<float*>(p); p += rup(n_ids * 16, 256);
    s.d_px = reinterpret_cast<float*>(p); p += rup(n_px * 4, 256);
    s.d_out = reinterpret_cast<int64_t*>(p);
    MG_CHECK_CUDA(cudaMemcpyAsync(s.d_ids, ids, n_ids * 8, cudaMemcpyHostToDevice, cs));
    if (mask) MG_CHECK_CUDA(cudaMemcpyAsync(s.d_mask, mask, n_ids * 8, cudaMemcpyHostToDevice, cs));
    MG_CHECK_CUDA(cudaMemcpyAsync(s.d_box, bbox, n_ids * 16, cudaMemcpyHostToDevice, cs));
    MG_CHECK_CUDA(cudaMemcpyAsync(s.d_px, px, n_px * 4, cudaMemcpyHostToDevice, cs));
    MG_CHECK_CUDA(cudaEventRecord(s.ready, cs));
    s.k_ids = ids; s.k_box = bbox; s.k_px = px; s.k_mask = mask;
    s.B = B; s.Lt = Lt;
    s.valid = true;
    s.seq = ++stage_seq;
  }

  // ---- encoder run-ahead (mg_encode_ahead): while batch i decodes -- an HBM-bound loop that leaves the tensor cores
  // idle -- the encoder of batch i+1 runs on a small SM partition of its own (CUDA green contexts); the decode loop
  // then runs on the remaining SMs.  An EncSlot is one encoded batch together with the arenas it lives in; taking a
  // slot swaps it with the model's current view (cur_*, mem*, persist, scratch), so nothing is copied.
  struct EncSlot {
    Arena persist, scratch;
    int B = 0, Lt = 0, S = 0, Sp = 0, M = 0, Mp = 0;
    float* mem = nullptr;
    Planes mem_pl;
    int* mem_mask = nullptr;
    const void *k_ids = nullptr, *k_box = nullptr, *k_px = nullptr, *k_mask = nullptr;  // inputs it was computed from
    bool valid = false;   // holds an encoded batch that no generate call has consumed yet
    uint64_t seq = 0;
    cudaEvent_t begin = nullptr, done = nullptr;  // on the small partition's stream
    int64_t n_launch = 0;
  };
  EncSlot ahead[2];
  float last_ahead_ms = 0.f;  // encoder time of the batch the last generate call took from a slot (on the small partition)
  uint64_t ahead_seq = 0;
  struct Partitions {
    bool tried = false, ok = false;
    CUgreenCtx g_small = nullptr, g_big = nullptr;
    cudaStream_t s_small = nullptr, s_big = nullptr;
    int n_small = 0, n_big = 0;
    cudaEvent_t order = nullptr;
  } part;
  int run_ctas = 0;  // CTAs the fused decode step may use in the current generate call (0 = all SMs)
  void swap_view(EncSlot& s) {
    std::swap(persist, s.persist); std::swap(scratch, s.scratch);
    std::swap(cur_B, s.B); std::swap(cur_Lt, s.Lt); std::swap(cur_S, s.S); std::swap(cur_Sp, s.Sp);
    std::swap(cur_M, s.M); std::swap(cur_Mp, s.Mp);
    std::swap(mem, s.mem); std::swap(mem_pl, s.mem_pl); std::swap(mem_mask, s.mem_mask);
  }
  bool ensure_partitions();
  int find_ahead(int B, int Lt, const void* ids, const void* box, const void* px, const void* mask) const {
    int best = -1;  // the OLDEST match: a caller that reuses its tensors has the next batch armed under the same keys
    for (int i = 0; i < 2; ++i) {
      const EncSlot& s = ahead[i];
      if (s.valid && s.k_ids == ids && s.k_box == box && s.k_px == px && s.k_mask == mask && s.B == B && s.Lt == Lt &&
          (best < 0 || s.seq < ahead[best].seq))
        best = i;
    }
    return best;
  }
  void encode_ahead(cudaStream_t after, int B, int Lt, const int64_t* ids, const float* bbox, const float* px,
                    const int64_t* amask);

  int64_t* ids_buf = nullptr;
  int64_t ids_cap = 0;
  int64_t* persist_ids(int64_t n) {  // scratch for local ids of mg_generate_dist (kept across calls)
    if (n > ids_cap) {
      ids_buf = own<int64_t>(n);
      ids_cap = n;
    }
    return ids_buf;
  }
  Planes planes(Arena& a, int64_t n) {
    Planes p;
    p.hi = a.get<bf16>(n);
    p.lo = split2 ? a.get<bf16>(n) : nullptr;
    return p;
  }
  static Planes offset(Planes p, int64_t off) { return Planes{p.hi + off, p.lo ? p.lo + off : nullptr}; }

  // y[rows, N] = x[rows, K] * W^T  (x planes with row stride ldx)
  void linear(cudaStream_t st, const LinearW& W, Planes x, int64_t ldx, int64_t rows, GemmEpilogue ep,
              bool use_bias = true) {
    GemmOperand A, B;
    A.hi = x.hi; A.lo = x.lo; A.rows = rows; A.ld = ldx;
    B.hi = W.w.hi; B.lo = W.w.lo; B.rows = W.N; B.ld = W.ldk;
    if (use_bias && W.bias) ep.bias = W.bias;
    if (ep.ld_r == 0) ep.ld_r = W.N;
    launch_gemm(st, A, B, (int)rows, W.N, W.K, 1, 1, 1, ep, W.N <= 32 ? 32 : (W.N <= 64 ? 64 : 128));
    ++launches;
  }

  // Decode-chain linear for MORE than 128 activation rows (beam search over large batches, batches > 128 per GPU;
  // measured: 500 rows 18.7 -> 13.4 ms per step, but 128 rows 6.37 -> 6.66, so one skinny launch keeps those): the activations are
  // normalised / rectified and split into planes ONCE by a small kernel, then the persistent TMA-fed tcgen05 GEMM runs
  // with the weights on the M axis and split-K atomics into the output -- instead of skinny_tc_kernel<4>, where every
  // CTA re-loads and re-splits the whole activation tile (one L2 round trip per k-block) and launches of 128 rows each
  // re-read the weights.  out[b][n] += sum_k pro(x)[b][k] W[n][k];  pro: 0 none, 1 RMSNorm (weight lnw), 2 ReLU.
  int wide_linear(cudaStream_t st, int pro, const float* x, int ldx, const LinearW& W, float* out, int ld_out,
                  const float* lnw, float scale, int rows, Planes xs, float* zero_ptr, int64_t zero_n) {
    int n = 0;
    if (pro == 1) {
      MG_REQUIRE(ldx == W.K, "wide linear: RMSNorm input must be contiguous rows");
      launch_rmsnorm(st, x, lnw, rows, W.K, cfg.ln_eps, scale, xs, nullptr, 0, 0, 0);
    } else if (pro == 2) {
      MG_REQUIRE(ldx == W.K, "wide linear: ReLU input must be contiguous rows");
      launch_relu_split(st, x, (int64_t)rows * W.K, xs);
    } else {
      launch_split(st, x, rows, W.K, ldx, xs, W.K);
    }
    ++n;
    if (zero_ptr) MG_CHECK_CUDA(cudaMemsetAsync(zero_ptr, 0, sizeof(float) * (size_t)zero_n, st));
    GemmOperand A, B;
    A.hi = W.w.hi; A.lo = W.w.lo; A.rows = W.N; A.ld = W.ldk;
    B.hi = xs.hi; B.lo = xs.lo; B.rows = rows; B.ld = W.K;
    GemmEpilogue ep;
    ep.out_f32 = out;
    ep.ld_r = 1;        // feature index (GEMM row) is the fast axis of out[b][n]
    ep.ld_c = ld_out;   // activation row (GEMM column) strides by the output's leading dimension
    ep.atomic = 1;
    const int tiles = ((W.N + 127) / 128) * ((rows + 127) / 128);
    const int ksplit = std::max(1, std::min((W.K + 63) / 64, 148 / std::max(1, tiles)));
    launch_gemm(st, A, B, W.N, rows, W.K, 1, 1, ksplit, ep, 128);
    return n + 1;
  }

  void finalize(cudaStream_t st);
  void encode(cudaStream_t st, int B, int Lt, const int64_t* ids, const float* bbox, const float* px,
              const int64_t* amask);
  void swin_forward(cudaStream_t st, int B0, int Bc, const float* px);
  void vtl_forward(cudaStream_t st, int B0, int Bc, int Lt, const int64_t* ids, const float* bbox, const float* px,
                   const int64_t* amask);
  void project_cross_kv(cudaStream_t st, int B, std::vector<float*>& ckt, std::vector<float*>& cv,
                        std::vector<uint8_t*>* ckv24 = nullptr);
  void prepare_decoder_memory(cudaStream_t st, int B);
  void generate(cudaStream_t st, int B, int max_length, int64_t* out_ids, int32_t* out_len, float* step_logits,
                int32_t* steps_run, const int64_t* forced = nullptr, int forced_ld = 0);
  void generate_beam(cudaStream_t st, int B, int nb, int max_length, int64_t* out_ids, int32_t* out_len,
                     int32_t* steps_run);
};

// ================================================================================================= finalize
void mg_model::finalize(cudaStream_t st) {
  const mg_config& c = cfg;
  MG_REQUIRE(c.d_kv == 64 && c.num_heads * c.d_kv == c.d_model, "this build needs d_kv == 64 and heads*d_kv == d_model");
  MG_REQUIRE(c.d_model % 128 == 0 && c.d_model <= 1024, "d_model must be a multiple of 128, <= 1024");
  MG_REQUIRE(c.image_size % c.patch_size == 0, "image_size must be a multiple of patch_size");
  split2 = c.precision == 0;
  const int d = c.d_model, H = c.num_heads, V = c.vocab_size;
  np_side = c.image_size / c.patch_size;
  NP = np_side * np_side;

  shared = copy_f32(st, "shared.weight", (int64_t)V * d);
  patch_embed = make_linear(st, {"encoder.embed_patches.proj.weight"}, {"encoder.embed_patches.proj.bias"},
                            3 * c.patch_size * c.patch_size);
  cell_x = copy_f32(st, "encoder.cell_2d_embedding.x_position_embeddings.weight", (int64_t)c.max_2d * d);
  cell_y = copy_f32(st, "encoder.cell_2d_embedding.y_position_embeddings.weight", (int64_t)c.max_2d * d);
  tab1d = copy_f32(st, "encoder.relative_bias.biases.0.relative_attention_bias.weight", (int64_t)c.rel_buckets * H);
  tabh = copy_f32(st, "encoder.relative_bias.biases.1.relative_attention_bias.weight", (int64_t)c.rel_buckets * H);
  tabv = copy_f32(st, "encoder.relative_bias.biases.2.relative_attention_bias.weight", (int64_t)c.rel_buckets * H);
  enc.resize(c.num_layers);
  for (int i = 0; i < c.num_layers; ++i) {
    const std::string p = "encoder.block." + std::to_string(i) + ".layer.";
    EncLayer& L = enc[i];
    L.ln1 = copy_f32(st, p + "0.layer_norm.weight", d);
    L.qk = make_linear(st, {p + "0.SelfAttention.q.weight", p + "0.SelfAttention.k.weight"}, {}, d);
    L.v = make_linear(st, {p + "0.SelfAttention.v.weight"}, {}, d);
    L.o = make_linear(st, {p + "0.SelfAttention.o.weight"}, {}, d);
    L.ln2 = copy_f32(st, p + "1.layer_norm.weight", d);
    L.wi = make_linear(st, {p + "1.DenseReluDense.wi.weight"}, {}, d);
    L.wo = make_linear(st, {p + "1.DenseReluDense.wo.weight"}, {}, c.d_ff);
  }
  enc_final_ln = copy_f32(st, "encoder.final_layer_norm.weight", d);
  dec.resize(c.num_decoder_layers);
  for (int i = 0; i < c.num_decoder_layers; ++i) {
    const std::string p = "decoder.block." + std::to_string(i) + ".layer.";
    DecLayer& L = dec[i];
    L.ln1 = copy_f32(st, p + "0.layer_norm.weight", d);
    L.qkv = make_linear(st, {p + "0.SelfAttention.q.weight", p + "0.SelfAttention.k.weight", p + "0.SelfAttention.v.weight"}, {}, d);
    L.o = make_linear(st, {p + "0.SelfAttention.o.weight"}, {}, d);
    L.ln2 = copy_f32(st, p + "1.layer_norm.weight", d);
    L.cq = make_linear(st, {p + "1.EncDecAttention.q.weight"}, {}, d);
    L.ck = make_linear(st, {p + "1.EncDecAttention.k.weight"}, {}, d);
    L.cv = make_linear(st, {p + "1.EncDecAttention.v.weight"}, {}, d);
    L.co = make_linear(st, {p + "1.EncDecAttention.o.weight"}, {}, d);
    L.ln3 = copy_f32(st, p + "2.layer_norm.weight", d);
    L.wi = make_linear(st, {p + "2.DenseReluDense.wi.weight"}, {}, d);
    L.wo = make_linear(st, {p + "2.DenseReluDense.wo.weight"}, {}, c.d_ff);
  }
  dec_bias = copy_f32(st, "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", (int64_t)c.rel_buckets * H);
  dec_final_ln = copy_f32(st, "decoder.final_layer_norm.weight", d);
  lm_head = make_linear(st, {"lm_head.weight"}, {}, d);
  MG_REQUIRE(d % 64 == 0 && c.d_ff % 64 == 0, "d_model and d_ff must be multiples of 64");

  // ---- Swin
  const std::string sp = "encoder.molscribe_encoder.";
  MG_REQUIRE(c.swin_num_stages >= 1 && c.swin_num_stages <= MG_MAX_SWIN_STAGES, "bad swin_num_stages");
  MG_REQUIRE(c.swin_image % c.swin_patch == 0, "swin image must be a multiple of its patch size");
  sw_patch = make_linear(st, {sp + "embeddings.patch_embeddings.projection.weight"},
                         {sp + "embeddings.patch_embeddings.projection.bias"}, 3 * c.swin_patch * c.swin_patch);
  sw_emb_w = copy_f32(st, sp + "embeddings.norm.weight", c.swin_embed);
  sw_emb_b = copy_f32(st, sp + "embeddings.norm.bias", c.swin_embed);
  sw.resize(c.swin_num_stages);
  int res = c.swin_image / c.swin_patch;
  int C = c.swin_embed;
  const int ntab = (2 * c.swin_window - 1) * (2 * c.swin_window - 1);
  for (int s = 0; s < c.swin_num_stages; ++s) {
    SwStage& S = sw[s];
    S.C = C;
    S.heads = c.swin_heads[s];
    S.res = res;
    MG_REQUIRE(C / S.heads == 32 && C % S.heads == 0, "Swin head_dim must be 32");
    MG_REQUIRE(res % std::min(res, c.swin_window) == 0 && res >= c.swin_window,
               "Swin stage resolution must be a multiple of (and at least) the window size");
    S.blocks.resize(c.swin_depths[s]);
    for (int b = 0; b < c.swin_depths[s]; ++b) {
      const std::string p = sp + "encoder.layers." + std::to_string(s) + ".blocks." + std::to_string(b) + ".";
      SwBlock& K = S.blocks[b];
      K.ln1w = copy_f32(st, p + "layernorm_before.weight", C);
      K.ln1b = copy_f32(st, p + "layernorm_before.bias", C);
      K.qkv = make_linear(st, {p + "attention.self.query.weight", p + "attention.self.key.weight", p + "attention.self.value.weight"},
                          {p + "attention.self.query.bias", p + "attention.self.key.bias", p + "attention.self.value.bias"}, C);
      K.table = copy_f32(st, p + "attention.self.relative_position_bias_table", (int64_t)ntab * S.heads);
      K.proj = make_linear(st, {p + "attention.output.dense.weight"}, {p + "attention.output.dense.bias"}, C);
      K.ln2w = copy_f32(st, p + "layernorm_after.weight", C);
      K.ln2b = copy_f32(st, p + "layernorm_after.bias", C);
      K.fc1 = make_linear(st, {p + "intermediate.dense.weight"}, {p + "intermediate.dense.bias"}, C);
      K.fc2 = make_linear(st, {p + "output.dense.weight"}, {p + "output.dense.bias"}, 4 * C);
    }
    if (s + 1 < c.swin_num_stages) {
      const std::string p = sp + "encoder.layers." + std::to_string(s) + ".downsample.";
      S.has_down = true;
      S.dnw = copy_f32(st, p + "norm.weight", 4 * C);
      S.dnb = copy_f32(st, p + "norm.bias", 4 * C);
      S.down = make_linear(st, {p + "reduction.weight"}, {}, 4 * C);
      MG_REQUIRE(res % 2 == 0, "Swin stage resolution must be even for patch merging");
      res /= 2;
      C *= 2;
    }
  }
  sw_dim = C;
  n_sw = res * res;
  sw_ln_w = copy_f32(st, sp + "layernorm.weight", C);
  sw_ln_b = copy_f32(st, sp + "layernorm.bias", C);
  proj1 = make_linear(st, {"encoder.molscribe_projector.0.weight"}, {"encoder.molscribe_projector.0.bias"}, C);
  proj2 = make_linear(st, {"encoder.molscribe_projector.2.weight"}, {"encoder.molscribe_projector.2.bias"}, c.proj_hidden);
  MG_REQUIRE(proj1.N == c.proj_hidden && proj2.N == d, "projector shapes do not match the config");

  // ---- integer LUTs of the bucket functions
  {
    std::vector<int32_t> l(c.rel_max_distance + 1);
    rel_bucket_lut(1, c.rel_buckets, c.rel_max_distance, (int)l.size(), l.data());
    lut_enc1d = upload_lut(st, l);
    lut_enc1d_n = (int)l.size();
    std::vector<int32_t> lh(101);
    rel_bucket_lut(1, c.rel_buckets, 100, (int)lh.size(), lh.data());  // RelativePositionBiasHorizontal/Vertical: max_distance 100
    lut_hv = upload_lut(st, lh);
    lut_hv_n = (int)lh.size();
    std::vector<int32_t> ld(c.rel_max_distance + 1);
    rel_bucket_lut(0, c.rel_buckets, c.rel_max_distance, (int)ld.size(), ld.data());
    // the decode kernel indexes lut[step - j] directly: extend to any distance by clamping on the host
    std::vector<int32_t> ldx(4096);
    for (int i = 0; i < 4096; ++i) ldx[i] = ld[std::min<int>(i, c.rel_max_distance)];
    lut_dec = upload_lut(st, ldx);
    lut_dec_n = 4096;
  }
  // ---- fused decode step: decoder weights re-tiled as a stream of pre-swizzled 32 KB shared-memory images
  {
    const char* mode = getenv("MG_DECODE");
    const bool want = !(mode && std::string(mode) == "chain");
    mega_ctas = (want && split2 && d % 128 == 0) ? mega_max_ctas() : 0;  // two-output linears split on 128-row tiles
    if (mega_ctas > 0) {
      const int NL = c.num_decoder_layers, dff = c.d_ff;
      // phases 0 and 2 are two-output linears: [Wqkv diag(ln1); Wcq diag(ln2)] on x and [Wo; (Wcq diag(ln2)) Wo] on the
      // self-attention context together give the cross query (decode_mega.cu); built here from the fp32 weights
      const bool fold_ff = mega_fold_ff();  // co folded into wi: one three-segment linear [Wco; Wi' Wco; Wi'] of d + 2 dff rows
      size_t per_layer = mega_lin_bytes(4 * d, d) + mega_lin_bytes(2 * d, d) + mega_lin_bytes(d, dff) +
                         (fold_ff ? mega_lin_bytes(d + 2 * dff, d) : mega_lin_bytes(d, d) + mega_lin_bytes(dff, d));
      uint8_t* buf = own<uint8_t>((int64_t)(per_layer * NL + mega_lin_bytes(V, d)));
      mega_layers.resize(NL);
      auto tile = [&](const LinearW& W, bool store) {
        MegaLin L = make_mega_lin(st, W.w, W.N, W.K, W.ldk, store, mega_ctas, buf);
        buf += mega_lin_bytes(W.N, W.K);
        return L;
      };
      float* fold = nullptr;  // [max(4d, d + 2 dff)][d] fp32 scratch + planes of the same shape, freed below
      bf16* fold_pl = nullptr;
      const int64_t ldk = rup(d, 8);
      const int64_t fold_rows = std::max<int64_t>(4 * d, fold_ff ? d + 2 * dff : 0);
      MG_CHECK_CUDA(cudaMalloc((void**)&fold, sizeof(float) * (size_t)fold_rows * d));
      MG_CHECK_CUDA(cudaMalloc((void**)&fold_pl, sizeof(bf16) * (size_t)2 * fold_rows * ldk));
      const Planes fp{fold_pl, fold_pl + fold_rows * ldk};
      auto tile_fold = [&](int N, int n_split, int n_split2 = 0) {
        launch_split(st, fold, N, d, d, fp, ldk);
        MegaLin L = make_mega_lin(st, fp, N, d, ldk, false, mega_ctas, buf, n_split, n_split2);
        buf += mega_lin_bytes(N, d);
        return L;
      };
      for (int i = 0; i < NL; ++i) {
        MegaLayer& M = mega_layers[i];
        const DecLayer& L = dec[i];
        const std::string p = "decoder.block." + std::to_string(i) + ".layer.";
        const char* qkv_names[3] = {"0.SelfAttention.q.weight", "0.SelfAttention.k.weight", "0.SelfAttention.v.weight"};
        for (int j = 0; j < 3; ++j) {
          const RawWeight& w = need(p + qkv_names[j]);
          MG_REQUIRE(w.numel() == (int64_t)d * d, "decoder self-attention weight is not [d_model, d_model]");
          launch_scale_cols(st, w.ptr, L.ln1, d, d, fold + (int64_t)j * d * d);
        }
        launch_scale_cols(st, need(p + "1.EncDecAttention.q.weight").ptr, L.ln2, d, d, fold + (int64_t)3 * d * d);
        M.lin[0] = tile_fold(4 * d, 3 * d);
        // rows [d, 2d) = (Wcq diag(ln2)) Wo, from the scaled copy still in fold rows [3d, 4d)
        launch_fold_product(st, fold + (int64_t)3 * d * d, need(p + "0.SelfAttention.o.weight").ptr, d, d, d, fold + (int64_t)d * d);
        MG_CHECK_CUDA(cudaMemcpyAsync(fold, need(p + "0.SelfAttention.o.weight").ptr, sizeof(float) * (size_t)d * d, cudaMemcpyDeviceToDevice, st));
        M.lin[1] = tile_fold(2 * d, d);
        if (fold_ff) {
          // rows [0, d) = Wco; [d, d + dff) = (Wi diag(ln3)) Wco (fp64 accumulation); [d + dff, d + 2 dff) = Wi diag(ln3)
          MG_REQUIRE(need(p + "1.EncDecAttention.o.weight").numel() == (int64_t)d * d &&
                         need(p + "2.DenseReluDense.wi.weight").numel() == (int64_t)dff * d, "decoder co / wi weight shapes");
          float* wi_s = fold + (int64_t)(d + dff) * d;
          launch_scale_cols(st, need(p + "2.DenseReluDense.wi.weight").ptr, L.ln3, dff, d, wi_s);
          launch_fold_product(st, wi_s, need(p + "1.EncDecAttention.o.weight").ptr, dff, d, d, fold + (int64_t)d * d);
          MG_CHECK_CUDA(cudaMemcpyAsync(fold, need(p + "1.EncDecAttention.o.weight").ptr, sizeof(float) * (size_t)d * d, cudaMemcpyDeviceToDevice, st));
          M.lin[2] = tile_fold(d + 2 * dff, d, d + dff);
          M.lin[3] = tile(L.wo, false);
          M.lin[4] = M.lin[3];
        } else {
          M.lin[2] = tile(L.co, false); M.lin[3] = tile(L.wi, false); M.lin[4] = tile(L.wo, false);
        }
        M.ln[0] = L.ln1; M.ln[1] = L.ln2; M.ln[2] = L.ln3;
        M.skb = M.svb = nullptr; M.ckv = nullptr; M.pad_ = nullptr;
      }
      MG_CHECK_CUDA(cudaStreamSynchronize(st));
      MG_CHECK_CUDA(cudaFree(fold));
      MG_CHECK_CUDA(cudaFree(fold_pl));
      mega_lm = tile(lm_head, true);
      mega_layers_dev = own<MegaLayer>(NL);
      mega_bar = own<unsigned>(4 + 256);  // [2] barrier counters (+2 spare), then one arrival flag per CTA (-DMK_FLAGBAR build)
      MG_CHECK_CUDA(cudaMemsetAsync(mega_bar, 0, (4 + 256) * sizeof(unsigned), st));
    }
  }
  MG_CHECK_CUDA(cudaMallocHost((void**)&pinned_flag, 64));
  for (auto& e : ev) MG_CHECK_CUDA(cudaEventCreate(&e));
  for (auto& e : ev_loop) MG_CHECK_CUDA(cudaEventCreate(&e));
  MG_CHECK_CUDA(cudaStreamSynchronize(st));
  raw.clear();
  finalized = true;
}

// ================================================================================================= Swin branch
void mg_model::swin_forward(cudaStream_t st, int B0, int Bc, const float* px_all) {
  const mg_config& c = cfg;
  Arena& a = scratch;
  const int d = c.d_model;
  const float* px = px_all + (int64_t)B0 * 3 * c.image_size * c.image_size;
  const int SI = c.swin_image;
  if (c.image_size != SI) {
    float* rs = a.get<float>((int64_t)Bc * 3 * SI * SI);
    launch_resize_bilinear(st, px, Bc, c.image_size, c.image_size, SI, SI, rs);
    ++launches;
    px = rs;
  }
  int res = SI / c.swin_patch;
  int64_t T = (int64_t)Bc * res * res;
  // patch embedding + LayerNorm
  Planes col = planes(a, T * sw_patch.ldk);
  launch_im2col(st, px, Bc, SI, SI, c.swin_patch, (int)sw_patch.ldk, col);
  ++launches;
  float* pe = a.get<float>(T * c.swin_embed);
  {
    GemmEpilogue ep;
    ep.out_f32 = pe;
    linear(st, sw_patch, col, sw_patch.ldk, T, ep);
  }
  float* x = a.get<float>(T * c.swin_embed);
  launch_layernorm_any(st, pe, nullptr, 1, c.swin_embed, T, sw_emb_w, sw_emb_b, c.swin_ln_eps, Planes{}, x);
  ++launches;

  for (size_t s = 0; s < sw.size(); ++s) {
    SwStage& S = sw[s];
    const int C = S.C;
    res = S.res;
    T = (int64_t)Bc * res * res;
    const int ws = std::min(c.swin_window, res);
    int* map0 = a.get<int>(T);
    int* map1 = a.get<int>(T);
    launch_window_rowmap(st, Bc, res, res, ws, 0, map0);
    ++launches;
    const bool can_shift = res > c.swin_window;
    if (can_shift) {
      launch_window_rowmap(st, Bc, res, res, ws, ws / 2, map1);
      ++launches;
    }
    Planes xw = planes(a, T * C);
    float* qkv = a.get<float>(T * 3 * C);
    Planes ctx = planes(a, T * C);
    Planes hid = planes(a, T * 4 * C);
    for (size_t b = 0; b < S.blocks.size(); ++b) {
      SwBlock& K = S.blocks[b];
      const int shift = (b % 2 == 1 && can_shift) ? ws / 2 : 0;
      const int* map = shift ? map1 : map0;
      launch_layernorm_any(st, x, map, 1, C, T, K.ln1w, K.ln1b, c.swin_ln_eps, xw, nullptr);
      ++launches;
      {
        GemmEpilogue ep;
        ep.out_f32 = qkv;
        linear(st, K.qkv, xw, C, T, ep);
      }
      launch_window_attn(st, qkv, K.table, T / (ws * ws), C, S.heads, ws, res, res, shift, ctx);
      ++launches;
      {
        GemmEpilogue ep;  // proj + bias, scattered back to spatial order and added to the shortcut (in place)
        ep.out_f32 = x;
        ep.residual = x;
        ep.row_map = map;
        ep.ld_r = C;
        linear(st, K.proj, ctx, C, T, ep);
      }
      launch_layernorm_any(st, x, nullptr, 1, C, T, K.ln2w, K.ln2b, c.swin_ln_eps, xw, nullptr);
      ++launches;
      {
        GemmEpilogue ep;
        ep.out_hi = hid.hi;
        ep.out_lo = hid.lo;
        ep.act = ACT_GELU_ERF;
        linear(st, K.fc1, xw, C, T, ep);
      }
      {
        GemmEpilogue ep;
        ep.out_f32 = x;
        ep.residual = x;
        linear(st, K.fc2, hid, 4 * C, T, ep);
      }
    }
    if (S.has_down) {
      const int64_t T2 = T / 4;
      int* mm = a.get<int>(T2 * 4);
      launch_merge_rowmap(st, Bc, res, res, mm);
      ++launches;
      Planes mg4 = planes(a, T2 * 4 * C);
      launch_layernorm_any(st, x, mm, 4, C, T2, S.dnw, S.dnb, c.swin_ln_eps, mg4, nullptr);
      ++launches;
      float* xn = a.get<float>(T2 * 2 * C);
      GemmEpilogue ep;
      ep.out_f32 = xn;
      linear(st, S.down, mg4, 4 * C, T2, ep);
      x = xn;
    }
  }
  // final LayerNorm -> projector -> e1 rows [0, n_sw) of the encoder memory
  const int64_t Tn = (int64_t)Bc * n_sw;
  Planes f = planes(a, Tn * sw_dim);
  launch_layernorm_any(st, x, nullptr, 1, sw_dim, Tn, sw_ln_w, sw_ln_b, c.swin_ln_eps, f, nullptr);
  ++launches;
  Planes ph = planes(a, Tn * c.proj_hidden);
  {
    GemmEpilogue ep;
    ep.out_hi = ph.hi;
    ep.out_lo = ph.lo;
    ep.act = ACT_GELU_ERF;
    linear(st, proj1, f, sw_dim, Tn, ep);
  }
  {
    GemmOperand A, Bop;
    A.hi = ph.hi; A.lo = ph.lo; A.rows = n_sw; A.ld = c.proj_hidden; A.use_b2 = true; A.bs2 = (int64_t)n_sw * c.proj_hidden;
    Bop.hi = proj2.w.hi; Bop.lo = proj2.w.lo; Bop.rows = proj2.N; Bop.ld = proj2.ldk;
    GemmEpilogue ep;
    ep.out_f32 = mem + (int64_t)B0 * cur_Mp * d;
    ep.ld_r = d;
    ep.bs2 = (int64_t)cur_Mp * d;
    ep.bias = proj2.bias;
    launch_gemm(st, A, Bop, n_sw, d, c.proj_hidden, 1, Bc, 1, ep, 128);
    ++launches;
  }
}

// ================================================================================================= VTL encoder
void mg_model::vtl_forward(cudaStream_t st, int B0, int Bc, int Lt, const int64_t* ids, const float* bbox,
                           const float* px_all, const int64_t* amask) {
  const mg_config& c = cfg;
  Arena& a = scratch;
  const int d = c.d_model, H = c.num_heads, Sp = cur_Sp, S = cur_S;
  const float* px = px_all + (int64_t)B0 * 3 * c.image_size * c.image_size;
  ids += (int64_t)B0 * Lt;
  bbox += (int64_t)B0 * Lt * 4;
  if (amask) amask += (int64_t)B0 * Lt;
  const int64_t T = (int64_t)Bc * Sp;

  // patch embeddings (Conv2d k16 s16 == GEMM over im2col rows)
  Planes col = planes(a, (int64_t)Bc * NP * patch_embed.ldk);
  launch_im2col(st, px, Bc, c.image_size, c.image_size, c.patch_size, (int)patch_embed.ldk, col);
  ++launches;
  float* pemb = a.get<float>((int64_t)Bc * NP * d);
  {
    GemmEpilogue ep;
    ep.out_f32 = pemb;
    linear(st, patch_embed, col, patch_embed.ldk, (int64_t)Bc * NP, ep);
  }
  // token/patch fusion + compaction + cell embeddings
  float* x = a.get<float>(T * d);
  double* bbox_ext = a.get<double>(T * 4);
  int* vmask = a.get<int>(T);
  int* ocr_pt = a.get<int>((int64_t)Bc * Lt);
  int* vis_src = a.get<int>((int64_t)Bc * NP);
  int* n_vis = a.get<int>(Bc);
  launch_combine(st, ids, bbox, amask, shared, pemb, cell_x, cell_y, Bc, Lt, np_side, Sp, d, c.max_2d, c.vocab_size,
                 ocr_pt, vis_src, n_vis, x, bbox_ext, vmask);
  launches += 2;
  // encoder attention: one fused tcgen05 flash kernel per layer (enc_flash.cu); MG_ENC_ATTN=unfused keeps the round-1
  // chain (scores GEMM -> softmax kernel -> P.V GEMM through a materialised (B,H,S,S) tensor) for A/B runs
  static const bool env_unfused = getenv("MG_ENC_ATTN") && std::string(getenv("MG_ENC_ATTN")) == "unfused";
  const bool flash = !env_unfused && split2 && c.rel_buckets <= 32 && Sp <= 1632;
  uchar2* hv = nullptr;
  uint16_t* code = nullptr;
  if (flash) {
    code = reinterpret_cast<uint16_t*>(a.alloc(enc_bias_code_bytes(Bc, Sp)));
    launch_enc_bias_code(st, bbox_ext, vmask, Bc, Sp, lut_hv, lut_hv_n, c.rel_buckets / 2, code);
  } else {
    hv = a.get<uchar2>(T * Sp);
    launch_relbucket_hv(st, bbox_ext, Bc, Sp, lut_hv, lut_hv_n, c.rel_buckets / 2, hv);
  }
  ++launches;

  Planes xn = planes(a, T * d);
  Planes qk = planes(a, T * 2 * d);
  Planes vt = planes(a, T * d);  // [Bc][d][Sp]
  float* scores = flash ? nullptr : a.get<float>((int64_t)Bc * H * Sp * Sp);
  Planes P = flash ? Planes{} : planes(a, (int64_t)Bc * H * Sp * Sp);
  Planes ctx = planes(a, T * d);
  Planes hid = planes(a, T * c.d_ff);

  for (int l = 0; l < c.num_layers; ++l) {
    EncLayer& L = enc[l];
    launch_rmsnorm(st, x, L.ln1, T, d, c.ln_eps, 1.f, xn, nullptr, 0, 0, 0);
    ++launches;
    {
      GemmEpilogue ep;
      ep.out_hi = qk.hi;
      ep.out_lo = qk.lo;
      linear(st, L.qk, xn, d, T, ep);
    }
    {  // V^T per image: vt[b][feature][token] = Wv . xn[b]^T
      GemmOperand A, Bop;
      A.hi = L.v.w.hi; A.lo = L.v.w.lo; A.rows = d; A.ld = L.v.ldk;
      Bop.hi = xn.hi; Bop.lo = xn.lo; Bop.rows = Sp; Bop.ld = d; Bop.use_b2 = true; Bop.bs2 = (int64_t)Sp * d;
      GemmEpilogue ep;
      ep.out_hi = vt.hi;
      ep.out_lo = vt.lo;
      ep.ld_r = Sp;
      ep.bs2 = (int64_t)d * Sp;
      launch_gemm(st, A, Bop, d, Sp, d, 1, Bc, 1, ep, 128);
      ++launches;
    }
    if (flash) {
      launch_enc_flash_attn(st, qk, vt, code, tab1d, tabh, tabv, lut_enc1d, lut_enc1d_n, c.rel_buckets / 2, c.rel_buckets,
                            Bc, H, d, Sp, ctx);
      ++launches;
    } else {
      {  // scores[b][h] = q k^T   (no 1/sqrt(d) scaling in T5)
        GemmOperand A, Bop;
        A.hi = qk.hi; A.lo = qk.lo; A.rows = Sp; A.ld = 2 * d;
        A.use_b1 = true; A.bs1 = 64; A.use_b2 = true; A.bs2 = (int64_t)Sp * 2 * d;
        Bop = A;
        Bop.hi = qk.hi + d;
        Bop.lo = qk.lo ? qk.lo + d : nullptr;
        GemmEpilogue ep;
        ep.out_f32 = scores;
        ep.ld_r = Sp;
        ep.bs1 = (int64_t)Sp * Sp;
        ep.bs2 = (int64_t)H * Sp * Sp;
        launch_gemm(st, A, Bop, Sp, Sp, 64, H, Bc, 1, ep, 128);
        ++launches;
      }
      launch_enc_softmax(st, scores, hv, vmask, tab1d, tabh, tabv, lut_enc1d, lut_enc1d_n, c.rel_buckets / 2,
                         c.rel_buckets, Bc, H, Sp, P);
      ++launches;
      {  // ctx[b][:, h*64:(h+1)*64] = P[b][h] . V[b][h]
        GemmOperand A, Bop;
        A.hi = P.hi; A.lo = P.lo; A.rows = Sp; A.ld = Sp;
        A.use_b1 = true; A.bs1 = (int64_t)Sp * Sp; A.use_b2 = true; A.bs2 = (int64_t)H * Sp * Sp;
        Bop.hi = vt.hi; Bop.lo = vt.lo; Bop.rows = 64; Bop.ld = Sp;
        Bop.use_b1 = true; Bop.bs1 = (int64_t)64 * Sp; Bop.use_b2 = true; Bop.bs2 = (int64_t)d * Sp;
        GemmEpilogue ep;
        ep.out_hi = ctx.hi;
        ep.out_lo = ctx.lo;
        ep.ld_r = d;
        ep.bs1 = 64;
        ep.bs2 = (int64_t)Sp * d;
        launch_gemm(st, A, Bop, Sp, 64, Sp, H, Bc, 1, ep, 64);
        ++launches;
      }
    }
    {
      GemmEpilogue ep;
      ep.out_f32 = x;
      ep.residual = x;
      linear(st, L.o, ctx, d, T, ep);
    }
    launch_rmsnorm(st, x, L.ln2, T, d, c.ln_eps, 1.f, xn, nullptr, 0, 0, 0);
    ++launches;
    {
      GemmEpilogue ep;
      ep.out_hi = hid.hi;
      ep.out_lo = hid.lo;
      ep.act = ACT_RELU;
      linear(st, L.wi, xn, d, T, ep);
    }
    {
      GemmEpilogue ep;
      ep.out_f32 = x;
      ep.residual = x;
      linear(st, L.wo, hid, c.d_ff, T, ep);
    }
  }
  // final RMSNorm straight into rows [n_sw, n_sw+Sp) of the encoder memory
  launch_rmsnorm(st, x, enc_final_ln, T, d, c.ln_eps, 1.f, Planes{}, mem + (int64_t)B0 * cur_Mp * d, Sp,
                 (int64_t)cur_Mp * d, (int64_t)n_sw * d);
  ++launches;
  launch_build_mem_mask(st, vmask, Bc, Sp, S, n_sw, cur_Mp, mem_mask + (int64_t)B0 * cur_Mp);
  ++launches;
  (void)S;
}

void mg_model::encode(cudaStream_t st, int B, int Lt, const int64_t* ids, const float* bbox, const float* px,
                      const int64_t* amask) {
  MG_REQUIRE(finalized, "mg_finalize has not been called");
  MG_REQUIRE(B > 0 && Lt > 0, "empty batch");
  const mg_config& c = cfg;
  const int d = c.d_model;
  cur_B = B;
  cur_Lt = Lt;
  cur_S = Lt + NP;
  cur_Sp = (int)rup(cur_S, 8);
  cur_M = n_sw + cur_S;
  cur_Mp = (int)rup(n_sw + cur_Sp, 8);
  persist.reset();
  mem = persist.get<float>((int64_t)B * cur_Mp * d);
  mem_pl = planes(persist, (int64_t)B * cur_Mp * d);
  mem_mask = persist.get<int>((int64_t)B * cur_Mp);
  MG_CHECK_CUDA(cudaMemsetAsync(mem, 0, sizeof(float) * (int64_t)B * cur_Mp * d, st));
  const int chunk = c.enc_chunk > 0 ? c.enc_chunk : 64;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bc = std::min(chunk, B - b0);
    scratch.reset();
    swin_forward(st, b0, bc, px);
    scratch.reset();
    vtl_forward(st, b0, bc, Lt, ids, bbox, px, amask);
  }
  launch_split(st, mem, (int64_t)B * cur_Mp, d, d, mem_pl, d);
  ++launches;
}

// cross K^T / V of every decoder layer, once per generate (UdopAttention :575-583). K^T [B][H][64][Mp] via operand
// swap, V head-major [B][H][Mp][64]: every (image, head) block is one contiguous stream for the decode kernels.
// With ckv24 the fp32 results of a layer are repacked into kv24 blocks (decode.cu: 3 bytes per element) and the two
// fp32 buffers are reused by the next layer.
// Drops the masked memory positions before anything decoder-side touches them (scratch arena; call after its reset).
// One host read of B counters per generate call sizes the compacted length.
void mg_model::prepare_decoder_memory(cudaStream_t st, int B) {
  const int d = cfg.d_model;
  const bool env_off = getenv("MG_COMPACT") && getenv("MG_COMPACT")[0] == '0';  // read per call (A/B, tests)
  dmem = mem; dmem_pl = mem_pl; dmask = mem_mask; dMp = cur_Mp; dM = cur_M;
  if (env_off) return;
  Arena& a = scratch;
  int* src = a.get<int>((int64_t)B * cur_Mp);
  int* nv = a.get<int>(B);
  launch_compact_plan(st, mem_mask, B, cur_Mp, src, nv);
  ++launches;
  std::vector<int> h(B);
  MG_CHECK_CUDA(cudaMemcpyAsync(h.data(), nv, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
  MG_CHECK_CUDA(cudaStreamSynchronize(st));
  int mx = 1;
  for (int v : h) mx = std::max(mx, v);
  const int Mc = (int)rup(mx, getenv("MG_COMPACT_PAD8") ? 8 : 16);  // 16: the fused step's K chunks may then hold an odd number of d-rows
  if (Mc >= cur_Mp) return;  // nothing to drop
  dM = mx;
  dMp = Mc;
  dmem = a.get<float>((int64_t)B * Mc * d);
  dmask = a.get<int>((int64_t)B * Mc);
  dmem_pl = planes(a, (int64_t)B * Mc * d);
  launch_compact_gather(st, mem, src, nv, B, cur_Mp, Mc, d, dmem, dmask);
  launch_split(st, dmem, (int64_t)B * Mc, d, d, dmem_pl, d);
  launches += 2;
}

void mg_model::project_cross_kv(cudaStream_t st, int B, std::vector<float*>& ckt, std::vector<float*>& cv,
                                std::vector<uint8_t*>* ckv24) {
  const mg_config& c = cfg;
  const int d = c.d_model, H = c.num_heads, Mp = dMp, NL = c.num_decoder_layers;
  const Planes mem_pl = dmem_pl;  // (shadows the member: the decoder-side view)
  Arena& a = scratch;
  float *tmp_k = nullptr, *tmp_v = nullptr;
  if (ckv24) {
    tmp_k = a.get<float>((int64_t)B * d * Mp);
    tmp_v = a.get<float>((int64_t)B * Mp * d);
  }
  for (int l = 0; l < NL; ++l) {
    ckt[l] = ckv24 ? tmp_k : a.get<float>((int64_t)B * d * Mp);
    cv[l] = ckv24 ? tmp_v : a.get<float>((int64_t)B * Mp * d);
    {
      GemmOperand A, Bop;
      A.hi = dec[l].ck.w.hi; A.lo = dec[l].ck.w.lo; A.rows = d; A.ld = dec[l].ck.ldk;
      Bop.hi = mem_pl.hi; Bop.lo = mem_pl.lo; Bop.rows = Mp; Bop.ld = d; Bop.use_b2 = true; Bop.bs2 = (int64_t)Mp * d;
      GemmEpilogue ep;
      ep.out_f32 = ckt[l];
      ep.ld_r = Mp;
      ep.bs2 = (int64_t)d * Mp;
      launch_gemm(st, A, Bop, d, Mp, d, 1, B, 1, ep, 128);
      ++launches;
    }
    {  // V, head-major: cv[b][h][m][64] so that every (image, head) block is one contiguous stream
      GemmOperand A, Bop;
      A.hi = mem_pl.hi; A.lo = mem_pl.lo; A.rows = Mp; A.ld = d; A.use_b2 = true; A.bs2 = (int64_t)Mp * d;
      Bop.hi = dec[l].cv.w.hi; Bop.lo = dec[l].cv.w.lo; Bop.rows = 64; Bop.ld = dec[l].cv.ldk;
      Bop.use_b1 = true; Bop.bs1 = (int64_t)64 * dec[l].cv.ldk;
      GemmEpilogue ep;
      ep.out_f32 = cv[l];
      ep.ld_r = 64;
      ep.bs1 = (int64_t)Mp * 64;
      ep.bs2 = (int64_t)Mp * d;
      launch_gemm(st, A, Bop, Mp, 64, d, H, B, 1, ep, 64);
      ++launches;
    }
    if (ckv24) {
      (*ckv24)[l] = a.get<uint8_t>((int64_t)B * H * 384 * Mp);
      launch_kv24_pack(st, ckt[l], cv[l], B, H, Mp, (*ckv24)[l]);
      ++launches;
    }
  }
}

// ================================================================================================= greedy decode
void mg_model::generate(cudaStream_t st, int B, int max_length, int64_t* out_ids, int32_t* out_len,
                        float* step_logits, int32_t* steps_run, const int64_t* forced, int forced_ld) {
  const mg_config& c = cfg;
  MG_REQUIRE(B == cur_B && mem != nullptr, "generate: encode must run first on the same batch");
  MG_REQUIRE(max_length >= 2 && max_length <= 4096, "max_length out of range");
  const int d = c.d_model, H = c.num_heads, V = c.vocab_size, NL = c.num_decoder_layers;
  Arena& a = scratch;
  a.reset();
  prepare_decoder_memory(st, B);
  const int Mp = dMp;
  int* const mem_mask = dmask;  // (shadows the member: the decoder-side view)
  // fused persistent decode step (decode_mega.cu) whenever the batch fits one activation tile
  // cross K/V with 24 significant bits (3 bytes / element) unless MG_KV24=0
  const bool env_kv24 = !(getenv("MG_KV24") && getenv("MG_KV24")[0] == '0');
  const bool kv24 = env_kv24 && Mp % 8 == 0 && Mp <= 2048;
  const bool use_mega = mega_ctas > 0 && kv24 && B <= 32 && max_length <= 512 && NL <= 24;
  const int Tp = (int)rup(max_length, use_mega ? 32 : 4);
  const int64_t Vld = rup(V, 4);

  std::vector<float*> ckt(NL), cv(NL);
  std::vector<uint8_t*> ckv(NL, nullptr);
  project_cross_kv(st, B, ckt, cv, kv24 ? &ckv : nullptr);
  // ---- decode state
  std::vector<float*> skt(NL), sv(NL);
  for (int l = 0; l < NL; ++l) {
    skt[l] = a.get<float>((int64_t)B * d * Tp);
    sv[l] = a.get<float>((int64_t)B * Tp * d);
  }
  float* x = a.get<float>((int64_t)B * d);
  float* qkv = a.get<float>((int64_t)B * 4 * d);  // q | k | v, then (fused path with the folded FF) dx = Wco ctx behind them
  float* q = a.get<float>((int64_t)B * d);
  float* ctx = a.get<float>((int64_t)B * d);
  float* hbuf = a.get<float>((int64_t)B * c.d_ff);
  float* logits = a.get<float>((int64_t)B * Vld);
  float* xalt = a.get<float>((int64_t)B * d);     // second residual buffer of the fused path (folded FF)
  const int n_part = (V + 127) / 128;  // LM-head tiles: per-tile (max, argmax) from the GEMM epilogue
  float* part_val = a.get<float>((int64_t)B * n_part);
  int* part_idx = a.get<int>((int64_t)B * n_part);
  int* finished = a.get<int>(B);
  int* gctr = a.get<int>(8);  // [3] = global n_unfinished (multi-GPU)
  const bool dist = comm != nullptr && dist_all_ids != nullptr && forced == nullptr;
  int* step_tok = a.get<int>(B);
  unsigned long long* step_ts = a.get<unsigned long long>(max_length);  // %globaltimer at the end of every step
  int* gathered = a.get<int>((int64_t)world * B);
  int* gfinished = a.get<int>((int64_t)world * B);
  float* rs_rows = a.get<float>(B);

  // ---- micro-batch lanes (experimental, MG_LANES=2; default 1).  The decode chain of one token is strictly
  // sequential and alternates between latency-bound kernels (skinny linears) and the HBM-bound cross-attention
  // stream; with two lanes the image batch is split in two and the same captured step runs on two streams so one
  // lane can stream K/V while the other runs its linears.  Images are independent, so results are identical.
  // Measured on B200 at batch 32: no gain (2.95 vs 2.93 ms/step) -- the half-batch kernels are as latency-bound
  // as the full-batch ones and contend for shared memory -- so the default stays at one lane.
  static const int env_lanes = getenv("MG_LANES") ? atoi(getenv("MG_LANES")) : 1;
  const int nlanes = (B >= 8 && env_lanes >= 2 && !use_mega) ? 2 : 1;
  // wide batches (> 128 rows, kernel chain, one lane): activation planes of the linears' inputs (MG_WIDE=0: skinny kernels only)
  const bool wide = split2 && nlanes == 1 && !(getenv("MG_WIDE") && getenv("MG_WIDE")[0] == '0');
  Planes wide_xs = (wide && B > 128) ? planes(a, (int64_t)B * std::max(d, c.d_ff)) : Planes{};
  struct Lane {
    int b0, bn;
    cudaStream_t st;
    int* ctr;  // [0]=step [1]=n_unfinished [2]=ticket
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
  };
  Lane lanes[2];
  if (nlanes == 2 && !aux_stream) {
    MG_CHECK_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
    MG_CHECK_CUDA(cudaEventCreateWithFlags(&lane_ev[0], cudaEventDisableTiming));
    MG_CHECK_CUDA(cudaEventCreateWithFlags(&lane_ev[1], cudaEventDisableTiming));
    MG_CHECK_CUDA(cudaEventCreateWithFlags(&lane_ev[2], cudaEventDisableTiming));
  }
  for (int i = 0; i < nlanes; ++i) {
    lanes[i].b0 = i == 0 ? 0 : (B + 1) / 2;
    lanes[i].bn = nlanes == 1 ? B : (i == 0 ? (B + 1) / 2 : B - (B + 1) / 2);
    lanes[i].st = i == 0 ? st : aux_stream;
    lanes[i].ctr = a.get<int>(8);
  }

  // tokens travel through NVLink peer stores issued by the selection kernel itself when every rank mapped every peer
  // at mg_comm_init (one micro-batch lane only); otherwise one ncclAllGather + scatter kernel per step
  const bool p2p = dist && px_ok && B <= px_bcap && nlanes == 1;
  PeerExchange px;
  if (p2p) {
    px.peers = px_peers_dev; px.world = world; px.rank = rank; px.bcap = px_bcap; px.B = B;
    px.base = (px_calls++) * 8192u;
    px.all_ids = dist_all_ids; px.ld = max_length; px.eos = c.eos_token_id;
    px.gfinished = gfinished; px.g_unfinished = gctr + 3; px.consumed = gctr + 4;
    MG_CHECK_CUDA(cudaMemsetAsync(gctr + 4, 0, sizeof(int), st));
  }
  if (dist) {
    MG_CHECK_CUDA(cudaMemsetAsync(gfinished, 0, sizeof(int) * (size_t)world * B, st));
    MG_CHECK_CUDA(cudaMemsetAsync(dist_all_ids, 0, sizeof(int64_t) * (size_t)world * B * max_length, st));
    // column 0 of every row = decoder start id; the global unfinished counter starts at world*B
    std::vector<int64_t> col0((size_t)world * B, (int64_t)c.decoder_start_token_id);
    MG_CHECK_CUDA(cudaMemcpy2DAsync(dist_all_ids, sizeof(int64_t) * max_length, col0.data(), sizeof(int64_t),
                                    sizeof(int64_t), (size_t)world * B, cudaMemcpyHostToDevice, st));
    const int wb = world * B;
    MG_CHECK_CUDA(cudaMemcpyAsync(gctr + 3, &wb, sizeof(int), cudaMemcpyHostToDevice, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));  // col0 / wb are host temporaries
  }
  int64_t* ids_dev = out_ids;
  for (int i = 0; i < nlanes; ++i) {
    Lane& L = lanes[i];
    launch_decode_init(st, shared, d, c.decoder_start_token_id, c.pad_token_id, L.bn, ids_dev + (int64_t)L.b0 * max_length, max_length,
                       finished + L.b0, L.ctr, L.ctr + 1, L.ctr + 2, x + (int64_t)L.b0 * d,
                       forced ? forced + (int64_t)L.b0 * forced_ld : nullptr, forced_ld);
    ++launches;
  }
  // split-K accumulation buffers start at zero; afterwards each is re-zeroed by a later kernel of the chain
  MG_CHECK_CUDA(cudaMemsetAsync(qkv, 0, sizeof(float) * (size_t)B * 4 * d, st));
  MG_CHECK_CUDA(cudaMemsetAsync(xalt, 0, sizeof(float) * (size_t)B * d, st));
  MG_CHECK_CUDA(cudaMemsetAsync(q, 0, sizeof(float) * (size_t)B * d, st));
  MG_CHECK_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(float) * (size_t)B * c.d_ff, st));
  if (nlanes == 2) {  // the second lane starts after the encoder, the K/V projection and the state reset
    MG_CHECK_CUDA(cudaEventRecord(lane_ev[0], st));
    MG_CHECK_CUDA(cudaStreamWaitEvent(aux_stream, lane_ev[0], 0));
  }

  // one decode step of one lane (rows [b0, b0+bn)) on the lane's stream
  MegaParams mp;
  if (use_mega) {
    for (int l = 0; l < NL; ++l) {
      mega_layers[l].skb = skt[l];
      mega_layers[l].svb = sv[l];
      mega_layers[l].ckv = ckv[l];
      mega_layers[l].pad_ = nullptr;
    }
    MG_CHECK_CUDA(cudaMemcpyAsync(mega_layers_dev, mega_layers.data(), sizeof(MegaLayer) * NL, cudaMemcpyHostToDevice, st));
    MG_CHECK_CUDA(cudaMemsetAsync(mega_bar, 0, (4 + 256) * sizeof(unsigned), st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));  // pageable source
    mp.layers = mega_layers_dev; mp.NL = NL; mp.lm_head = mega_lm; mp.final_ln = dec_final_ln;
    mp.logit_scale = c.logit_scale; mp.eps = c.ln_eps;
    mp.B = B; mp.H = H; mp.D = d; mp.DFF = c.d_ff; mp.Mp = Mp; mp.Tp = Tp;
    mp.x = x; mp.qkv = qkv; mp.q = q; mp.ctx = ctx; mp.hbuf = hbuf; mp.logits = logits; mp.ld_logits = (int)Vld;
    mp.xalt = xalt; mp.dx = qkv + (int64_t)B * 3 * d;
    mp.part_val = part_val; mp.part_idx = part_idx; mp.step_ptr = lanes[0].ctr; mp.mem_mask = mem_mask;
    mp.dec_bias = dec_bias; mp.lut = lut_dec; mp.bar_ctr = mega_bar;
    mp.rs = a.get<float>(3 * 32);
    mp.xp = a.get<float>((int64_t)B * H * (Mp + 4));
    mp.xflag = a.get<unsigned>((int64_t)B * H);
    MG_CHECK_CUDA(cudaMemsetAsync(mp.xflag, 0, sizeof(unsigned) * (size_t)B * H, st));
    mp.dbg_host = pinned_flag + 8;
    for (int i = 8; i < 16; ++i) pinned_flag[i] = 0;
    if (getenv("MG_MEGA_DBG")) mp.dbg = atoi(getenv("MG_MEGA_DBG"));
    if (getenv("MG_MEGA_GATE")) mp.gate = atoi(getenv("MG_MEGA_GATE"));
    if (getenv("MG_MEGA_INFLIGHT")) mp.max_inflight = std::max(1, std::min(5, atoi(getenv("MG_MEGA_INFLIGHT"))));
    if (getenv("MG_MEGA_CROSS_RK")) mp.cross_rk = atoi(getenv("MG_MEGA_CROSS_RK"));
    if (getenv("MG_MEGA_CROSS_VR")) mp.cross_vr = atoi(getenv("MG_MEGA_CROSS_VR"));
    if (getenv("MG_MEGA_L2PF")) mp.l2pf = std::max(0, atoi(getenv("MG_MEGA_L2PF"))) * 1024;  // KB per CTA and layer
    if (getenv("MG_MEGA_L2PF_PIECE")) mp.l2pf_piece = std::max(0, atoi(getenv("MG_MEGA_L2PF_PIECE")) & ~15);
    if (getenv("MG_MEGA_L2PF_GAP")) mp.l2pf_gap = std::max(0, atoi(getenv("MG_MEGA_L2PF_GAP")));
    if (getenv("MG_MEGA_L2PF_MASK")) mp.l2pf_mask = (int)strtol(getenv("MG_MEGA_L2PF_MASK"), nullptr, 0);
    if (getenv("MG_MEGA_PROF")) {  // debug: per-phase timestamps of the last step, dumped after the loop
      mp.prof = a.get<unsigned long long>((int64_t)mega_ctas * 1024);
      MG_CHECK_CUDA(cudaMemsetAsync(mp.prof, 0, sizeof(unsigned long long) * (size_t)mega_ctas * 1024, st));
    }
  }
  auto one_step = [&](Lane& Ln) {
    cudaStream_t ls = Ln.st;
    const int b0 = Ln.b0, bn = Ln.bn;
    if (use_mega) {
      launch_decode_step(ls, mp, run_ctas > 0 ? std::min(mega_ctas, run_ctas) : mega_ctas);
      launch_greedy_select(ls, part_val, part_idx, n_part, logits, bn, V, Vld, shared, d, c.eos_token_id, c.pad_token_id,
                           ids_dev, max_length, finished, Ln.ctr, Ln.ctr + 1, Ln.ctr + 2, x,
                           step_logits, (int64_t)(max_length - 1) * V, V, forced, forced_ld, step_tok, step_ts,
                           p2p ? &px : nullptr);
      launches += 2;
      return;
    }
    float* const rs_scratch = rs_rows;  // RMSNorm row scales of the wide (> 32 rows) launches
    auto lin = [&](int pro, const float* xin, int ldx, const LinearW& W, float* out, int ld_out, const float* lnw,
                   float scale, float* zp, int64_t zn, bool store, bool amax = false) {
      if (wide && bn > 128 && !store) {
        launches += wide_linear(ls, pro, xin + (int64_t)b0 * ldx, ldx, W, out + (int64_t)b0 * ld_out, ld_out, lnw, scale, bn,
                                wide_xs, zp, zn);
        return;
      }
      for (int r0 = 0; r0 < bn; r0 += 128) {
        const int bc = std::min(128, bn - r0);
        launch_skinny_tc(ls, pro, xin + (int64_t)(b0 + r0) * ldx, ldx, W.w, W.ldk, out + (int64_t)(b0 + r0) * ld_out,
                         ld_out, bc, W.N, W.K, lnw, c.ln_eps, scale, r0 == 0 ? zp : nullptr, zn, store,
                         amax ? part_val + (int64_t)(b0 + r0) * n_part : nullptr,
                         amax ? part_idx + (int64_t)(b0 + r0) * n_part : nullptr, rs_scratch + b0 + r0);
        launches += (pro == 1 && bc > 32) ? 2 : 1;
      }
    };
    for (int l = 0; l < NL; ++l) {
      DecLayer& L = dec[l];
      // self-attention block: RMSNorm fused into the QKV projection; zero duty: FF hidden buffer
      lin(1, x, d, L.qkv, qkv, 3 * d, L.ln1, 1.f, hbuf + (int64_t)b0 * c.d_ff, (int64_t)bn * c.d_ff, false);
      launch_dec_self_attn(ls, qkv + (int64_t)b0 * 3 * d, bn, H, d, skt[l] + (int64_t)b0 * d * Tp, Tp, (int64_t)d * Tp,
                           sv[l] + (int64_t)b0 * Tp * d, d, (int64_t)Tp * d, Ln.ctr, Tp, dec_bias, lut_dec,
                           ctx + (int64_t)b0 * d);
      lin(0, ctx, d, L.o, x, d, nullptr, 1.f, nullptr, 0, false);  // x += o(ctx)
      // cross-attention block; zero duty: the QKV buffer just consumed by self-attention
      lin(1, x, d, L.cq, q, d, L.ln2, 1.f, qkv + (int64_t)b0 * 3 * d, (int64_t)bn * 3 * d, false);
      if (kv24)
        launch_cross_attn_stream24(ls, q + (int64_t)b0 * d, bn, H, d, ckv[l] + (int64_t)b0 * H * 384 * Mp, Mp,
                                   mem_mask + (int64_t)b0 * Mp, ctx + (int64_t)b0 * d);
      else
        launch_cross_attn_stream(ls, q + (int64_t)b0 * d, bn, H, d, ckt[l] + (int64_t)b0 * d * Mp,
                                 cv[l] + (int64_t)b0 * Mp * d, Mp, mem_mask + (int64_t)b0 * Mp, ctx + (int64_t)b0 * d);
      lin(0, ctx, d, L.co, x, d, nullptr, 1.f, nullptr, 0, false);
      // feed-forward: RMSNorm fused into wi, ReLU fused into wo's operand load; zero duty: cross-attention q
      lin(1, x, d, L.wi, hbuf, c.d_ff, L.ln3, 1.f, q + (int64_t)b0 * d, (int64_t)bn * d, false);
      lin(2, hbuf, c.d_ff, L.wo, x, d, nullptr, 1.f, nullptr, 0, false);
      launches += 2;
    }
    // final RMSNorm * d_model^-0.5 fused into the LM head (modeling_udop.py:1585-1590), direct store
    lin(1, x, d, lm_head, logits, (int)Vld, dec_final_ln, c.logit_scale, nullptr, 0, true, true);
    launch_greedy_select(ls, part_val + (int64_t)b0 * n_part, part_idx + (int64_t)b0 * n_part, n_part,
                         logits + (int64_t)b0 * Vld, bn, V, Vld, shared, d, c.eos_token_id, c.pad_token_id,
                         ids_dev + (int64_t)b0 * max_length, max_length, finished + b0, Ln.ctr, Ln.ctr + 1, Ln.ctr + 2,
                         x + (int64_t)b0 * d,
                         step_logits ? step_logits + (int64_t)b0 * (max_length - 1) * V : nullptr,
                         (int64_t)(max_length - 1) * V, V, forced ? forced + (int64_t)b0 * forced_ld : nullptr,
                         forced_ld, step_tok + b0, b0 == 0 ? step_ts : nullptr, p2p ? &px : nullptr);
    launches += 1;
  };
  // one step of every lane; in multi-GPU mode followed by the exchange of the step's token ids
  auto exchange = [&](int col) {
    if (p2p) return;  // already done by the step's selection kernel
    if (nlanes == 2) {  // the all-gather reads both lanes' tokens ...
      MG_CHECK_CUDA(cudaEventRecord(lane_ev[1], aux_stream));
      MG_CHECK_CUDA(cudaStreamWaitEvent(st, lane_ev[1], 0));
    }
    MG_CHECK_NCCL(nccl_api().AllGather(step_tok, gathered, (size_t)B, ncclInt32, comm, st));
    launch_scatter_step(st, gathered, world * B, col, max_length, c.eos_token_id, dist_all_ids, gfinished, gctr + 3);
    launches += 2;
    if (nlanes == 2) {  // ... and lane 1 must not overwrite them before it has run
      MG_CHECK_CUDA(cudaEventRecord(lane_ev[2], st));
      MG_CHECK_CUDA(cudaStreamWaitEvent(aux_stream, lane_ev[2], 0));
    }
  };

  prof_ckt = ckt;
  prof_cv = cv;
  prof_ckv = ckv;
  prof_kv24 = kv24;
  prof_q = q;
  prof_ctx = ctx;
  prof_bn = lanes[0].bn;
  const int total_steps = max_length - 1;
  int done_steps = 0;
  MG_CHECK_CUDA(cudaEventRecord(ev_loop[0], st));
  // step 0 runs eagerly (lazy one-time initialisation happens outside graph capture) ...
  for (int i = 0; i < nlanes; ++i) one_step(lanes[i]);
  done_steps = 1;
  if (dist) exchange(1);
  if (total_steps > 1 && use_mega) {
    // two launches per token: nothing to capture, the host simply runs ahead of the GPU
    pinned_flag[0] = pinned_flag[1] = B;
    bool stop = false;
    try {
    while (done_steps < total_steps && !stop) {
      const int n = std::min(16, total_steps - done_steps);
      for (int i = 0; i < n; ++i) {
        one_step(lanes[0]);
        if (dist) exchange(done_steps + i + 1);
      }
      done_steps += n;
      MG_CHECK_CUDA(cudaEventSynchronize(ev[3]));  // poll the "all finished" counter one window late
      if (pinned_flag[0] == 0) stop = true;
      MG_CHECK_CUDA(cudaMemcpyAsync(pinned_flag, dist ? gctr + 3 : lanes[0].ctr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      MG_CHECK_CUDA(cudaEventRecord(ev[3], st));
    }
    MG_CHECK_CUDA(cudaStreamSynchronize(st));
    } catch (const Error& e) {
      if (pinned_flag[8] != 0)  // the fused kernel's watchdog fired: say where
        throw Error(e.code, std::string(e.what()) + " [decode_step_kernel watchdog: code " + std::to_string(pinned_flag[8]) +
                                " cta " + std::to_string(pinned_flag[9]) + " thread " + std::to_string(pinned_flag[10]) + " a " +
                                std::to_string(pinned_flag[11]) + " b " + std::to_string(pinned_flag[12]) + "]");
      throw;
    }
  } else if (total_steps > 1) {
    // ... then one step per lane is captured and replayed; every kernel reads the step index from device memory
    const int64_t before = launches;
    for (int i = 0; i < nlanes; ++i) {
      MG_CHECK_CUDA(cudaStreamBeginCapture(lanes[i].st, cudaStreamCaptureModeThreadLocal));
      try {
        one_step(lanes[i]);
      } catch (...) {
        cudaGraph_t g2;
        cudaStreamEndCapture(lanes[i].st, &g2);
        throw;
      }
      MG_CHECK_CUDA(cudaStreamEndCapture(lanes[i].st, &lanes[i].graph));
      MG_CHECK_CUDA(cudaGraphInstantiate(&lanes[i].exec, lanes[i].graph, 0));
    }
    if (getenv("MG_DUMP_GRAPH")) cudaGraphDebugDotPrint(lanes[0].graph, getenv("MG_DUMP_GRAPH"), 0);
    const int64_t step_launches = launches - before;
    launches = before;
    pinned_flag[0] = pinned_flag[1] = B;
    const int check_every = 16;
    bool stop = false;
    while (done_steps < total_steps && !stop) {
      const int n = std::min(check_every, total_steps - done_steps);
      for (int i = 0; i < n; ++i) {
        for (int k = 0; k < nlanes; ++k) MG_CHECK_CUDA(cudaGraphLaunch(lanes[k].exec, lanes[k].st));
        if (dist) exchange(done_steps + i + 1);
      }
      launches += step_launches * n;
      done_steps += n;
      // poll the "all finished" counters one window late so the GPU never drains
      MG_CHECK_CUDA(cudaEventSynchronize(ev[3]));
      if (nlanes == 2) MG_CHECK_CUDA(cudaEventSynchronize(ev[5]));
      if (dist ? pinned_flag[0] == 0 : (pinned_flag[0] + (nlanes == 2 ? pinned_flag[1] : 0)) == 0) stop = true;
      MG_CHECK_CUDA(cudaMemcpyAsync(pinned_flag, dist ? gctr + 3 : lanes[0].ctr + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      MG_CHECK_CUDA(cudaEventRecord(ev[3], st));
      if (nlanes == 2) {
        MG_CHECK_CUDA(cudaMemcpyAsync(pinned_flag + 1, lanes[1].ctr + 1, sizeof(int), cudaMemcpyDeviceToHost, aux_stream));
        MG_CHECK_CUDA(cudaEventRecord(ev[5], aux_stream));
      }
    }
  }
  if (nlanes == 2) {  // join the second lane back into the caller's stream
    MG_CHECK_CUDA(cudaEventRecord(lane_ev[1], aux_stream));
    MG_CHECK_CUDA(cudaStreamWaitEvent(st, lane_ev[1], 0));
  }
  if (p2p) {  // the last step's tokens of all ranks
    launch_peer_drain(st, px, done_steps - 1);
    ++launches;
  }
  MG_CHECK_CUDA(cudaEventRecord(ev_loop[1], st));
  if (out_len) {
    launch_out_len(st, ids_dev, B, max_length, std::min(done_steps + 1, max_length), c.eos_token_id, out_len);
    ++launches;
  }
  MG_CHECK_CUDA(cudaStreamSynchronize(st));
  MG_CHECK_CUDA(cudaEventElapsedTime(&last_loop_ms, ev_loop[0], ev_loop[1]));
  last_loop_steps = done_steps;
  last_fused = use_mega ? 1 : 0;
  {  // true per-step latencies: differences of the device-side end-of-step stamps (greedy_select_kernel)
    std::vector<unsigned long long> ts((size_t)done_steps);
    MG_CHECK_CUDA(cudaMemcpy(ts.data(), step_ts, sizeof(unsigned long long) * ts.size(), cudaMemcpyDeviceToHost));
    std::vector<float> per;
    for (size_t i = 1; i < ts.size(); ++i) per.push_back((float)((double)(ts[i] - ts[i - 1]) * 1e-6));
    std::sort(per.begin(), per.end());
    last_step_p50_ms = per.empty() ? 0.f : per[per.size() / 2];
    last_step_p99_ms = per.empty() ? 0.f : per[std::min(per.size() - 1, (size_t)(per.size() * 0.99))];
  }
  if (use_mega && mp.prof) {
    std::vector<unsigned long long> h((size_t)mega_ctas * 1024);
    MG_CHECK_CUDA(cudaMemcpy(h.data(), mp.prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(getenv("MG_MEGA_PROF"), "wb")) {
      fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
      fclose(f);
    }
  }
  for (int i = 0; i < nlanes; ++i) {
    if (lanes[i].exec) cudaGraphExecDestroy(lanes[i].exec);
    if (lanes[i].graph) cudaGraphDestroy(lanes[i].graph);
  }
  if (steps_run) *steps_run = done_steps;
}

// ================================================================================================= beam search
void mg_model::generate_beam(cudaStream_t st, int B, int nb, int max_length, int64_t* out_ids, int32_t* out_len,
                             int32_t* steps_run) {
  const mg_config& c = cfg;
  MG_REQUIRE(B == cur_B && mem != nullptr, "generate: encode must run first on the same batch");
  MG_REQUIRE(max_length >= 2 && max_length <= 4096, "max_length out of range");
  MG_REQUIRE(nb >= 2 && nb <= 8, "2 <= num_beams <= 8");
  const int d = c.d_model, H = c.num_heads, V = c.vocab_size, NL = c.num_decoder_layers;
  const int R = B * nb;
  Arena& a = scratch;
  a.reset();
  prepare_decoder_memory(st, B);
  const int Mp = dMp;
  int* const mem_mask = dmask;  // (shadows the member: the decoder-side view)
  const int Tp = (int)rup(max_length, 4);
  const int64_t Vld = rup(V, 4);
  // cross K/V as kv24 blocks (3 bytes / element, decode.cu) streamed once per IMAGE for all its beams, unless MG_KV24=0
  const bool env_kv24 = !(getenv("MG_KV24") && getenv("MG_KV24")[0] == '0');
  const bool kv24 = env_kv24 && Mp % 8 == 0 && Mp <= 2048;
  std::vector<float*> ckt(NL), cv(NL);
  std::vector<uint8_t*> ckv(NL, nullptr);
  project_cross_kv(st, B, ckt, cv, kv24 ? &ckv : nullptr);
  const bool dist = comm != nullptr && dist_all_ids != nullptr;
  std::vector<float*> skt(NL), sv(NL);
  for (int l = 0; l < NL; ++l) {
    skt[l] = a.get<float>((int64_t)R * d * Tp);
    sv[l] = a.get<float>((int64_t)R * Tp * d);
  }
  float* x = a.get<float>((int64_t)R * d);
  float* qkv = a.get<float>((int64_t)R * 3 * d);
  float* q = a.get<float>((int64_t)R * d);
  float* ctx = a.get<float>((int64_t)R * d);
  float* hbuf = a.get<float>((int64_t)R * c.d_ff);
  float* logits = a.get<float>((int64_t)R * Vld);
  BeamState bs;
  bs.B = B; bs.nb = nb; bs.L = max_length; bs.anc_ld = Tp;
  bs.run_seq = a.get<int64_t>((int64_t)R * max_length);
  bs.fin_seq = a.get<int64_t>((int64_t)R * max_length);
  bs.tmp_seq = a.get<int64_t>((int64_t)R * max_length);
  bs.run_score = a.get<float>(R);
  bs.fin_score = a.get<float>(R);
  bs.fin_flag = a.get<int>(R);
  bs.fin_len = a.get<int>(R);
  bs.unsat = a.get<int>(B);
  bs.anc0 = a.get<int>((int64_t)R * Tp);
  bs.anc1 = a.get<int>((int64_t)R * Tp);
  bs.ctrl = a.get<int>(8);
  int* gathered = nullptr;
  int* gctr = a.get<int>(4);  // [0] = ranks whose search is not done (multi-GPU)
  if (dist) {
    bs.step_tok = a.get<int>(B + 1);
    gathered = a.get<int>((int64_t)world * (B + 1));
    MG_CHECK_CUDA(cudaMemsetAsync(bs.step_tok, 0, sizeof(int) * (size_t)(B + 1), st));
    MG_CHECK_CUDA(cudaMemcpyAsync(gctr, &world, sizeof(int), cudaMemcpyHostToDevice, st));
    MG_CHECK_CUDA(cudaStreamSynchronize(st));  // `world` is read by the copy
  }
  // fill value of unfinished tail positions: stock transformers computes `pad_token_id or eos_token_id[0]`
  // (generation/utils.py:3165), i.e. EOS when the pad id is 0 as it is for UDOP/T5 -- mirrored for id parity
  const int fill = c.pad_token_id != 0 ? c.pad_token_id : c.eos_token_id;
  launch_beam_init(st, bs, shared, d, c.decoder_start_token_id, fill, x);
  ++launches;
  MG_CHECK_CUDA(cudaMemsetAsync(qkv, 0, sizeof(float) * (size_t)R * 3 * d, st));
  MG_CHECK_CUDA(cudaMemsetAsync(q, 0, sizeof(float) * (size_t)R * d, st));
  MG_CHECK_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(float) * (size_t)R * c.d_ff, st));

  float* rs_rows = a.get<float>(R);  // RMSNorm row scales of the wide (> 32 rows) launches
  const bool wide = split2 && !(getenv("MG_WIDE") && getenv("MG_WIDE")[0] == '0');
  Planes wide_xs = (wide && R > 128) ? planes(a, (int64_t)R * std::max(d, c.d_ff)) : Planes{};
  auto lin = [&](int pro, const float* xin, int ldx, const LinearW& W, float* out, int ld_out, const float* lnw,
                 float scale, float* zp, int64_t zn, bool store) {
    if (wide && R > 128 && !store) {
      launches += wide_linear(st, pro, xin, ldx, W, out, ld_out, lnw, scale, R, wide_xs, zp, zn);
      return;
    }
    for (int b0 = 0; b0 < R; b0 += 128) {
      const int bc = std::min(128, R - b0);
      launch_skinny_tc(st, pro, xin + (int64_t)b0 * ldx, ldx, W.w, W.ldk, out + (int64_t)b0 * ld_out, ld_out, bc, W.N,
                       W.K, lnw, c.ln_eps, scale, b0 == 0 ? zp : nullptr, zn, store, nullptr, nullptr, rs_rows + b0);
      launches += (pro == 1 && bc > 32) ? 2 : 1;
    }
  };
  auto one_step = [&]() {
    for (int l = 0; l < NL; ++l) {
      DecLayer& L = dec[l];
      lin(1, x, d, L.qkv, qkv, 3 * d, L.ln1, 1.f, hbuf, (int64_t)R * c.d_ff, false);
      launch_beam_self_attn(st, qkv, R, H, d, skt[l], Tp, (int64_t)d * Tp, sv[l], d, (int64_t)Tp * d, bs.ctrl,
                            bs.ctrl + 1, bs.anc0, bs.anc1, Tp, dec_bias, lut_dec, ctx);
      lin(0, ctx, d, L.o, x, d, nullptr, 1.f, nullptr, 0, false);
      lin(1, x, d, L.cq, q, d, L.ln2, 1.f, qkv, (int64_t)R * 3 * d, false);
      if (kv24)
        launch_beam_cross_attn24(st, q, B, nb, H, d, ckv[l], Mp, mem_mask, ctx);
      else
        launch_beam_cross_attn(st, q, B, nb, H, d, ckt[l], cv[l], Mp, mem_mask, ctx);
      lin(0, ctx, d, L.co, x, d, nullptr, 1.f, nullptr, 0, false);
      lin(1, x, d, L.wi, hbuf, c.d_ff, L.ln3, 1.f, q, (int64_t)R * d, false);
      lin(2, hbuf, c.d_ff, L.wo, x, d, nullptr, 1.f, nullptr, 0, false);
      launches += 2;
    }
    lin(1, x, d, lm_head, logits, (int)Vld, dec_final_ln, c.logit_scale, nullptr, 0, true);
    launch_beam_select(st, bs, logits, V, Vld, shared, d, c.eos_token_id, max_length, x);
    launches += 1;
  };
  // multi-GPU: one all-gather per step of {token of every image's best running beam, this rank's done flag}; every rank
  // keeps stepping (frozen once its own search is done) until ALL ranks are done, so all stop on the same step
  auto exchange = [&](int col) {
    MG_CHECK_NCCL(nccl_api().AllGather(bs.step_tok, gathered, (size_t)(B + 1), ncclInt32, comm, st));
    launch_beam_scatter_step(st, gathered, world, B, col, max_length, dist_all_ids, gctr);
    launches += 2;
  };
  if (dist) MG_CHECK_CUDA(cudaMemsetAsync(dist_all_ids, 0, sizeof(int64_t) * (size_t)world * B * max_length, st));
  const int total_steps = max_length - 1;
  MG_CHECK_CUDA(cudaEventRecord(ev_loop[0], st));
  one_step();
  int done_steps = 1;
  if (dist) exchange(1);
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  if (total_steps > 1) {
    const int64_t before = launches;
    MG_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      one_step();
    } catch (...) {
      cudaGraph_t g2;
      cudaStreamEndCapture(st, &g2);
      throw;
    }
    MG_CHECK_CUDA(cudaStreamEndCapture(st, &graph));
    MG_CHECK_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
    const int64_t step_launches = launches - before;
    launches = before;
    pinned_flag[0] = dist ? world : 0;  // device-side stop rule: "done" flag, or (multi-GPU) ranks not yet done
    bool stop = false;
    while (done_steps < total_steps && !stop) {
      const int n = std::min(16, total_steps - done_steps);
      for (int i = 0; i < n; ++i) {
        MG_CHECK_CUDA(cudaGraphLaunch(gexec, st));
        if (dist) exchange(done_steps + i + 1);
      }
      launches += step_launches * n;
      done_steps += n;
      MG_CHECK_CUDA(cudaEventSynchronize(ev[3]));
      if (dist ? pinned_flag[0] == 0 : pinned_flag[0] != 0) stop = true;
      MG_CHECK_CUDA(cudaMemcpyAsync(pinned_flag, dist ? gctr : bs.ctrl + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
      MG_CHECK_CUDA(cudaEventRecord(ev[3], st));
    }
  }
  launch_beam_finalize(st, bs, fill, out_ids, out_len);
  ++launches;
  MG_CHECK_CUDA(cudaEventRecord(ev_loop[1], st));
  if (dist) {  // the final sequences of every rank's images, in global image order
    MG_CHECK_NCCL(nccl_api().AllGather(out_ids, dist_all_ids, (size_t)B * max_length, ncclInt64, comm, st));
    ++launches;
  }
  MG_CHECK_CUDA(cudaStreamSynchronize(st));
  MG_CHECK_CUDA(cudaEventElapsedTime(&last_loop_ms, ev_loop[0], ev_loop[1]));
  last_loop_steps = done_steps;
  last_fused = 0;
  last_step_p50_ms = last_step_p99_ms = 0.f;
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  if (steps_run) *steps_run = done_steps;
}

// ================================================================================================= C ABI
// (a failed launch of an EARLIER call leaves a non-sticky error behind that the next cudaGetLastError() would report
// against an innocent kernel: every entry point starts from a clean slate)
#define MG_API_BEGIN \
  try {              \
    cudaGetLastError();
#define MG_API_END                                        \
  return 0;                                               \
  }                                                       \
  catch (const mg::Error& e) { return mg::set_error(e); } \
  catch (const std::exception& e) { return mg::set_error(e); }

// ================================================================================================= encoder run-ahead
// Splits the SMs into a small group for the run-ahead encoder and the rest for the decode loop.  MG_AHEAD=0 switches the
// feature off, MG_AHEAD_SMS sizes the small group (default 16: the fused decode step keeps 132 CTAs, and every linear
// of the step has 128 work items, so only its cross-attention and LM-head phases lose SMs).
bool mg_model::ensure_partitions() {
  if (part.tried) return part.ok;
  part.tried = true;
  if (getenv("MG_AHEAD") && getenv("MG_AHEAD")[0] == '0') return false;
  const int want = getenv("MG_AHEAD_SMS") ? std::max(8, atoi(getenv("MG_AHEAD_SMS"))) : 16;
  auto entry = [](const char* name) -> void* {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return fn;
  };
  auto pDevGet = reinterpret_cast<CUresult (*)(CUdevice*, int)>(entry("cuDeviceGet"));
  auto pGetRes = reinterpret_cast<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>(entry("cuDeviceGetDevResource"));
  auto pSplit = reinterpret_cast<CUresult (*)(CUdevResource*, unsigned*, const CUdevResource*, CUdevResource*, unsigned, unsigned)>(
      entry("cuDevSmResourceSplitByCount"));
  auto pDesc = reinterpret_cast<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned)>(entry("cuDevResourceGenerateDesc"));
  auto pCreate = reinterpret_cast<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned)>(entry("cuGreenCtxCreate"));
  auto pStream = reinterpret_cast<CUresult (*)(CUstream*, CUgreenCtx, unsigned, int)>(entry("cuGreenCtxStreamCreate"));
  if (!pDevGet || !pGetRes || !pSplit || !pDesc || !pCreate || !pStream) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  CUdevice cudev;
  CUdevResource all, grp, rest;
  unsigned n = 1;
  CUdevResourceDesc d_small, d_big;
  CUstream ss = nullptr, sb = nullptr;
  if (pDevGet(&cudev, dev) != CUDA_SUCCESS || pGetRes(cudev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  if ((int)all.sm.smCount < want + 64) return false;
  if (pSplit(&grp, &n, &all, &rest, 0, (unsigned)want) != CUDA_SUCCESS || n < 1 || rest.sm.smCount < 64) return false;
  if (pDesc(&d_small, &grp, 1) != CUDA_SUCCESS || pDesc(&d_big, &rest, 1) != CUDA_SUCCESS) return false;
  if (pCreate(&part.g_small, d_small, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  if (pCreate(&part.g_big, d_big, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
  if (pStream(&ss, part.g_small, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
  if (pStream(&sb, part.g_big, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
  part.s_small = reinterpret_cast<cudaStream_t>(ss);
  part.s_big = reinterpret_cast<cudaStream_t>(sb);
  part.n_small = (int)grp.sm.smCount;
  part.n_big = (int)rest.sm.smCount;
  if (cudaEventCreateWithFlags(&part.order, cudaEventDisableTiming) != cudaSuccess) return false;
  part.ok = true;
  return true;
}

// Runs the encoder of a batch on the small partition, ordered after everything queued on `after` (the inputs may
// still be in flight there), into the free slot; returns once the work is queued.
void mg_model::encode_ahead(cudaStream_t after, int B, int Lt, const int64_t* ids, const float* bbox, const float* px,
                            const int64_t* amask) {
  MG_REQUIRE(finalized, "mg_finalize has not been called");
  MG_REQUIRE(part.ok, "encode_ahead without SM partitions");
  int k = 0;  // a slot without a pending batch, else the older one (its batch is dropped: it was never asked for)
  if (ahead[0].valid && (!ahead[1].valid || ahead[1].seq < ahead[0].seq)) k = 1;
  EncSlot& S = ahead[k];
  if (!S.done) {
    MG_CHECK_CUDA(cudaEventCreate(&S.done));
    MG_CHECK_CUDA(cudaEventCreate(&S.begin));
  }
  S.valid = false;
  MG_CHECK_CUDA(cudaEventRecord(part.order, after));
  MG_CHECK_CUDA(cudaStreamWaitEvent(part.s_small, part.order, 0));
  MG_CHECK_CUDA(cudaEventRecord(S.begin, part.s_small));
  const int64_t l0 = launches;
  swap_view(S);
  try {
    encode(part.s_small, B, Lt, ids, bbox, px, amask);
  } catch (...) {
    swap_view(S);
    throw;
  }
  swap_view(S);  // the slot now holds the encoded batch, the model its previous view
  MG_CHECK_CUDA(cudaEventRecord(S.done, part.s_small));
  S.k_ids = ids; S.k_box = bbox; S.k_px = px; S.k_mask = amask;
  S.valid = true;
  S.seq = ++ahead_seq;
  S.n_launch = launches - l0;
}

extern "C" {

int mg_device_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n > 0 ? 1 : 0;
}

int mg_rel_bucket_lut(int bidirectional, int num_buckets, int max_distance, int n_entries, int32_t* lut_host) {
  MG_API_BEGIN
  MG_REQUIRE(lut_host && n_entries > 0 && num_buckets >= 4, "bad arguments");
  rel_bucket_lut(bidirectional, num_buckets, max_distance, n_entries, lut_host);
  MG_API_END
}

int mg_create(const mg_config* cfg, mg_model** out) {
  MG_API_BEGIN
  MG_REQUIRE(cfg && out, "null argument");
  mg_model* m = new mg_model();
  m->cfg = *cfg;
  *out = m;
  MG_API_END
}

void mg_destroy(mg_model* m) { delete m; }

int mg_load_weight(mg_model* m, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int rank) {
  try {
    MG_REQUIRE(m && name && dev_ptr && rank >= 0 && rank <= 8, "bad arguments");
    MG_REQUIRE(dtype == 0, "only fp32 weights are accepted");
    MG_REQUIRE(!m->finalized, "model already finalized");
    RawWeight w;
    w.ptr = static_cast<const float*>(dev_ptr);
    w.shape.assign(shape, shape + rank);
    m->raw[name] = w;
    return 0;
  } catch (const mg::Error& e) {
    return mg::set_error(e);
  } catch (const std::exception& e) {
    return mg::set_error(e);
  }
}

int mg_finalize(mg_model* m, void* stream) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  m->finalize(static_cast<cudaStream_t>(stream));
  MG_API_END
}

int mg_encode(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
              const float* pixel_values, const int64_t* attn_mask, float* enc_out, int32_t* enc_mask, int32_t* M_out) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  m->encode(st, B, Lt, input_ids, bbox, pixel_values, attn_mask);
  const int d = m->cfg.d_model, M = m->cur_M, Mp = m->cur_Mp;
  if (enc_out)
    MG_CHECK_CUDA(cudaMemcpy2DAsync(enc_out, sizeof(float) * (size_t)M * d, m->mem, sizeof(float) * (size_t)Mp * d,
                                    sizeof(float) * (size_t)M * d, B, cudaMemcpyDeviceToDevice, st));
  if (enc_mask)
    MG_CHECK_CUDA(cudaMemcpy2DAsync(enc_mask, sizeof(int) * (size_t)M, m->mem_mask, sizeof(int) * (size_t)Mp,
                                    sizeof(int) * (size_t)M, B, cudaMemcpyDeviceToDevice, st));
  if (M_out) *M_out = M;
  MG_API_END
}

int mg_generate(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                const float* pixel_values, const int64_t* attn_mask, int num_beams, int max_length, int64_t* out_ids,
                int32_t* out_len, float* step_logits, int32_t* steps_run) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values && out_ids, "null argument");
  MG_REQUIRE(num_beams >= 1 && num_beams <= 8, "1 <= num_beams <= 8");
  MG_REQUIRE(num_beams == 1 || step_logits == nullptr, "step_logits is only available for greedy decoding");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (st == nullptr || st == cudaStreamLegacy) {
    // the decode step is replayed from a captured CUDA graph and the legacy default stream cannot be
    // captured: run on the model's own stream, ordered after everything already queued on the device
    MG_CHECK_CUDA(cudaDeviceSynchronize());
    if (!m->own_stream) MG_CHECK_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    st = m->own_stream;
  }
  const int64_t l0 = m->launches;
  // a batch whose encoder ran ahead (mg_encode_ahead) is taken from its slot; while ANOTHER batch is being encoded ahead
  // this call runs on the large SM partition, so that the two never compete for an SM
  const int slot = m->find_ahead(B, Lt, input_ids, bbox, pixel_values, attn_mask);
  bool other_pending = false;  // ... and its encoder has not finished yet (a finished one needs no SMs any more)
  for (int i = 0; i < 2; ++i) {
    if (i == slot || !m->ahead[i].valid) continue;
    if (cudaEventQuery(m->ahead[i].done) != cudaSuccess) other_pending = true;
    cudaGetLastError();  // (cudaErrorNotReady is an answer, not an error)
  }
  m->run_ctas = 0;
  // NCCL collectives inside the decode loop (beam search across ranks, MG_DIST=nccl) stay on the caller's stream: let the
  // run-ahead encoder finish first instead of partitioning
  const bool nccl_in_loop = m->comm != nullptr && m->dist_all_ids != nullptr && (num_beams > 1 || !m->px_ok);
  if (m->part.ok && other_pending && nccl_in_loop) MG_CHECK_CUDA(cudaStreamSynchronize(m->part.s_small));
  if (m->part.ok && other_pending && !nccl_in_loop) {
    MG_CHECK_CUDA(cudaEventRecord(m->part.order, st));
    MG_CHECK_CUDA(cudaStreamWaitEvent(m->part.s_big, m->part.order, 0));
    st = m->part.s_big;
    m->run_ctas = m->part.n_big;
  }
  MG_CHECK_CUDA(cudaEventRecord(m->ev[0], st));
  int took_slot = -1;
  if (slot >= 0) {
    mg_model::EncSlot& S = m->ahead[slot];
    m->swap_view(S);
    S.valid = false;
    MG_CHECK_CUDA(cudaStreamWaitEvent(st, S.done, 0));
    took_slot = slot;
  } else {
    m->encode(st, B, Lt, input_ids, bbox, pixel_values, attn_mask);
  }
  MG_CHECK_CUDA(cudaEventRecord(m->ev[1], st));
  MG_CHECK_CUDA(cudaEventRecord(m->ev[3], st));
  MG_CHECK_CUDA(cudaEventRecord(m->ev[5], st));
  if (num_beams == 1)
    m->generate(st, B, max_length, out_ids, out_len, step_logits, steps_run);
  else
    m->generate_beam(st, B, num_beams, max_length, out_ids, out_len, steps_run);
  m->run_ctas = 0;
  MG_CHECK_CUDA(cudaEventRecord(m->ev[2], st));
  MG_CHECK_CUDA(cudaEventSynchronize(m->ev[2]));  // (also orders the caller's stream: the call is host-synchronous)
  MG_CHECK_CUDA(cudaEventElapsedTime(&m->last_encode_ms, m->ev[0], m->ev[1]));
  MG_CHECK_CUDA(cudaEventElapsedTime(&m->last_decode_ms, m->ev[1], m->ev[2]));
  m->last_launches = m->launches - l0;
  m->last_ahead_ms = 0.f;
  if (took_slot >= 0) {  // the encoder ran ahead on the small partition: its own duration and launches
    MG_CHECK_CUDA(cudaEventElapsedTime(&m->last_ahead_ms, m->ahead[took_slot].begin, m->ahead[took_slot].done));
    m->last_launches += m->ahead[took_slot].n_launch;
  }
  MG_API_END
}

int mg_nccl_unique_id(void* out_128_bytes) {
  MG_API_BEGIN
  MG_REQUIRE(out_128_bytes, "null argument");
  ncclUniqueId id;
  MG_CHECK_NCCL(nccl_api().GetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  MG_API_END
}

int mg_comm_init(mg_model* m, int world, int rank, const void* id_128_bytes) {
  MG_API_BEGIN
  MG_REQUIRE(m && id_128_bytes && world >= 1 && rank >= 0 && rank < world, "bad arguments");
  MG_REQUIRE(m->comm == nullptr, "communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, id_128_bytes, sizeof(id));
  MG_CHECK_NCCL(nccl_api().CommInitRank(&m->comm, world, id, rank));
  m->world = world;
  m->rank = rank;
  // ---- peer-memory token exchange: every rank exports its exchange buffer (cudaIpc) and maps everybody else's
  const bool want_p2p = !(getenv("MG_DIST") && std::string(getenv("MG_DIST")) == "nccl") && world <= PX_MAXW;
  {
    const size_t n_int = peer_exchange_ints(world, m->px_bcap);
    m->px_local = m->own<int>((int64_t)n_int);
    MG_CHECK_CUDA(cudaMemset(m->px_local, 0, n_int * sizeof(int)));
    char* xb = m->own<char>((int64_t)(sizeof(cudaIpcMemHandle_t) + 8) * (world + 1));
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;  // handle + "ok so far" flag
    std::vector<char> mine(rec, 0), all(rec * world, 0);
    int ok = want_p2p ? 1 : 0;
    cudaIpcMemHandle_t h;
    if (ok && cudaIpcGetMemHandle(&h, m->px_local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ok) memcpy(mine.data(), &h, sizeof(h));
    memcpy(mine.data() + sizeof(h), &ok, sizeof(int));
    MG_CHECK_CUDA(cudaMemcpy(xb, mine.data(), rec, cudaMemcpyHostToDevice));
    MG_CHECK_NCCL(nccl_api().AllGather(xb, xb + rec, rec, ncclChar, m->comm, nullptr));
    MG_CHECK_CUDA(cudaMemcpy(all.data(), xb + rec, rec * world, cudaMemcpyDeviceToHost));
    std::vector<int*> bases(world, nullptr);
    for (int r = 0; r < world; ++r) {
      int okr;
      memcpy(&okr, all.data() + r * rec + sizeof(h), sizeof(int));
      ok = ok && okr;
    }
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) { bases[r] = m->px_local; continue; }
      cudaIpcMemHandle_t hr;
      memcpy(&hr, all.data() + r * rec, sizeof(hr));
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      m->px_opened.push_back(q);
      bases[r] = static_cast<int*>(q);
    }
    // second round: the mapping must have worked on EVERY rank, or all of them stay with the NCCL all-gather
    MG_CHECK_CUDA(cudaMemcpy(xb, &ok, sizeof(int), cudaMemcpyHostToDevice));
    MG_CHECK_NCCL(nccl_api().AllGather(xb, xb + rec, sizeof(int), ncclChar, m->comm, nullptr));
    std::vector<int> oks(world, 0);
    MG_CHECK_CUDA(cudaMemcpy(oks.data(), xb + rec, sizeof(int) * world, cudaMemcpyDeviceToHost));
    for (int r = 0; r < world; ++r) ok = ok && oks[r];
    if (ok) {
      m->px_peers_dev = reinterpret_cast<int**>(m->own<int*>(world));
      MG_CHECK_CUDA(cudaMemcpy(m->px_peers_dev, bases.data(), sizeof(int*) * world, cudaMemcpyHostToDevice));
    }
    m->px_ok = ok != 0;
  }
  MG_API_END
}

int mg_dist_mode(mg_model* m) {
  // 0 = single GPU, 1 = ncclAllGather of the token ids per decode step, 2 = NVLink peer stores fused into the selection kernel
  if (!m || !m->comm) return 0;
  return m->px_ok ? 2 : 1;
}

int mg_generate_dist(mg_model* m, void* stream, int B_local, int Lt, const int64_t* input_ids, const float* bbox,
                     const float* pixel_values, const int64_t* attn_mask, int num_beams, int max_length, int64_t* all_ids,
                     int32_t* steps_run) {
  MG_API_BEGIN
  MG_REQUIRE(m && m->comm, "mg_comm_init has not been called");
  MG_REQUIRE(all_ids, "null argument");
  MG_REQUIRE(num_beams >= 1 && num_beams <= 8, "1 <= num_beams <= 8");
  {  // every rank must bring the same shard size and length: unequal counts would hang or corrupt the all-gather
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (!m->dist_chk) m->dist_chk = m->own<int>(4 + 4 * 64);
    MG_REQUIRE(m->world <= 64, "at most 64 ranks");
    const int mine[4] = {B_local, max_length, num_beams, 0};
    MG_CHECK_CUDA(cudaMemcpyAsync(m->dist_chk, mine, sizeof(mine), cudaMemcpyHostToDevice, cs));
    MG_CHECK_NCCL(nccl_api().AllGather(m->dist_chk, m->dist_chk + 4, 4, ncclInt32, m->comm, cs));
    std::vector<int> all((size_t)4 * m->world);
    MG_CHECK_CUDA(cudaMemcpyAsync(all.data(), m->dist_chk + 4, sizeof(int) * all.size(), cudaMemcpyDeviceToHost, cs));
    MG_CHECK_CUDA(cudaStreamSynchronize(cs));
    for (int r = 0; r < m->world; ++r)
      MG_REQUIRE(all[4 * r] == B_local && all[4 * r + 1] == max_length && all[4 * r + 2] == num_beams,
                 "mg_generate_dist: rank " + std::to_string(r) + " brought B_local=" + std::to_string(all[4 * r]) +
                     " max_length=" + std::to_string(all[4 * r + 1]) + " num_beams=" + std::to_string(all[4 * r + 2]) +
                     ", this rank B_local=" + std::to_string(B_local) + " max_length=" + std::to_string(max_length) +
                     " num_beams=" + std::to_string(num_beams) + " (pad the shards to one size)");
  }
  int64_t* local = m->persist_ids((int64_t)B_local * max_length);
  m->dist_all_ids = all_ids;
  int rc = mg_generate(m, stream, B_local, Lt, input_ids, bbox, pixel_values, attn_mask, num_beams, max_length, local,
                       nullptr, nullptr, steps_run);
  m->dist_all_ids = nullptr;
  if (rc != 0) return rc;
  MG_API_END
}

int mg_forward_logits(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                      const float* pixel_values, const int64_t* attn_mask, const int64_t* decoder_input_ids, int T,
                      float* logits) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values && decoder_input_ids && logits, "null argument");
  MG_REQUIRE(T >= 1, "empty decoder input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (st == nullptr || st == cudaStreamLegacy) {
    MG_CHECK_CUDA(cudaDeviceSynchronize());
    if (!m->own_stream) MG_CHECK_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    st = m->own_stream;
  }
  m->encode(st, B, Lt, input_ids, bbox, pixel_values, attn_mask);
  // teacher forcing through the cached decode path: step t consumes decoder_input_ids[:, t] and its logits are
  // position t of the (B, T, V) output; T steps == max_length T+1
  int64_t* ids_scratch = m->persist.get<int64_t>((int64_t)B * (T + 1));
  MG_CHECK_CUDA(cudaEventRecord(m->ev[3], st));
  MG_CHECK_CUDA(cudaEventRecord(m->ev[5], st));
  m->generate(st, B, T + 1, ids_scratch, nullptr, logits, nullptr, decoder_input_ids, T);
  MG_API_END
}

int mg_generate_host(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                     const float* pixel_values, const int64_t* attn_mask, int num_beams, int max_length,
                     int64_t* out_ids, int32_t* out_len, int32_t* steps_run) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values && out_ids, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (st == nullptr || st == cudaStreamLegacy) {
    MG_CHECK_CUDA(cudaDeviceSynchronize());
    if (!m->own_stream) MG_CHECK_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    st = m->own_stream;
  }
  const mg_config& c = m->cfg;
  // a slot already filled by mg_prefetch_host from these very host buffers is used as is; otherwise stage now
  int slot = -1;
  for (int i = 0; i < 2; ++i) {
    const mg_model::HostStage& s = m->stage[i];
    if (s.valid && s.k_ids == input_ids && s.k_box == bbox && s.k_px == pixel_values && s.k_mask == attn_mask &&
        s.B == B && s.Lt == Lt && s.cap >= (size_t)((char*)s.d_out - (char*)s.buf) + (size_t)B * max_length * 8)
      slot = i;
  }
  if (slot >= 0) {
    MG_CHECK_CUDA(cudaStreamWaitEvent(st, m->stage[slot].ready, 0));
  } else {
    slot = m->pick_stage_slot();
    m->stage_inputs(m->stage[slot], st, B, Lt, input_ids, bbox, pixel_values, attn_mask, max_length);
  }
  mg_model::HostStage& S = m->stage[slot];
  S.valid = false;  // consumed: a later call with the same host pointers copies again (the caller may have refilled them)
  int64_t* d_out = S.d_out;
  int rc = mg_generate(m, st, B, Lt, S.d_ids, S.d_box, S.d_px, attn_mask ? S.d_mask : nullptr, num_beams, max_length, d_out,
                       nullptr, nullptr, steps_run);
  if (rc != 0) return rc;
  MG_CHECK_CUDA(cudaMemcpyAsync(out_ids, d_out, (size_t)B * max_length * 8, cudaMemcpyDeviceToHost, st));
  MG_CHECK_CUDA(cudaStreamSynchronize(st));
  if (out_len) {
    for (int b = 0; b < B; ++b) {
      int len = max_length;
      for (int t = 1; t < max_length; ++t)
        if (out_ids[(size_t)b * max_length + t] == c.eos_token_id) {
          len = t + 1;
          break;
        }
      out_len[b] = len;
    }
  }
  MG_API_END
}

int mg_prefetch_host(mg_model* m, int B, int Lt, const int64_t* input_ids, const float* bbox, const float* pixel_values,
                     const int64_t* attn_mask, int max_length) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values, "null argument");
  MG_REQUIRE(m->finalized, "mg_finalize has not been called");
  MG_REQUIRE(B > 0 && Lt > 0 && max_length >= 2, "bad sizes");
  if (!m->copy_stream) MG_CHECK_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  const int slot = m->pick_stage_slot();  // never a slot whose batch is still waiting for its generate call
  m->stage_inputs(m->stage[slot], m->copy_stream, B, Lt, input_ids, bbox, pixel_values, attn_mask, max_length);
  MG_API_END
}

int mg_encode_ahead(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                    const float* pixel_values, const int64_t* attn_mask, int32_t* armed) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values, "null argument");
  MG_REQUIRE(m->finalized, "mg_finalize has not been called");
  MG_REQUIRE(B > 0 && Lt > 0, "empty batch");
  if (armed) *armed = 0;
  if (m->ensure_partitions()) {
    m->encode_ahead(static_cast<cudaStream_t>(stream), B, Lt, input_ids, bbox, pixel_values, attn_mask);
    if (armed) *armed = 1;
  }
  MG_API_END
}

int mg_encode_ahead_host(mg_model* m, int B, int Lt, const int64_t* input_ids, const float* bbox,
                         const float* pixel_values, const int64_t* attn_mask, int max_length, int32_t* armed) {
  MG_API_BEGIN
  MG_REQUIRE(m && input_ids && bbox && pixel_values, "null argument");
  MG_REQUIRE(m->finalized, "mg_finalize has not been called");
  MG_REQUIRE(B > 0 && Lt > 0 && max_length >= 2, "bad sizes");
  if (armed) *armed = 0;
  if (!m->copy_stream) MG_CHECK_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  const int slot = m->pick_stage_slot();
  mg_model::HostStage& S = m->stage[slot];
  m->stage_inputs(S, m->copy_stream, B, Lt, input_ids, bbox, pixel_values, attn_mask, max_length);
  if (m->ensure_partitions()) {
    m->encode_ahead(m->copy_stream, B, Lt, S.d_ids, S.d_box, S.d_px, attn_mask ? S.d_mask : nullptr);
    if (armed) *armed = 1;
  }
  MG_API_END
}

int mg_ahead_reset(mg_model* m) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (m->part.ok) MG_CHECK_CUDA(cudaStreamSynchronize(m->part.s_small));
  for (auto& a : m->ahead) a.valid = false;
  MG_API_END
}

int mg_last_ahead(mg_model* m, float* encoder_ms, int32_t* sms_encoder, int32_t* sms_decoder) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (encoder_ms) *encoder_ms = m->last_ahead_ms;
  if (sms_encoder) *sms_encoder = m->part.ok ? m->part.n_small : 0;
  if (sms_decoder) *sms_decoder = m->part.ok ? m->part.n_big : 0;
  MG_API_END
}

int mg_profile_cross_attn(mg_model* m, void* stream, int reps, float* ms_per_launch, int64_t* bytes_per_launch,
                          int32_t* n_launches) {
  MG_API_BEGIN
  MG_REQUIRE(m && m->prof_q && !m->prof_ckt.empty(), "mg_profile_cross_attn needs a preceding mg_generate call");
  MG_REQUIRE(reps > 0, "reps must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (st == nullptr || st == cudaStreamLegacy) {
    MG_CHECK_CUDA(cudaDeviceSynchronize());
    if (!m->own_stream) MG_CHECK_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    st = m->own_stream;
  }
  const mg_config& c = m->cfg;
  // the production launch shape: one micro-batch lane (prof_bn images) per launch
  const int B = m->prof_bn > 0 ? m->prof_bn : m->cur_B;
  const int d = c.d_model, H = c.num_heads, Mp = m->dMp, NL = (int)m->prof_ckt.size();
  auto pass = [&]() {
    for (int l = 0; l < NL; ++l) {
      if (m->prof_kv24)
        launch_cross_attn_stream24(st, m->prof_q, B, H, d, m->prof_ckv[l], Mp, m->dmask, m->prof_ctx);
      else
        launch_cross_attn_stream(st, m->prof_q, B, H, d, m->prof_ckt[l], m->prof_cv[l], Mp, m->dmask, m->prof_ctx);
    }
  };
  pass();  // warm-up
  MG_CHECK_CUDA(cudaEventRecord(m->ev[0], st));
  for (int r = 0; r < reps; ++r) pass();
  MG_CHECK_CUDA(cudaEventRecord(m->ev[1], st));
  MG_CHECK_CUDA(cudaEventSynchronize(m->ev[1]));
  float ms = 0.f;
  MG_CHECK_CUDA(cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]));
  if (ms_per_launch) *ms_per_launch = ms / (float)(reps * NL);
  // algorithmic bytes of one launch: K and V of the true memory length M (fp32), the mask, q in, ctx planes out
  if (bytes_per_launch)
    *bytes_per_launch = (int64_t)B * ((int64_t)2 * m->dMp * d * (m->prof_kv24 ? 3 : 4) + (int64_t)m->dMp * 4 +
                                      (int64_t)d * 4 + (int64_t)d * (m->split2 ? 4 : 2));
  if (n_launches) *n_launches = reps * NL;
  MG_API_END
}

int mg_last_decode_p50(mg_model* m, float* step_p50_ms) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (step_p50_ms) *step_p50_ms = m->last_step_p50_ms;
  MG_API_END
}

int mg_last_decode_latency(mg_model* m, float* step_p50_ms, float* step_p99_ms) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (step_p50_ms) *step_p50_ms = m->last_step_p50_ms;
  if (step_p99_ms) *step_p99_ms = m->last_step_p99_ms;
  MG_API_END
}

int mg_last_decode_loop(mg_model* m, float* loop_ms, int32_t* steps, int32_t* fused) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (loop_ms) *loop_ms = m->last_loop_ms;
  if (steps) *steps = m->last_loop_steps;
  if (fused) *fused = m->last_fused;
  MG_API_END
}

int mg_last_memory_len(mg_model* m, int32_t* M_encoder, int32_t* M_decoder) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (M_encoder) *M_encoder = m->cur_M;
  if (M_decoder) *M_decoder = m->dMp;
  MG_API_END
}

int mg_launch_count(mg_model* m, int64_t* kernels_launched) {
  MG_API_BEGIN
  MG_REQUIRE(m && kernels_launched, "null argument");
  *kernels_launched = m->launches;
  MG_API_END
}

int mg_last_stats(mg_model* m, float* encode_ms, float* decode_ms, int64_t* kernels_launched) {
  MG_API_BEGIN
  MG_REQUIRE(m, "null model");
  if (encode_ms) *encode_ms = m->last_encode_ms;
  if (decode_ms) *decode_ms = m->last_decode_ms;
  if (kernels_launched) *kernels_launched = m->last_launches;
  MG_API_END
}

}  // extern "C"
