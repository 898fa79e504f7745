// Launch wrappers of all non-GEMM kernels (definitions in ops.cu / swin.cu / decode.cu).
#pragma once
#include <vector>

#include "mg_internal.h"

namespace mg {

// ---- ops.cu
void rel_bucket_lut(int bidirectional, int num_buckets, int max_distance, int n_entries, int32_t* lut);
void launch_rmsnorm(cudaStream_t st, const float* x, const float* w, int64_t rows, int D, float eps, float scale,
                    Planes out, float* out_f32, int64_t rows_per_b, int64_t out_bs, int64_t out_off);
void launch_im2col(cudaStream_t st, const float* px, int B, int H, int W, int p, int ldk, Planes out);
void launch_combine(cudaStream_t st, const int64_t* ids, const float* bbox, const int64_t* attn_mask,
                    const float* tok_emb, const float* patch_emb, const float* cell_x, const float* cell_y, int B,
                    int Lt, int np, int Sp, int D, int max_2d, int vocab, int* ocr_point, int* vis_src, int* n_vis,
                    float* x, double* bbox_ext, int* mask);
void launch_relbucket_hv(cudaStream_t st, const double* bbox_ext, int B, int Sp, const int* lut_hv, int lut_n,
                         int half_buckets, uchar2* hv);
void launch_enc_softmax(cudaStream_t st, const float* scores, const uchar2* hv, const int* mask, const float* tab1d,
                        const float* tabh, const float* tabv, const int* lut1d, int lut1d_n, int half_buckets,
                        int nbuckets, int B, int H, int Sp, Planes P);
void launch_fill_f32(cudaStream_t st, float* p, int64_t n, float v);
void launch_build_mem_mask(cudaStream_t st, const int* vtl_mask, int B, int Sp, int S, int n_sw, int Mp, int* out);
// decoder-side compaction of the encoder memory (masked positions dropped, valid rows in order, padded to Mc per image)
void launch_compact_plan(cudaStream_t st, const int* mask, int B, int Mp, int* src, int* n_valid);
void launch_compact_gather(cudaStream_t st, const float* mem, const int* src, const int* n_valid, int B, int Mp, int Mc,
                           int D, float* out, int* mask_out);

// ---- enc_flash.cu: fused encoder self-attention (scores, bucketed bias, online softmax, P.V in one tcgen05 kernel)
size_t enc_bias_code_bytes(int B, int Sp);
void launch_enc_bias_code(cudaStream_t st, const double* bbox_ext, const int* mask, int B, int Sp, const int* lut_hv,
                          int lut_n, int half_buckets, uint16_t* code);
// qk planes [B*Sp][2D] (q | k), vt planes [B][D][Sp], ctx planes [B*Sp][D]
void launch_enc_flash_attn(cudaStream_t st, Planes qk, Planes vt, const uint16_t* code, const float* tab1d,
                           const float* tabh, const float* tabv, const int* lut1d, int lut1d_n, int half_buckets,
                           int nbuckets, int B, int H, int D, int Sp, Planes ctx);

// ---- swin.cu
void launch_resize_bilinear(cudaStream_t st, const float* in, int B, int Hi, int Wi, int Ho, int Wo, float* out);
void launch_window_rowmap(cudaStream_t st, int B, int Hs, int Ws, int ws, int shift, int* map);
void launch_merge_rowmap(cudaStream_t st, int B, int Hs, int Ws, int* map);
void launch_layernorm_any(cudaStream_t st, const float* x, const int* src_rows, int G, int C, int64_t rows,
                          const float* w, const float* b, float eps, Planes out, float* out_f32);
void launch_window_attn(cudaStream_t st, const float* qkv, const float* table, int64_t n_windows, int C, int heads,
                        int ws, int Hs, int Ws, int shift, Planes out);

// ---- decode.cu
void launch_dec_self_attn(cudaStream_t st, const float* qkv, int B, int H, int D, float* kt, int64_t kt_ld,
                          int64_t kt_bs, float* v, int64_t v_ld, int64_t v_bs, const int* step_ptr, int max_keys,
                          const float* dec_bias, const int* lut, float* ctx);
void launch_dec_cross_attn(cudaStream_t st, const float* q, int B, int H, int D, const float* kt, int64_t kt_ld,
                           int64_t kt_bs, const float* v, int64_t v_ld, int64_t v_bs, int n_keys, const int* mask,
                           int mask_ld, float* ctx);
// streaming cross-attention: kt [B][H][64][Mp], v [B][H][Mp][64] (both contiguous per (b,h))
void launch_cross_attn_stream(cudaStream_t st, const float* q, int B, int H, int D, const float* kt, const float* v,
                              int Mp, const int* mask, float* ctx);
// kv24 cross K/V (decode.cu): fp32 rounded to 24 significant bits, stored as a 16-bit + an 8-bit plane
void launch_kv24_pack(cudaStream_t st, const float* kt, const float* v, int B, int H, int Mp, uint8_t* out);
void launch_kv24_roundtrip(cudaStream_t st, const float* kt, const float* v, int B, int H, int Mp, uint8_t* packed,
                           float* out);
void launch_cross_attn_stream24(cudaStream_t st, const float* q, int B, int H, int D, const uint8_t* kv, int Mp,
                                const int* mask, float* ctx);
void launch_relu_split(cudaStream_t st, const float* x, int64_t n, Planes out);
// Multi-GPU token exchange over NVLink peer memory, fused into the kernel that ends a decode step (decode.cu): every
// rank owns one exchange buffer  [flags: PX_RING x PX_MAXW][slots: PX_RING x world x bcap]  mapped into all peers
// (cudaIpc); greedy_select_kernel stores the step's token of each of its images into slot (step % PX_RING, rank) of
// EVERY rank's buffer and then raises that rank's flag to base + step + 1; an extra CTA of the next step's launch
// waits for all flags of the previous step and scatters the world x B tokens into the global id matrix (same stop
// rule on every rank).  No collective kernel runs inside the decode loop.
constexpr int PX_RING = 4;    // a rank can be at most two steps ahead of the slowest one (it waits for everyone's step s-1 in step s)
constexpr int PX_MAXW = 64;   // ranks
struct PeerExchange {
  int* const* peers = nullptr;  // device array [world] of exchange-buffer bases (this rank's own included); null = off
  int world = 1, rank = 0, bcap = 0, B = 0;
  unsigned base = 0;            // flag epoch of this generate call
  int64_t* all_ids = nullptr;   // (world * B, ld) global id matrix of this rank
  int ld = 0, eos = 1;
  int* gfinished = nullptr;     // [world * B]
  int* g_unfinished = nullptr;
  int* consumed = nullptr;      // launches of the consumer CTA so far in this generate call (device counter, starts at 0)
};
inline size_t peer_exchange_ints(int world, int bcap) { return (size_t)PX_RING * PX_MAXW + (size_t)PX_RING * world * bcap; }
// consumes the tokens of decode step `step` (after the loop: the last one)
void launch_peer_drain(cudaStream_t st, const PeerExchange& px, int step);
// part_val/part_idx: optional [B][n_part] partial maxima from the LM-head epilogue (then `logits` is only dumped)
void launch_greedy_select(cudaStream_t st, const float* part_val, const int* part_idx, int n_part, const float* logits, int B, int V, int64_t ld, const float* emb, int D,
                          int eos, int pad, int64_t* out_ids, int out_ld, int* finished, int* step_ptr,
                          int* n_unfinished, int* ticket, float* x_next, float* logits_dump, int64_t dump_bs,
                          int64_t dump_ss, const int64_t* forced, int forced_ld, int* step_tok,
                          unsigned long long* step_ts = nullptr, const PeerExchange* px = nullptr);
void launch_scatter_step(cudaStream_t st, const int* gathered, int n_rows, int col, int ld, int eos, int64_t* all_ids,
                         int* gfinished, int* g_unfinished);
void launch_decode_init(cudaStream_t st, const float* emb, int D, int start, int pad, int B, int64_t* out_ids,
                        int out_ld, int* finished, int* step_ptr, int* n_unfinished, int* ticket, float* x,
                        const int64_t* forced, int forced_ld);

void launch_out_len(cudaStream_t st, const int64_t* ids, int B, int ld, int ncols, int eos, int* len);

// ---- beam.cu
struct BeamState {
  int64_t* run_seq;
  int64_t* fin_seq;
  float* run_score;
  float* fin_score;
  int* fin_flag;
  int* fin_len;
  int* unsat;
  int* anc0;
  int* anc1;
  int* ctrl;
  int64_t* tmp_seq;  // [B][nb][L] scratch for the slot permutations
  int* step_tok = nullptr;  // multi-GPU: [B + 1] token of the best running beam per image (-1 once frozen) + this rank's done flag
  int B, nb, L, anc_ld;
};
void launch_beam_self_attn(cudaStream_t st, const float* qkv, int R, int H, int D, float* kt, int64_t kt_ld,
                           int64_t kt_bs, float* v, int64_t v_ld, int64_t v_bs, const int* step_ptr,
                           const int* anc_sel, const int* anc0, const int* anc1, int anc_ld, const float* dec_bias,
                           const int* lut, float* ctx);
void launch_beam_cross_attn(cudaStream_t st, const float* q, int B, int nq, int H, int D, const float* kt,
                            const float* v, int Mp, const int* mask, float* ctx);
void launch_beam_cross_attn24(cudaStream_t st, const float* q, int B, int nq, int H, int D, const uint8_t* kv, int Mp,
                              const int* mask, float* ctx);
// multi-GPU beam search: gathered [world][B + 1] (tokens + done flag per rank) -> provisional column `col` of all_ids,
// *n_not_done = ranks whose stop condition has not been reached
void launch_beam_scatter_step(cudaStream_t st, const int* gathered, int world, int B, int col, int ld, int64_t* all_ids,
                              int* n_not_done);
void launch_beam_init(cudaStream_t st, const BeamState& s, const float* emb, int D, int start, int pad, float* x);
void launch_beam_select(cudaStream_t st, const BeamState& s, const float* logits, int V, int64_t ld, const float* emb,
                        int D, int eos, int max_length, float* x_next);
void launch_beam_finalize(cudaStream_t st, const BeamState& s, int pad, int64_t* out_ids, int* out_len);

// gemm_tc.cu: tensor-core skinny linear with fused prologue (pro: 0 none, 1 RMSNorm, 2 ReLU), split-K atomics
void launch_skinny_tc(cudaStream_t st, int pro, const float* x, int ldx, Planes W, int64_t ldw, float* out, int ld_out,
                      int B, int N, int K, const float* lnw, float eps, float scale, float* zero_ptr, int64_t zero_n,
                      bool store, float* amax_val = nullptr, int* amax_idx = nullptr, float* rs_scratch = nullptr);
// (rs_scratch: >= B floats of device scratch; with more than 32 rows and a fused RMSNorm the row scales are computed once
//  by a small kernel in front of the linear instead of by every CTA of it)

// ---- pack.cu: GPU input packing (Pillow-exact resize + image-processor normalisation)
int resample_coeffs(int in_size, int out_size, int filter, std::vector<int>& bounds, std::vector<int>& kk);
void launch_pack_pixels(cudaStream_t st, int B, int Hin, int Win, const uint8_t* src, int Hout, int Wout, int filter,
                        const float* mean3, const float* std3, uint8_t* tmp, float* out);

// ---- detok.cu: batched detokeniser (ids -> bytes; flags / text table built on the host)
void launch_detok_measure(cudaStream_t st, const int64_t* ids, int B, int T, const int* lens, const int* text_off,
                          const uint8_t* flags, int vocab, int* tok_off, int* tok_len, int64_t* row_len, int64_t* row_off);
void launch_detok_write(cudaStream_t st, const int64_t* ids, int B, int T, const int* text_off, const uint8_t* text,
                        int vocab, const int* tok_off, const int* tok_len, const int64_t* row_off, uint8_t* out);

// ---- decode_mega.cu: the fused persistent decode step (one cooperative kernel per generated token)
struct MegaLin {           // a linear layer as a stream of pre-swizzled 32 KB (tile, k-block) weight tiles
  const uint8_t* w = nullptr;  // [tiles][num_kb][hi 16 KB | lo 16 KB]
  int N = 0, K = 0;
  int tiles = 0, num_kb = 0;
  int kb_per_item = 0, ksplit = 0;  // work items = tiles * ksplit, item -> (tile = it / ksplit, k-slice = it % ksplit)
  int n_split = 0;  // two-output linears: rows [0, n_split) go to the first output, [n_split, N) to the second (0 = one output)
  int n_split2 = 0; // three-segment linear (folded co|wi, -DMK_FOLD_FF): rows >= n_split2 read the SECOND activation source
};
struct MegaLayer {
  MegaLin lin[5];      // qkv|cq_x (ln1 / ln2 folded into the rows), o|cq_ctx, co, wi, wo -- decode_mega.cu phase table
  const float* ln[3];  // RMSNorm weights before qkv / cq / wi (the kernel reads only ln[2]; the others are folded)
  float* skb;        // self K cache [B][H][Tp/32][64][32]
  float* svb;        // self V cache [B][H][Tp][64]
  const uint8_t* ckv;  // cross K/V, kv24 blocks [B][H][384 * Mp] (decode.cu)
  const void* pad_;
};
struct MegaParams {
  const MegaLayer* layers;  // device array [NL]
  int NL;
  MegaLin lm_head;
  const float* final_ln;
  float logit_scale, eps;
  int B, H, D, DFF, Mp, Tp;
  float *x, *qkv, *q, *ctx, *hbuf, *logits;
  float* xalt = nullptr;  // -DMK_FOLD_FF: second residual-stream buffer (layer l reads X[l & 1], its wo phase writes the other)
  float* dx = nullptr;    // -DMK_FOLD_FF: [B][D] Wco . ctx of the current layer (contiguous behind qkv: zeroed together)
  int ld_logits;
  float* part_val;
  int* part_idx;
  const int* step_ptr;
  const int* mem_mask;
  const float* dec_bias;
  const int* lut;
  float* xp;          // [B*H][Mp + 4] softmax numerators (+ their sum) of cross-attention items split over two CTAs
  unsigned* xflag;    // [B*H] epoch flags of xp, zero before the first step
  float* rs;          // [3][32] RMSNorm row scales published by the qkv (slot 0) and wi (slot 2) phases for their consumers; slot 1 unused
  unsigned* bar_ctr;  // [2], zero before the first step
  int gate = 0;          // 1: attention K/V streams wait until the consumers enter their phase (measured: no gain)
  int* dbg_host = nullptr;  // host-mapped pinned words for the watchdog's diagnostics (may be null)
  int dbg = 0;           // -DMK_FINE builds: 1 = skip the reductions, 2 = skip the TMEM loads; any build: 4 = evict-normal stream, 8 / 16 = no cross K / V arithmetic (stream only)
  int max_inflight = 5;  // bulk loads one SM keeps in flight (<= ring stages)
  int cross_rk = 0, cross_vr = 0;  // experiments: d-rows per cross-K chunk / keys per cross-V chunk (0 = as many as fit a stage)
  // decode_mega.cu, producer: paced L2 prefetch of the first bytes of the coming cross-attention phase (0 = off)
  int l2pf = 384 * 1024;   // bytes per CTA and layer
  int l2pf_piece = 4096;   // bytes per prefetch instruction (multiple of 16)
  int l2pf_gap = 500;      // minimum cycles between two prefetch instructions of a CTA
  int l2pf_mask = 0x07;    // phases (bit = phase index) during whose loads the producer polls + prefetches; bit 8: poll in all
  unsigned long long* prof = nullptr;  // debug: [CTAs][256][2] globaltimer at (work done, barrier passed) per phase
};
size_t mega_lin_bytes(int N, int K);
// tiles the planes of one linear into dst (mega_lin_bytes(N, K) bytes) and fills the work split for n_ctas CTAs
MegaLin make_mega_lin(cudaStream_t st, Planes w, int N, int K, int64_t ldk, bool store, int n_ctas, uint8_t* dst, int n_split = 0,
                      int n_split2 = 0);
bool mega_fold_ff();  // true if decode_mega.cu was built with -DMK_FOLD_FF (co folded into wi: 6 phases per layer)
// finalize-time helpers of the folded cross query: dst[r][k] = src[r][k] * gain[k];  P[n][j] = sum_k A[n][k] * Bm[k][j] (fp64 accumulation)
void launch_scale_cols(cudaStream_t st, const float* src, const float* gain, int rows, int K, float* dst);
void launch_fold_product(cudaStream_t st, const float* A, const float* Bm, int N, int K, int J, float* P);
int mega_max_ctas();  // CTAs of the cooperative launch (= SM count), 0 if the device cannot run it
void launch_decode_step(cudaStream_t st, const MegaParams& p, int n_ctas);

}  // namespace mg
