/* mg_b200.h — C ABI of libmg_b200.so: the B200-native forward+generate hot path of MarkushGrapher-2
 * (Swin-B OCSR encoder -> MLP projector -> UDOP vision-text-layout encoder -> T5 decoder, greedy decode).
 *
 * The reference has no FFI for this path: its interface is the Python object contract of the (un-vendored)
 * transformers-fork class MarkushgrapherForConditionalGeneration as exercised by markushgrapher.core
 * (SURVEY.md §8b). Each entry point below names the reference call it stands behind:
 *
 *   mg_create / mg_load_weight / mg_finalize  <- MarkushgrapherForConditionalGeneration(config) /
 *        .from_pretrained(path, config=config).to(device)   reference markushgrapher/core/common/begin.py:128-133
 *        model.init_molscribe_weights() / model.safe_load() reference begin.py:138,151,166
 *   mg_encode          <- model.encoder(...) inside generate (encode once)
 *                         transformers/generation/utils.py:765-803, models/udop/modeling_udop.py:1064-1256
 *   mg_generate(_host) <- model.generate(input_ids=, bbox=, pixel_values=, labels=, num_beams=, max_length=)
 *                         reference markushgrapher/utils/ocsr/utils_evaluation.py:269-285
 *   mg_forward_logits  <- model(**batch).logits   reference markushgrapher/core/trainers/curriculumTrainer.py:655
 *
 * Conventions: every function returns 0 on success or a negative error code; the message is available from
 * mg_last_error() (thread-local). Unless a parameter says "host", pointers are DEVICE pointers (e.g.
 * torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as void*. Calls are asynchronous on `stream`
 * except where stated. One mg_model per device; not thread-safe across concurrent calls on the same model.
 * The library never allocates caller-visible memory. No torch / C++ types cross this boundary.
 */
#ifndef MG_B200_H
#define MG_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MG_MAX_SWIN_STAGES 4

typedef struct mg_config {
  /* UDOP / T5 dims (transformers/models/udop/configuration_udop.py) */
  int32_t vocab_size, d_model, d_kv, d_ff, num_layers, num_decoder_layers, num_heads;
  int32_t rel_buckets, rel_max_distance, max_2d, image_size, patch_size;
  float ln_eps;
  /* OCSR Swin encoder (MolScribe swin_base_patch4_window12_384 geometry) */
  int32_t swin_image, swin_patch, swin_embed, swin_num_stages;
  int32_t swin_depths[MG_MAX_SWIN_STAGES], swin_heads[MG_MAX_SWIN_STAGES];
  int32_t swin_window;
  float swin_ln_eps;
  int32_t proj_hidden;
  /* numerics: 0 = fp32-parity (split-bf16 tensor-core GEMMs, ~2^-16), 1 = plain bf16 GEMM inputs */
  int32_t precision;
  float logit_scale; /* d_model^-0.5 for the tied-head UDOP convention, 1.0 otherwise */
  int32_t decoder_start_token_id, eos_token_id, pad_token_id;
  int32_t enc_chunk; /* images per encoder pass (bounds the S x S score scratch); 0 = default 64 */
} mg_config;

typedef struct mg_model mg_model;

const char* mg_last_error(void);
/* 1 if a CUDA device is visible to the library, else 0 (no compute is attempted) */
int mg_device_available(void);

int mg_create(const mg_config* cfg, mg_model** out);
void mg_destroy(mg_model* m);

/* Register one fp32 parameter by its state_dict name (HF naming of the stock UDOP/Swin modules, see
 * INTEGRATION.md). dtype: 0 = fp32 (only value accepted). The pointer is borrowed until mg_finalize returns.
 * Returns 1 (not an error) if the name is not used by the path. */
int mg_load_weight(mg_model* m, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int rank);
/* Repack all weights into the library's own layouts (split planes, fused QKV, ...). Synchronises `stream`.
 * After it returns the caller may free its tensors. */
int mg_finalize(mg_model* m, void* stream);

/* Encoder: input_ids (B,Lt) i64, bbox (B,Lt,4) f32 in [0,1], pixel_values (B,3,image,image) f32,
 * attn_mask (B,Lt) i64 or NULL. Writes enc_out (B, M, d_model) f32 and enc_mask (B, M) i32 if non-NULL, with
 * M = swin_tokens + Lt + (image/patch)^2 (reported through M_out, host pointer, may be NULL). */
int mg_encode(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
              const float* pixel_values, const int64_t* attn_mask, float* enc_out, int32_t* enc_mask, int32_t* M_out);

/* Generate: num_beams == 1 greedy (GenerationMixin._sample), 2..8 beam search (GenerationMixin._beam_search with
 * the defaults the reference leaves in place: length_penalty 1.0, early_stopping False, 1 returned sequence).
 * out_ids (B, max_length) i64: column 0 = decoder start id, finished rows padded with pad id.
 * out_len (B) i32 = tokens up to and including EOS (or the columns generated).
 * step_logits: NULL or (B, max_length-1, vocab) f32 receiving the logits of every step (debug / parity).
 * steps_run (host, may be NULL): number of decode steps executed. Synchronises `stream` before returning. */
int mg_generate(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                const float* pixel_values, const int64_t* attn_mask, int num_beams, int max_length, int64_t* out_ids,
                int32_t* out_len, float* step_logits, int32_t* steps_run);

/* Same, but every buffer is a HOST pointer (pinned memory recommended): inputs are copied host->device and the
 * ids device->host inside the call. This is the end-to-end entry the reference's generate() call maps to. */
int mg_generate_host(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                     const float* pixel_values, const int64_t* attn_mask, int num_beams, int max_length,
                     int64_t* out_ids, int32_t* out_len, int32_t* steps_run);

/* Starts the host->device copy of the NEXT batch's inputs (same HOST pointers that a later mg_generate_host call will
 * pass, pinned memory; they must stay unchanged until that call) on the model's own copy stream and returns at once,
 * so the transfer overlaps the decode loop of the batch in flight.  mg_generate_host recognises the staged buffers by
 * pointer identity and skips its own copies.  Optional: without it mg_generate_host copies inside the call. */
int mg_prefetch_host(mg_model* m, int B, int Lt, const int64_t* input_ids, const float* bbox, const float* pixel_values,
                     const int64_t* attn_mask, int max_length);

/* Encoder run-ahead (throughput mode for a stream of batches): queues the ENCODER of the NEXT batch (Swin, projector,
 * VTL encoder -- the tensor-core-bound part) on a small SM partition of its own (CUDA green contexts, 16 SMs by
 * default: MG_AHEAD_SMS; MG_AHEAD=0 disables) and returns at once.  The mg_generate / mg_generate_dist call that later
 * passes the very same device pointers (contents unchanged in between) takes the finished encoder memory instead of
 * encoding; a generate call made while another batch is being encoded ahead runs its decode loop -- HBM-bound, the
 * tensor cores idle -- on the remaining SMs, so the two overlap.  `stream`: the work is ordered after what is queued
 * there (the inputs).  *armed (may be NULL) = 1 if the encoder was queued, 0 if SM partitioning is unavailable (then
 * nothing happens and the later generate call encodes as usual).  Results are identical either way.
 * The reference has no counterpart: its evaluation loop encodes and decodes one batch at a time
 * (markushgrapher/utils/ocsr/utils_evaluation.py:262-285). */
int mg_encode_ahead(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                    const float* pixel_values, const int64_t* attn_mask, int32_t* armed);
/* Same for HOST inputs: mg_prefetch_host + run-ahead encoder on the staged copy; consumed by the mg_generate_host call
 * that passes the same host pointers. */
int mg_encode_ahead_host(mg_model* m, int B, int Lt, const int64_t* input_ids, const float* bbox,
                         const float* pixel_values, const int64_t* attn_mask, int max_length, int32_t* armed);
/* After a generate call that took a run-ahead batch: the encoder's duration on its partition (ms, 0 if the call encoded
 * itself) and the SM counts of the two partitions (0 if unavailable). */
int mg_last_ahead(mg_model* m, float* encoder_ms, int32_t* sms_encoder, int32_t* sms_decoder);
/* Waits for a run-ahead encoder still in flight and drops every batch that no generate call has taken (end of a stream
 * of batches: later generate calls use all SMs again). */
int mg_ahead_reset(mg_model* m);

/* ---- multi-GPU: image-batch sharding, one process per GPU, one NCCL all-gather of token ids per decode step ----
 * mg_nccl_unique_id: rank 0 obtains a 128-byte NCCL id (host buffer) and ships it to the other ranks by any means
 * (bench.py uses torch.distributed). mg_comm_init: every rank joins. mg_generate_dist: like mg_generate on this rank's
 * B_local images; after every step the new token ids of ALL ranks are all-gathered (ncclAllGather of B_local int32 per
 * rank) so every rank fills all_ids (world*B_local, max_length) i64 in global image order and all ranks stop on the
 * same step.  num_beams > 1 (the reference's predict.yaml default decodes with beams): the beams of an image stay on
 * its GPU; the per-step all-gather carries the token of every image's best running beam plus the rank's "done" flag
 * (B_local + 1 int32), ranks whose search is over stay frozen until all are, and one final all-gather replaces the
 * provisional columns of all_ids with the finished sequences. B_local and max_length must be equal on all ranks: the call starts with a handshake
 * (one small all-gather) and fails on every rank with a message naming the offending rank otherwise.
 *
 * Greedy decoding moves the tokens WITHOUT a collective kernel when every rank can map every peer's memory (cudaIpc,
 * NVLink / NVSwitch): the kernel that ends a decode step stores each image's token straight into an exchange buffer on
 * every rank and raises a flag; an extra CTA of the next step's launch scatters the previous step's world x B_local
 * tokens into all_ids and maintains the global stop counter, so ranks run at most two steps apart and stop together.
 * MG_DIST=nccl (environment, read at mg_comm_init) or any rank failing to map a peer keeps the per-step ncclAllGather.
 * mg_dist_mode: 0 single GPU, 1 NCCL all-gather per step, 2 peer stores. */
int mg_nccl_unique_id(void* out_128_bytes);
int mg_comm_init(mg_model* m, int world, int rank, const void* id_128_bytes);
int mg_dist_mode(mg_model* m);
int mg_generate_dist(mg_model* m, void* stream, int B_local, int Lt, const int64_t* input_ids, const float* bbox,
                     const float* pixel_values, const int64_t* attn_mask, int num_beams, int max_length, int64_t* all_ids,
                     int32_t* steps_run);

/* Teacher-forced logits, the reference's `model(**batch).logits` with decoder_input_ids = shift_right(labels):
 * decoder_input_ids (B, T) i64 -> logits (B, T, vocab) f32. Runs T cached decode steps with the next input forced
 * (same kernels as mg_generate). Synchronises `stream` before returning. */
int mg_forward_logits(mg_model* m, void* stream, int B, int Lt, const int64_t* input_ids, const float* bbox,
                      const float* pixel_values, const int64_t* attn_mask, const int64_t* decoder_input_ids, int T,
                      float* logits);

/* Decode-path switches (environment, read when the model is finalised / at every generate call):
 *   MG_DECODE=chain  use the per-operation kernel chain (CUDA-graph replayed) instead of the fused persistent
 *                    decode-step kernel, which is the default whenever B <= 32 and max_length <= 512;
 *   MG_KV24=0        keep the cross K/V in fp32 (kernel chain only);
 *   MG_MEGA_L2PF=<KB> bytes of the coming cross-attention phase each CTA of the fused kernel prefetches into L2
 *                    (default 384, 0 = off; timing only, results identical). */

/* Memory lengths of the last generate call: M_encoder = swin tokens + text + patches (what mg_encode returns);
 * M_decoder = positions the decoder actually holds K/V for.  Masked positions (text padding, the padded tail behind the
 * surviving patches) contribute exp(finfo.min - max) = 0 to every cross-attention, so they are dropped before the cross
 * K/V projection: valid rows in order, every image padded with masked rows to the batch maximum (rounded up to 8).
 * Results are unchanged; the bytes streamed per generated token shrink accordingly.  MG_COMPACT=0 keeps all positions. */
int mg_last_memory_len(mg_model* m, int32_t* M_encoder, int32_t* M_decoder);
/* kernels this model has launched since mg_finalize (every entry point counts its own launches) */
int mg_launch_count(mg_model* m, int64_t* kernels_launched);
/* statistics of the last mg_generate call (host pointers, any may be NULL) */
int mg_last_stats(mg_model* m, float* encode_ms, float* decode_ms, int64_t* kernels_launched);

/* The decode-step loop of the last greedy mg_generate call alone (encoder and cross-K/V projection excluded), timed
 * with CUDA events on the launch stream: total milliseconds, steps executed, and whether the fused persistent
 * decode-step kernel ran (1) or the per-operation kernel chain (0). Host pointers, any may be NULL. */
int mg_last_decode_loop(mg_model* m, float* loop_ms, int32_t* steps, int32_t* fused);
/* Per-step decode latency of the last greedy generate (BASELINE.json "p50 decode lat"): the kernel that ends a decode
 * step (greedy_select_kernel) stamps %globaltimer on the device; the figures are the median / 99th percentile of the
 * differences between consecutive steps' stamps, i.e. true per-step latencies, both decode paths. */
int mg_last_decode_p50(mg_model* m, float* step_p50_ms);
int mg_last_decode_latency(mg_model* m, float* step_p50_ms, float* step_p99_ms);

/* Measurement hook for bench.py: re-launches the decode cross-attention kernel (the dominant, HBM-bound kernel of
 * the path) over the cross-KV buffers of the preceding mg_generate call, every decoder layer in turn (each
 * layer's K/V is larger than L2, so every launch is cache-cold), `reps` times between CUDA events on `stream`.
 * Outputs are HOST pointers: average ms per launch, algorithmic bytes per launch, launches timed. */
int mg_profile_cross_attn(mg_model* m, void* stream, int reps, float* ms_per_launch, int64_t* bytes_per_launch,
                          int32_t* n_launches);

/* ---- unit-level entry points used by the parity tests (device pointers, fp32) ---- */

/* c[M,N] = act(a[M,K] * b[N,K]^T + bias[N]) + residual[M,N]; row-major.
 * planes: 2 = split-bf16 (near-fp32), 1 = plain bf16. swap_out: write c transposed ([N,M] row-major). */
int mg_op_gemm(void* stream, int M, int N, int K, const float* a, const float* b, float* c, const float* bias,
               const float* residual, int act, int planes, int block_n, int ksplit, int swap_out);

/* kv24, the storage format of the decoder's cross-attention K/V (the tensors the decode loop re-reads once per
 * generated token; reference: the fp32 cached key/value states of UdopAttention, modeling_udop.py:569-583): fp32
 * rounded to nearest-even at 24 bits (sign, 8 exponent, 15 mantissa bits) and stored as a 16-bit plane plus an 8-bit
 * plane. kt, v: (B, H, 64*Mp) fp32 each; out: (2, B, H, 64*Mp) fp32 = decode(encode(kt)), decode(encode(v)). */
int mg_op_kv24_roundtrip(void* stream, int B, int H, int Mp, const float* kt, const float* v, float* out);

/* ---- GPU input packing: the step in front of the path --------------------------------------
 * src: B page images, uint8 RGB, (B, Hin, Win, 3) on the device.  out: pixel_values (B, 3, Hout, Wout) fp32 =
 *   page_image.resize((Wout, Hout), resample=Image.LANCZOS | BILINEAR)      reference core/datasets/mdu_dataset.py:118
 *   -> image processor: x * 1/255, (x - mean) / std                           reference utils/common.py:34-42
 * The resize restates Pillow's two-pass fixed-point resampler exactly (bit-identical to PIL, tests/test_pack_*.py);
 * filter: 0 = BILINEAR, 1 = LANCZOS.  mean / std: 3 host floats each.  Asynchronous on `stream`. */
int mg_pack_pixels(void* stream, int B, int Hin, int Win, const uint8_t* src, int Hout, int Wout, int filter,
                   const float* mean3_host, const float* std3_host, float* out);
/* HOST-only: Pillow's resampling coefficient table for one axis (precompute_coeffs + normalize_coeffs_8bpc):
 * bounds_host (2*out_size: first source index, count), kk_host (out_size * ksize 22-bit fixed-point weights). */
int mg_resample_coeffs(int in_size, int out_size, int filter, int32_t* ksize_out, int32_t* bounds_host, int32_t* kk_host,
                       int kk_capacity);

/* ---- batched detokeniser: the step behind the path --------------------------------------
 * ids (B, T) i64 on the device -> the strings of MarkushTokenizer.decode_plus_decode_other_tokens, reference
 * markushgrapher/core/common/markush_tokenizer.py:615-670.  The table is built on the host from the tokenizer's
 * id -> piece list and the reference's <other_N> vocabulary (markushgrapher_b200/detok.py evaluates every string
 * predicate of the reference once per vocabulary entry): text = UTF-8 bytes emitted for an id (space marker
 * stripped; "<mapped> " for a known <other_N>), text_off (vocab+1) their offsets, flags (vocab): 1 contains <i>,
 * 2 equals </i>, 4 contains </i>, 8 location token, 16 <other_*> token, 32 "the previous token gets a space".
 * mg_detok_measure: row lengths (lens (B) i32 or NULL = T tokens per row) -> row_off_host (B+1) byte offsets
 * (synchronises); mg_detok_write: the bytes, rows back to back, into out (row_off_host[B] bytes, device). */
typedef struct mg_detok mg_detok;
int mg_detok_create(int vocab, const uint8_t* text_host, const int32_t* text_off_host, const uint8_t* flags_host,
                    mg_detok** out);
void mg_detok_destroy(mg_detok* d);
int mg_detok_measure(mg_detok* d, void* stream, int B, int T, const int64_t* ids, const int32_t* lens,
                     int64_t* row_off_host);
int mg_detok_write(mg_detok* d, void* stream, int B, int T, const int64_t* ids, uint8_t* out);

/* HOST-only: T5/UDOP relative-position bucket LUT, lut[n] = bucket of |relative_position| = n without the
 * bidirectional sign offset (transformers/models/udop/modeling_udop.py:466-512). Needs no GPU. */
int mg_rel_bucket_lut(int bidirectional, int num_buckets, int max_distance, int n_entries, int32_t* lut_host);

#ifdef __cplusplus
}
#endif
#endif
