// Batched detokeniser on the device: generated token ids -> the CXSMILES / substituent-table
// string of reference markushgrapher/core/common/markush_tokenizer.py:615-670 (decode_plus_decode_other_tokens):
// a per-token table lookup plus a small state machine (skip <i>...</i> index spans, skip <loc_*>, map <other_N> to
// its vocabulary string + " ", strip the sentencepiece space marker and insert a space when the NEXT token starts a
// word or is an <other_*>).  Every string predicate of the reference is evaluated once per vocabulary entry on the
// host (markushgrapher_b200/detok.py) and shipped as flags; the kernels only follow the control flow.
#include <algorithm>

#include "kernels.h"

namespace mg {

__device__ __forceinline__ int dt_clamp_id(int64_t v, int vocab) { return v < 0 ? 0 : (v >= vocab ? vocab - 1 : (int)v); }

enum { DT_I_OPEN = 1, DT_I_CLOSE_EQ = 2, DT_I_CLOSE_IN = 4, DT_LOC = 8, DT_OTHER = 16, DT_NEXT_SPACE = 32 };

// one thread per row: token -> (byte offset inside the row, emitted length incl. the optional trailing space)
__global__ void detok_measure_kernel(const int64_t* __restrict__ ids, int B, int T, const int* __restrict__ lens,
                                     const int* __restrict__ text_off, const uint8_t* __restrict__ flags, int vocab,
                                     int* __restrict__ tok_off, int* __restrict__ tok_len, int64_t* __restrict__ row_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = lens ? min(max(lens[b], 0), T) : T;
  const int64_t* row = ids + (int64_t)b * T;
  int off = 0;
  bool skip = false;
  for (int t = 0; t < T; ++t) {
    int len = 0;
    if (t < n) {
      const int id = dt_clamp_id(row[t], vocab);
      const int f = flags[id];
      bool emit = true;
      if (skip && !(f & DT_I_CLOSE_EQ)) emit = false;          // inside an <i> ... </i> span
      if (emit) {
        skip = false;
        if (f & DT_I_OPEN) { skip = true; emit = false; }
        else if (f & (DT_I_CLOSE_IN | DT_LOC)) emit = false;
      }
      if (emit) {
        len = text_off[id + 1] - text_off[id];
        if (!(f & DT_OTHER) && t + 1 < n) {
          const int nid = dt_clamp_id(row[t + 1], vocab);
          if (flags[nid] & DT_NEXT_SPACE) len += 1;
        }
      }
    }
    tok_off[(int64_t)b * T + t] = off;
    tok_len[(int64_t)b * T + t] = len;
    off += len;
  }
  row_len[b] = off;
}

// exclusive scan of the row lengths (one block; B is a batch size)
__global__ void detok_scan_kernel(const int64_t* __restrict__ row_len, int B, int64_t* __restrict__ row_off) {
  if (threadIdx.x == 0) {
    int64_t s = 0;
    for (int b = 0; b < B; ++b) {
      row_off[b] = s;
      s += row_len[b];
    }
    row_off[B] = s;
  }
}

// one thread per (row, token): copy the token's bytes (+ the space) to its place
__global__ void detok_write_kernel(const int64_t* __restrict__ ids, int B, int T, const int* __restrict__ text_off,
                                   const uint8_t* __restrict__ text, int vocab, const int* __restrict__ tok_off,
                                   const int* __restrict__ tok_len, const int64_t* __restrict__ row_off,
                                   uint8_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T) return;
  const int len = tok_len[i];
  if (len == 0) return;
  const int b = (int)(i / T);
  const int id = dt_clamp_id(ids[i], vocab);
  const int tl = text_off[id + 1] - text_off[id];
  uint8_t* o = out + row_off[b] + tok_off[i];
  const uint8_t* s = text + text_off[id];
  for (int k = 0; k < tl; ++k) o[k] = s[k];
  if (len > tl) o[tl] = ' ';
}

void launch_detok_measure(cudaStream_t st, const int64_t* ids, int B, int T, const int* lens, const int* text_off,
                          const uint8_t* flags, int vocab, int* tok_off, int* tok_len, int64_t* row_len, int64_t* row_off) {
  detok_measure_kernel<<<(B + 127) / 128, 128, 0, st>>>(ids, B, T, lens, text_off, flags, vocab, tok_off, tok_len, row_len);
  MG_CHECK_CUDA(cudaGetLastError());
  detok_scan_kernel<<<1, 32, 0, st>>>(row_len, B, row_off);
  MG_CHECK_CUDA(cudaGetLastError());
}
void launch_detok_write(cudaStream_t st, const int64_t* ids, int B, int T, const int* text_off, const uint8_t* text,
                        int vocab, const int* tok_off, const int* tok_len, const int64_t* row_off, uint8_t* out) {
  const int64_t n = (int64_t)B * T;
  detok_write_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ids, B, T, text_off, text, vocab, tok_off, tok_len, row_off, out);
  MG_CHECK_CUDA(cudaGetLastError());
}

}  // namespace mg
