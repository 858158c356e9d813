// common.h — host-side helpers shared by the translation units of libpllb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pllb.h"

namespace pllb {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern thread_local int64_t g_launch_counter;   // kernels launched by this library on this thread

#define PLLB_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::pllb::fail(PLLB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
  } while (0)

#define PLLB_LAUNCH_CHECK(name)                                                                \
  do {                                                                                         \
    ++::pllb::g_launch_counter;                                                                \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return ::pllb::fail(PLLB_ERR_CUDA, std::string("launch ") + name + ": " + cudaGetErrorString(_e)); \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count();   // SMs of the current device (cached)

// 2-D row-major tensor [rows, cols] of `elt_bytes`-byte elements, box = [box_rows, box_cols] with
// 128-byte swizzle (box_cols * elt_bytes must be 128).  Encoded once per distinct
// (base, shape, box) and cached per host thread (api.cu).
int get_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elt_bytes, uint64_t rows, uint64_t cols,
                uint32_t box_rows, uint32_t box_cols);

// K-blocked 16-bit activation [rows, cols] (cols % 64 == 0, buffer padded to a multiple of 32 rows): tiles of
// 32 rows x 64 columns (4 KB) stored contiguously, tile (rb, cb) at ((rb * cols/64 + cb) * 4096) bytes, so
// that every TMA box of a GEMM operand / epilogue is ONE contiguous burst instead of 32 pieces of 128 bytes at
// the row pitch (tools/pitch_probe.cu: 6.2-6.6 against 5.2-5.3 TB/s when streaming).  EXPERIMENT, off by
// default: inside the FFN2 mainloop the row-major form is faster.  4-D tensor map, coordinates
// {column in tile, row in tile, column tile, row tile}, box {64, 32, 1, 1}, 128-byte swizzle: the
// shared-memory image of a box is the one of the 2-D row-major map.
int get_tmap_blocked(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel (keyed by its address — every
// instantiation with the same signature has the same pointer TYPE) and device.
template <typename K>
inline cudaError_t opt_in_smem(K kern, int bytes) {
  static thread_local std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  const void* key = reinterpret_cast<const void*>(kern);
  for (const auto& d : done)
    if (d.first == key && d.second == dev) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.emplace_back(key, dev);
  return e;
}

// ---- GEMM (gemm_tcgen05.cu) -------------------------------------------------
enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // out 16-bit = acc + bias
  EPI_BIAS_GELU_BF16 = 1,  // out 16-bit = gelu_erf(acc + bias)
  EPI_BIAS_F32 = 2,        // out fp32 = acc + bias
  EPI_BIAS_GELU_F32 = 3,   // out fp32 = gelu_erf(acc + bias)
  EPI_LSE = 4              // vocab-tiled online logsumexp + label-column pick (no C output)
};

// Operand-type bits of a GEMM launch:
//   bit 0: A (activations) is IEEE fp16, else bf16      bit 1: W (weights) is fp16, else bf16
//   bit 2: the 16-bit output is written as fp16, else bf16
// tcgen05 kind::f16 encodes the A and B formats in separate descriptor fields, but a launch
// with A = bf16, B = fp16 dies with "illegal instruction" on B200 (measured, round 2): the two
// operands must have the same type, so only the combinations below are instantiated.
enum GemmDtype : int {
  DT_BF16 = 0,          // bf16 x bf16 -> bf16 out
  DT_BF16_OUT16 = 4,    // bf16 x bf16 -> fp16 out (gemm_ln only: feeds the fp16 MLM head of operand mode 2)
  DT_FP16 = 7           // fp16 x fp16 -> fp16 out
};

struct LseArgs {
  const int32_t* labels;   // [M] label column of each row
  float2* partials;        // [M, 2 * n_tiles_n] (running max, sum of exp) per 128-column half tile
  float* label_logit;      // [M] written by the tile that owns the label column
  int32_t vocab;           // real vocab size (columns >= vocab are masked)
};

// C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).  A, W bf16 row-major (K contiguous).
// N % 256 == 0 (pad W rows with zeros), K % 64 == 0.  ldc = N.
// dt: DT_BF16 or DT_FP16.
// c_blocked: the 16-bit output C is written in the K-blocked layout (get_tmap_blocked) instead of row-major.
int launch_gemm_tcgen05(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K,
                        int epilogue, const LseArgs* lse, int dt, cudaStream_t stream, bool c_blocked = false);
// hidden_f32 = LayerNorm(A * W^T + bias + hidden_f32) in place, hidden_16 = 16-bit copy
// (gemm_ln_tcgen05.cu: cluster of H/256 CTAs per 128-row block, row statistics through DSMEM).
// a_blocked: A is stored in the K-blocked layout.
int launch_gemm_ln(const void* A, const void* W, const float* bias, const float* gamma, const float* beta, float eps,
                   float* hidden_f32, void* hidden_16, int64_t M, int H, int K, int dt, cudaStream_t stream,
                   bool a_blocked = false);
int launch_gemm_simt(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K,
                     int epilogue, cudaStream_t stream);

// ---- fp32 residual stream layout ("T32") ----------------------------------------
// The fp32 hidden states are private to this library, so they are stored in the layout the
// fused GEMM+LayerNorm epilogue wants: rows in blocks of 32; inside a block the 16-byte
// groups of 4 columns of the 32 rows are contiguous:
//     float index of (row, col) = ((row/32 * H/4 + col/4) * 32 + row%32) * 4 + col%4
// A thread that owns one row (the TMEM register layout) then reads / writes consecutive
// 16-byte pieces together with its 31 warp neighbours: 512 contiguous bytes per access.
// Buffers are padded to a multiple of 32 rows.  The 16-bit operand copies stay row-major
// (they are TMA-loaded as GEMM A operands).
__host__ __device__ inline size_t t32_index(int64_t row, int col, int H) {
  return (((size_t)(row >> 5) * (size_t)(H >> 2) + (size_t)(col >> 2)) * 32 + (size_t)(row & 31)) * 4 + (size_t)(col & 3);
}

// ---- encoder pieces (encoder_kernels.cu) -------------------------------------
struct CopyPlan {          // device arrays, one entry per masked copy of the chunk
  int32_t* seq_start;      // first packed row of the copy
  int32_t* seq_len;        // T = L + 2
  int32_t* mask_row;       // packed row index of the [MASK] token
  int32_t* label;          // original token at the masked position
  int32_t* hyp;            // chunk-local hypothesis index
  int32_t* uniq_base;      // first unique row of the copy's hypothesis (layer-0 sharing)
  int32_t* row_src;        // [packed rows] unique row every packed row equals (layer-0 sharing)
};

int launch_expand_plan(const int32_t* tokens, const int32_t* hyp_tok_off, const int32_t* hyp_copy_base,
                       const int32_t* hyp_row_base, int32_t n_hyp, int32_t vocab, bool whole_sequence, CopyPlan plan,
                       cudaStream_t s);
int launch_expand_ids(const int32_t* tokens, const int32_t* hyp_tok_off, CopyPlan plan, int32_t n_copies,
                      int32_t cls_id, int32_t sep_id, int32_t mask_id, int32_t* out_ids, int32_t* out_mask_pos,
                      int32_t* out_labels, cudaStream_t s);
int launch_embed_ln(const int32_t* tokens, const int32_t* hyp_tok_off, CopyPlan plan, int32_t n_copies,
                    const float* word_emb, const float* pos_emb, const float* type_emb, const float* g,
                    const float* b, float eps, int H, int32_t cls_id, int32_t sep_id, int32_t mask_id, int32_t vocab,
                    float* hidden_f32_rowmajor, void* hidden_bf16, bool fp16, cudaStream_t s);
// hidden = LN(y + hidden) (in place), hidden_bf16 = bf16(hidden)
int launch_residual_ln(const float* y, float* hidden_f32, void* hidden_bf16, const float* g, const float* b,
                       float eps, int64_t rows, int H, bool fp16, cudaStream_t s);
// out_bf16 = bf16(LN(x)); no residual (MLM head transform)
int launch_plain_ln_bf16(const float* x, void* out_bf16, const float* g, const float* b, float eps, int64_t rows,
                         int H, bool fp16, cudaStream_t s);
// shared_rows: qkv holds the unique rows of every hypothesis (2L+2 per hypothesis) instead of one row per packed row
// qkv_rows: rows of the [rows, 3H] buffer behind qkv_bf16 (extent of the TMA tensor maps)
int launch_attention(const void* qkv_bf16, void* ctx_bf16, CopyPlan plan, int32_t n_copies, int H, int NH,
                     int max_T, bool fp16, bool shared_rows, int64_t qkv_rows, cudaStream_t s);
// Embeddings of the unique rows of every hypothesis (fp32 row-major + 16-bit), and the packed-row -> unique-row map.
int launch_embed_unique(const int32_t* tokens, const int32_t* hyp_tok_off, int32_t n_hyp, const float* word_emb,
                        const float* pos_emb, const float* type_emb, const float* g, const float* b, float eps, int H,
                        int32_t cls_id, int32_t sep_id, int32_t mask_id, int32_t vocab, float* u_f32, void* u_bf16,
                        bool fp16, cudaStream_t s);
int launch_row_src(CopyPlan plan, int32_t n_copies, cudaStream_t s);
// last layer: one query row per copy (q [copies,H]) against the K|V projection of the whole sequence (kv [rows,2H])
int launch_attention_row(const void* q_bf16, const void* kv_bf16, void* out_bf16, CopyPlan plan, int32_t n_copies, int H,
                         int NH, int max_T, bool fp16, cudaStream_t s);
int launch_gather_rows_bf16(const void* hidden_bf16, const int32_t* rows, int32_t n, int H, void* out,
                            cudaStream_t s);
int launch_gather_rows_f32(const float* src, const int32_t* rows, int32_t n, int H, float* out, cudaStream_t s);
int launch_cls_linear(const float* hid_t32, const float* w, float b, int32_t n, int H, float* out, cudaStream_t s);
int launch_lse_finish(const float2* partials, const float* label_logit, int32_t n_copies, int n_tiles,
                      float* tok_logp, cudaStream_t s);
int launch_hyp_sum(const float* tok_logp, const int32_t* hyp_copy_base, int32_t n_hyp, double* out_pll,
                   float* out_tok_logp, cudaStream_t s);
// T32 blocked fp32 [rows, H] -> row-major fp32 (debug / parity output)
// row_src (optional): packed row r is read from src row row_src[r]
int launch_rowmajor_to_t32(const float* src, const int32_t* row_src, float* dst, int64_t rows, int H, cudaStream_t s);
int launch_t32_to_rowmajor(const float* src, float* dst, int64_t rows, int H, cudaStream_t s);
int launch_f32_to_bf16(const float* src, void* dst, int64_t n, bool fp16, cudaStream_t s);

// ---- combiner (rescore_kernels.cu, compiled with -fmad=false) ------------------
int launch_rescore_sweep(const double* am, const double* lm, const int64_t* len, const int32_t* dist, int32_t N,
                         int32_t n_best, const double* weights, int32_t W, int32_t variant, int32_t* out_argmax,
                         int64_t* out_edit_sum, cudaStream_t s);
int launch_rescore_scores(const double* am, const double* lm, const int64_t* len, int32_t N, int32_t n_best,
                          double weight, int32_t variant, double* out, cudaStream_t s);
int launch_tokenize_count(const int32_t* table, int table_size, const int32_t* cp, const int64_t* cp_off, int32_t n_hyp,
                          int32_t* counts, uint8_t* needs_host, cudaStream_t s);
int launch_tokenize_write(const int32_t* table, int table_size, const int32_t* cp, const int64_t* cp_off, int32_t n_hyp,
                          const uint8_t* needs_host, const int64_t* out_off, int32_t* out_ids, cudaStream_t s);
int launch_levenshtein(const int32_t* ref_cp, const int64_t* ref_off, const int32_t* hyp_cp, const int64_t* hyp_off,
                       const int32_t* pair_ref, int32_t n_pairs, int32_t max_len, int32_t* out, cudaStream_t s);

}  // namespace pllb
