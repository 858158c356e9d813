/*
 * pllb.h — C ABI of libpllb200.so, the B200-native MLM-PLL N-best scoring path.
 *
 * The reference (ishine/ASR-Rescoring) is pure Python and has no FFI for this
 * path; the boundary that exists is a set of Python functions.  Each entry
 * point below names the reference interface it replaces (file:line relative to
 * the reference tree).  INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns 0 on success or a
 *     non-zero pllb_status, and pllb_last_error() gives the message.
 *   - "_host" entry points take HOST buffers and do their own H2D/D2H copies
 *     (this is what the end-to-end benchmark times).  The others take DEVICE
 *     pointers owned by the caller (torch allocations) plus a cudaStream_t
 *     passed as void*; they never synchronise the device unless stated.
 *   - one handle per GPU, driven from one host thread.  There is no CPU
 *     fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef PLLB_H_
#define PLLB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLLB_ABI_VERSION 1

typedef enum pllb_status {
  PLLB_OK = 0,
  PLLB_ERR_INVALID = 1,     /* bad argument / unsupported model shape          */
  PLLB_ERR_CUDA = 2,        /* CUDA runtime / driver error                      */
  PLLB_ERR_NO_DEVICE = 3,   /* no sm_100 GPU visible (there is no CPU fallback) */
  PLLB_ERR_OOM = 4,         /* workspace does not fit                           */
  PLLB_ERR_TOO_LONG = 5     /* a hypothesis exceeds max_position_embeddings-2   */
} pllb_status;

typedef struct pllb_context* pllb_handle;

/* Model shape.  Mirrors transformers.BertConfig as used by
 * MLM_PLL/main.py:184 (BertForMaskedLM.from_pretrained(config.model.bert)). */
typedef struct pllb_model_desc {
  int32_t num_layers;     /* 12 (bert-base-chinese) / 24                     */
  int32_t hidden;         /* 768 / 1024; multiple of 256                     */
  int32_t num_heads;      /* hidden / 64 (head dim is fixed at 64)           */
  int32_t intermediate;   /* 3072 / 4096; multiple of 256                    */
  int32_t vocab;          /* 21128                                           */
  int32_t max_position;   /* 512                                             */
  float   ln_eps;         /* 1e-12                                           */
  int32_t cls_id, sep_id, mask_id; /* 101, 102, 103                          */
  int32_t operand_dtype;  /* GEMM / attention operand type, fp32 accumulate
                             in every mode:
                             0 = bf16 activations and weights;
                             1 = IEEE fp16 activations and weights (same
                                 tensor-core rate, 3 more mantissa bits: ~8x
                                 smaller PLL error; conversions saturate at
                                 +-65504 instead of overflowing);
                             2 = bf16 encoder, fp16 MLM head (transform +
                                 decoder operands; 1 % of the FLOPs, removes
                                 the largest single rounding site);
                             3 = bf16 for encoder layers < NL/2, fp16 for the
                                 layers >= NL/2 and the head;
                             16 + k = bf16 for layers < k, fp16 from layer k on
                                 and in the head (16 = mode 1, 16 + NL = mode 2) */
} pllb_model_desc;

/* One encoder layer; DEVICE pointers to fp32 tensors in nn.Linear layout
 * ([out_features, in_features] row-major), i.e. the tensors of the
 * BertForMaskedLM state_dict loaded at MLM_PLL/main.py:185-187. */
typedef struct pllb_layer_weights {
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b;   /* attention.self.{query,key,value} */
  const float *ao_w, *ao_b, *ao_ln_g, *ao_ln_b;     /* attention.output.{dense,LayerNorm} */
  const float *ff1_w, *ff1_b;                       /* intermediate.dense  */
  const float *ff2_w, *ff2_b, *out_ln_g, *out_ln_b; /* output.{dense,LayerNorm} */
} pllb_layer_weights;

typedef struct pllb_weights {
  const float *word_emb, *pos_emb, *type_emb, *emb_ln_g, *emb_ln_b; /* bert.embeddings.* */
  const pllb_layer_weights* layers;  /* HOST array of num_layers structs */
  const float *head_w, *head_b, *head_ln_g, *head_ln_b; /* cls.predictions.transform.* */
  const float *decoder_w;   /* cls.predictions.decoder.weight [vocab, hidden] (tied to word_emb) */
  const float *decoder_b;   /* cls.predictions.bias [vocab] */
} pllb_weights;

/* Counters filled by pllb_get_stats (since create or the last reset). */
typedef struct pllb_stats {
  int64_t kernel_launches;   /* kernels of this library launched            */
  int64_t hyps_scored, copies_scored, tokens_expanded;
  int64_t chunks;
  double  gemm_flops;        /* algorithmic FLOPs issued to the GEMM kernel */
  float   last_gemm_ms;      /* device time of the GEMM kernels of the last
                                pllb_score call when timing is enabled      */
  float   last_total_ms;     /* device time of the last pllb_score call     */
  int64_t last_gemm_launches;
} pllb_stats;

const char* pllb_last_error(void);
int pllb_abi_version(void);

/* Number of CUDA devices usable by this library (compute capability 10.x).
 * Returns 0 when none: callers must fail, there is no fallback. */
int pllb_device_count(void);

/* ---- PLL scoring: replaces MLM_PLL/main.py:73-114 (run_one_epoch, scoring
 * branch), :28-54 (collate) and MLM_PLL/preprocess.py:9-30 (masked-copy
 * expansion).  Weights are converted to bf16 GEMM operands once here.
 * max_chunk_tokens bounds the expanded tokens (sum over hyps of L*(L+2))
 * processed per internal chunk and therefore the workspace size; 0 = default. */
int pllb_create(pllb_handle* out, const pllb_model_desc* desc,
                const pllb_weights* weights, int64_t max_chunk_tokens, int device);
int pllb_destroy(pllb_handle h);

/* Bytes of device workspace owned by the handle. */
int64_t pllb_workspace_bytes(pllb_handle h);

/* hyp_tokens   DEVICE int32[hyp_offsets[n_hyp]]  wordpiece ids, no specials; hypothesis i
 *                                                is hyp_tokens[hyp_offsets[i] .. hyp_offsets[i+1])
 * hyp_offsets  HOST   int64[n_hyp+1]             (control metadata)
 * out_pll      DEVICE double[n_hyp]   sum over the L masked copies of
 *                                     log_softmax(logits[mask_pos])[token]
 *                                     (a hypothesis with L == 0 gets 0.0)
 * out_token_logp DEVICE float[hyp_offsets[n_hyp]] or NULL; the individual terms
 * Asynchronous on `stream`: hyp_offsets is consumed before the call returns, the
 * device buffers when the stream reaches the work.  A handle owns ONE workspace:
 * use it from one host thread and issue its calls on one stream (calls may be queued
 * back to back without synchronising; work on different streams must be ordered by
 * the caller). */
int pllb_score(pllb_handle h, const int32_t* hyp_tokens, const int64_t* hyp_offsets,
               int32_t n_hyp, double* out_pll, float* out_token_logp, void* stream);

/* Same with HOST buffers; copies in, scores, copies out, synchronises. */
int pllb_score_host(pllb_handle h, const int32_t* hyp_tokens, const int64_t* hyp_offsets,
                    int32_t n_hyp, double* out_pll, float* out_token_logp);

/* ---- Sequence-level scoring: replaces the forward of RescoreBert/model.py:13-21 inside
 * RescoreBert/main.py run_one_epoch (scoring, :232-285): every hypothesis goes through the
 * encoder ONCE as [CLS] t [SEP] and out[i] = dot(last_hidden_state[i, 0, :], linear_w) +
 * linear_b.  The handle may be created without MLM head weights (head_w == NULL ...).
 * linear_w DEVICE float[hidden] (HOST in the _host variant); out_scores float[n_hyp]. */
int pllb_score_cls(pllb_handle h, const int32_t* hyp_tokens, const int64_t* hyp_offsets,
                   int32_t n_hyp, const float* linear_w, float linear_b, float* out_scores,
                   void* stream);
int pllb_score_cls_host(pllb_handle h, const int32_t* hyp_tokens, const int64_t* hyp_offsets,
                        int32_t n_hyp, const float* linear_w, float linear_b, float* out_scores);

/* Stage-1 alone, for parity tests against MLM_PLL/preprocess.py:9-30:
 * expands the hypotheses into the packed masked copies.
 * out_ids      DEVICE int32[sum L*(L+2)]  input_ids of every copy, packed
 * out_mask_pos DEVICE int32[sum L]        mask_pos of every copy (1-based incl. [CLS])
 * out_labels   DEVICE int32[sum L]        labels[mask_pos] of every copy        */
int pllb_expand(pllb_handle h, const int32_t* hyp_tokens, const int64_t* hyp_offsets,
                int32_t n_hyp, int32_t* out_ids, int32_t* out_mask_pos, int32_t* out_labels,
                void* stream);

int pllb_get_stats(pllb_handle h, pllb_stats* out);
int pllb_reset_stats(pllb_handle h);
/* enable=1: bracket GEMM launches with CUDA events on the launch stream so
 * last_gemm_ms is filled (adds event-record overhead only). */
int pllb_set_timing(pllb_handle h, int enable);
/* Device time (ms) per GEMM kind of the last timed pllb_score call and FLOPs issued per
 * kind since the last reset.  Order: QKV, attention-output, FFN1, FFN2, head transform,
 * decoder(+logsumexp).  Either pointer may be NULL. */
int pllb_get_gemm_breakdown(pllb_handle h, float* ms6, double* flops6);

/* ---- Debug / parity hooks (encoder internals; used by tests only) ---------
 * C[M,N] = A[M,K] * W[N,K]^T + bias[N], through the tcgen05 GEMM kernel.
 * A, W: DEVICE bf16 (raw uint16) row-major; bias fp32; epilogue:
 *   0 = bias -> bf16 out, 1 = bias+GELU(erf) -> bf16 out,
 *   2 = bias -> fp32 out, 3 = bias+GELU(erf) -> fp32 out.
 * N % 256 == 0, K % 64 == 0. */
int pllb_debug_gemm(const uint16_t* A, const uint16_t* W, const float* bias, void* C,
                    int32_t M, int32_t N, int32_t K, int32_t epilogue, void* stream);
/* Same with an explicit operand type: 0 = A and W bf16, 1 = A and W IEEE fp16 (raw uint16);
 * 16-bit outputs are written in the operand type. */
int pllb_debug_gemm_dt(const uint16_t* A, const uint16_t* W, const float* bias, void* C,
                       int32_t M, int32_t N, int32_t K, int32_t epilogue, int32_t operand_dtype, void* stream);
/* Same arithmetic through the plain SIMT validation kernel (slow). */
int pllb_debug_gemm_simt(const uint16_t* A, const uint16_t* W, const float* bias, void* C,
                         int32_t M, int32_t N, int32_t K, int32_t epilogue, void* stream);
/* Final hidden states (fp32 [tokens, hidden]) of the packed masked copies of
 * the given hypotheses after `upto_layer` encoder layers (0 = embeddings). */
int pllb_debug_hidden(pllb_handle h, const int32_t* hyp_tokens, const int64_t* hyp_offsets,
                      int32_t n_hyp, int32_t upto_layer, float* out_hidden, void* stream);

/* ---- Text front end: replaces BertTokenizer.tokenize + convert_tokens_to_ids
 * (MLM_PLL/preprocess.py:10,16-27,34) for hypotheses made of CJK ideographs,
 * punctuation and whitespace, where BERT's BasicTokenizer makes every character
 * its own token.  Strings are packed Unicode code points (HOST arrays).
 * table int32[table_size], indexed by code point:
 *   >= 0              wordpiece id of the character as a token ([UNK] if not in the vocab)
 *   PLLB_TOK_SPACE    whitespace: separates, emits nothing
 *   PLLB_TOK_REMOVED  control character, U+0000, U+FFFD: removed
 *   PLLB_TOK_WORD     part of a word run (Latin, digits, kana, marks ...): the hypothesis
 *                     needs the host wordpiece tokenizer; needs_host[h] = 1 and it gets
 *                     0 tokens here.  Code points >= table_size count as PLLB_TOK_WORD.
 * out_ids: capacity cp_off[n_hyp]; out_off int64[n_hyp+1]; needs_host uint8[n_hyp]. */
#define PLLB_TOK_SPACE (-1)
#define PLLB_TOK_REMOVED (-2)
#define PLLB_TOK_WORD (-3)
int pllb_tokenize_host(const int32_t* table, int32_t table_size, const int32_t* cp,
                       const int64_t* cp_off, int32_t n_hyp, int32_t* out_ids,
                       int64_t* out_off, uint8_t* needs_host);

/* ---- Levenshtein: replaces jiwer.cer's per-pair edit distance at
 * rescore.py:40,118 (and espnet_data/preprocess/main.py:59-60).
 * Strings are packed arrays of Unicode code points.
 * pair i compares ref[pair_ref[i]] with hyp i.  out_dist[i] = edit distance. */
int pllb_levenshtein(const int32_t* ref_cp, const int64_t* ref_off,
                     const int32_t* hyp_cp, const int64_t* hyp_off,
                     const int32_t* pair_ref, int32_t n_pairs, int32_t max_len,
                     int32_t* out_dist, void* stream);
int pllb_levenshtein_host(const int32_t* ref_cp, const int64_t* ref_off, int32_t n_ref,
                          const int32_t* hyp_cp, const int64_t* hyp_off,
                          const int32_t* pair_ref, int32_t n_pairs, int32_t* out_dist);

/* ---- Combiner: replaces rescore.py:47-58 (rescore + get_highest_score_hyp)
 * and the per-weight CER numerator of rescore.py:37-43 (find_best_weight).
 * variant 0: (1-w)*am/len + w*lm/len   (rescore.py:51, current source)
 * variant 1: (1-w)*am     + w*lm       (rescore_result/MLM_PLL/rescore.log:28)
 * variant 2: (1-w)*am/len + w*lm       (rescore_result/RMBR/BertScore/rescore_mbr_normalize.log:29)
 * All arrays [N, n_best] row-major; fp64, no FMA contraction, numpy op order.
 * out_argmax int32[W, N]; out_edit_sum int64[W] = sum_u dist[u, argmax]. */
int pllb_rescore_sweep(const double* am, const double* lm, const int64_t* len,
                       const int32_t* dist, int32_t N, int32_t n_best,
                       const double* weights, int32_t W, int32_t variant,
                       int32_t* out_argmax, int64_t* out_edit_sum, void* stream);
int pllb_rescore_sweep_host(const double* am, const double* lm, const int64_t* len,
                            const int32_t* dist, int32_t N, int32_t n_best,
                            const double* weights, int32_t W, int32_t variant,
                            int32_t* out_argmax, int64_t* out_edit_sum);
/* The [N, n_best] score matrix for one weight (rescore.py:47-53), DEVICE fp64. */
int pllb_rescore_scores(const double* am, const double* lm, const int64_t* len,
                        int32_t N, int32_t n_best, double weight, int32_t variant,
                        double* out_scores, void* stream);

/* ---- MLM fine-tuning: replaces the train_mode=True branch of run_one_epoch
 * (MLM_PLL/main.py:73-99: BertForMaskedLM.forward with labels, loss.backward(),
 * torch.optim.AdamW.step(), zero_grad) and its loss-only twin (train_mode=False,
 * do_scoring=False — the dev pass of mlm_finetune_bert, MLM_PLL/main.py:146-153).
 * A batch is what collate (MLM_PLL/main.py:28-54) builds: B sequences zero-padded to T
 * positions; the loss is CrossEntropyLoss() over ALL B*T positions (labels are the whole
 * unmasked sequence, pad positions carry label 0 = [PAD]; nothing is ignored).
 * fp32 master weights / gradients / Adam moments; bf16 operands, fp32 accumulation in
 * every GEMM (forward, dgrad, wgrad).  The decoder weight is tied to the word embeddings
 * and the decoder bias is cls.predictions.bias. */
typedef struct pllb_train_desc {
  float lr;                 /* config.lr (MLM_PLL/config/train.yaml:5: 1e-5)          */
  float beta1, beta2;       /* torch.optim.AdamW defaults 0.9, 0.999                  */
  float adam_eps;           /* 1e-8                                                    */
  float weight_decay;       /* 0.01, applied to EVERY parameter (model.parameters())  */
  float hidden_dropout;     /* BertConfig.hidden_dropout_prob (0.1); 0 for parity runs */
  float attention_dropout;  /* BertConfig.attention_probs_dropout_prob (0.1)          */
  uint64_t seed;            /* dropout stream (the masks are a stateless hash)        */
  int32_t pad_id;           /* BertConfig.pad_token_id (0): the embedding lookup sends
                               no gradient to this row (nn.Embedding padding_idx)     */
  int32_t max_rows;         /* capacity: B * T of the largest batch                    */
  int32_t max_seq;          /* capacity: largest T (<= max_position, <= 512)           */
} pllb_train_desc;

typedef struct pllb_trainer_ctx* pllb_trainer;

/* weights: DEVICE fp32 state_dict tensors (copied; the MLM head is required). */
int pllb_train_create(pllb_trainer* out, const pllb_model_desc* desc, const pllb_weights* weights,
                      const pllb_train_desc* train, int device);
int pllb_train_destroy(pllb_trainer t);
int64_t pllb_train_workspace_bytes(pllb_trainer t);
int64_t pllb_train_kernel_launches(pllb_trainer t);
/* Steps that ran as a replay of a captured CUDA graph (a batch shape (B, T, mode) runs eagerly
 * once, is captured on its second occurrence and replayed afterwards; PLLB_TRAIN_GRAPH=0 disables). */
int64_t pllb_train_graph_replays(pllb_trainer t);
/* A fresh optimizer: zero moments, step count 0, learning rate lr — the reference
 * re-creates AdamW at the start of every epoch (MLM_PLL/main.py:76). */
int pllb_train_reset_optimizer(pllb_trainer t, float lr);
/* One batch, HOST buffers.  input_ids, labels: int32[B*T] row-major (zero-padded as collate
 * does); n_valid: int32[B], the number of leading positions with attention_mask 1.
 * mode 0: loss only (model.eval(): no dropout, no update);
 * mode 1: forward (dropout active) + backward + AdamW step;
 * mode 2: forward + backward, no update (gradients stay readable: parity tests).
 * out_loss: the batch loss (output.loss.item(), MLM_PLL/main.py:109).  Synchronises. */
int pllb_train_step_host(pllb_trainer t, const int32_t* input_ids, const int32_t* n_valid,
                         const int32_t* labels, int32_t B, int32_t T, int32_t mode, float* out_loss);
/* Per-position losses of the last step, HOST float[B*T]: loss[b*T + t] = -log_softmax(logits[b, t])[labels[b, t]].
 * With the for_scoring rows of preprocess.py (labels = the unmasked sequence) the entry at the row's mask_pos is
 * minus the PLL term the scoring branch adds (MLM_PLL/main.py:101-107) — this is what makes
 * run_one_epoch(train_mode=..., do_scoring=True) work on a trainer. */
int pllb_train_row_losses_host(pllb_trainer t, float* out, int32_t n);
/* Writes the fp32 master weights (model.state_dict(), MLM_PLL/main.py:155) / the gradients of
 * the last mode-1/2 step into the caller's DEVICE tensors; a NULL pointer skips that tensor;
 * q/k/v are un-stacked; decoder_w is written only if it is not the word_emb pointer. */
int pllb_train_export(pllb_trainer t, const pllb_weights* dst);
int pllb_train_export_grads(pllb_trainer t, const pllb_weights* dst);

#ifdef __cplusplus
}
#endif
#endif  /* PLLB_H_ */
