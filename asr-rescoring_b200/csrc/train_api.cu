// train_api.cu — C ABI of the MLM fine-tuning path (include/pllb.h, pllb_train_*): replaces the
// train_mode=True branch of run_one_epoch (MLM_PLL/main.py:73-99) — BertForMaskedLM.forward with
// labels, loss.backward(), torch.optim.AdamW.step() — and its loss-only twin (train_mode=False,
// do_scoring=False, the dev pass of mlm_finetune_bert, :146-153).
//
// Numerics: fp32 master parameters, gradients and Adam moments in four flat buffers with one
// layout; every matrix product (forward, dgrad, wgrad) on the tcgen05 GEMM with bf16 operands and
// fp32 accumulation; LayerNorm / softmax / cross entropy / GELU in fp32.  The GEMM takes two
// K-contiguous operands, so a linear layer Y = X W^T keeps W and W^T as bf16 copies (refreshed after
// every optimizer step) and the backward pass transposes dY and X on the fly:
//     dX [R, K] = dY [R, N] . (W^T [K, N])^T          dW [N, K] = dY^T [N, R] . (X^T [K, R])^T
// with R padded to a multiple of 64 by zero columns.  Weight gradients are written by the GEMM
// epilogue straight into the flat gradient buffer.
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "train.h"

using namespace pllb;

namespace {

struct Span { int64_t off = 0, n = 0; };

struct TLayer {
  Span qkv_w, qkv_b, ao_w, ao_b, ao_g, ao_be, ff1_w, ff1_b, ff2_w, ff2_b, out_g, out_be;
  __nv_bfloat16 *qkv16, *qkvT16, *ao16, *aoT16, *ff116, *ff1T16, *ff216, *ff2T16;     // operand copies
  // activations kept for the backward pass
  __nv_bfloat16 *x16, *qkv, *ctx, *h1_16, *g16;
  float *lse, *xhat1, *rstd1, *f, *xhat2, *rstd2;
};

enum Site : uint32_t { SITE_EMB = 1, SITE_ATT = 2, SITE_AO = 3, SITE_FF2 = 4 };   // + 8 * layer

}  // namespace

struct pllb_trainer_ctx {
  pllb_model_desc d{};
  pllb_train_desc t{};
  int device = 0;
  int Vp = 0;
  int64_t cap_rows = 0, cap_rows_pad = 0;
  std::vector<void*> owned;
  int64_t owned_bytes = 0;
  // flat fp32 buffers, one layout: parameters, gradients, Adam first / second moments
  float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr;
  int64_t n_flat = 0;
  Span word, pos, type, emb_g, emb_b, head_w, head_b, head_g, head_be, dec_b;
  std::vector<TLayer> L;
  __nv_bfloat16 *head16 = nullptr, *headT16 = nullptr, *E16 = nullptr, *ET16 = nullptr;
  // batch inputs
  int32_t *ids = nullptr, *labels = nullptr, *n_valid = nullptr;
  int32_t* host_stage = nullptr;     // pinned
  // stash outside the layers
  float *xhat_e = nullptr, *rstd_e = nullptr, *t_f32 = nullptr, *xhat_h = nullptr, *rstd_h = nullptr;
  __nv_bfloat16 *xlast16 = nullptr, *tn16 = nullptr, *dlogits16 = nullptr;
  float *logits = nullptr, *loss_rows = nullptr, *loss_dev = nullptr;
  // scratch
  float *h32 = nullptr, *y32 = nullptr, *dh32 = nullptr, *dz32 = nullptr, *dzd32 = nullptr, *Dq = nullptr, *zeros = nullptr,
        *colsum_scratch = nullptr;
  __nv_bfloat16 *d16 = nullptr, *dT16 = nullptr, *xT16 = nullptr;
  int64_t step = 0;        // optimizer steps since the last reset (Adam bias correction)
  int64_t calls = 0;       // forward passes since create (dropout stream)
  int64_t launches = 0;
  int last_rows = 0;       // B * T of the last step (pllb_train_row_losses_host)
  // per-step scalars (device copy read by the kernels; pinned host copy filled before every step)
  TrainStepParams* params_dev = nullptr;
  TrainStepParams* params_host = nullptr;
  // CUDA graphs of the step, one per (B, T, mode): the step is ~575 small launches (launch-bound at the
  // reference's batch size), all with shape-static arguments.  A shape runs eagerly the first time
  // (shared-memory opt-ins, tensor-map cache), is captured the second time and replayed from then on.
  cudaStream_t stream = nullptr;     // every step runs here (stream capture is not possible on the legacy stream)
  bool use_graph = true;   // PLLB_TRAIN_GRAPH=0 disables
  struct GraphEntry { int seen = 0; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
  std::unordered_map<uint64_t, GraphEntry> graphs;
  int64_t graph_replays = 0;
};

namespace {

#define RC(expr)            \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

template <typename T>
int talloc(pllb_trainer_ctx* c, T** out, int64_t count, bool zero = false) {
  void* p = nullptr;
  const int64_t bytes = std::max<int64_t>(count, 1) * (int64_t)sizeof(T);
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(PLLB_ERR_OOM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
  }
  if (zero) cudaMemset(p, 0, (size_t)bytes);
  c->owned.push_back(p);
  c->owned_bytes += bytes;
  *out = reinterpret_cast<T*>(p);
  return PLLB_OK;
}

Span take(int64_t& cursor, int64_t n) {
  Span s{cursor, n};
  cursor += round_up(n, 64);
  return s;
}

int gemm(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K, int epi, cudaStream_t s) {
  return launch_gemm_tcgen05(A, W, bias, C, M, N, K, epi, nullptr, DT_BF16, s);
}

// bf16 copies (row-major and transposed) of one GEMM weight [N, K] of the flat parameter buffer
int refresh_weight(pllb_trainer_ctx* c, Span w, int N, int K, __nv_bfloat16* w16, __nv_bfloat16* wT16, cudaStream_t s) {
  return launch_train_cast_transpose(c->P + w.off, false, N, K, N, w16, wT16, s);
}

int refresh_all(pllb_trainer_ctx* c, cudaStream_t s) {
  const int H = c->d.hidden, I = c->d.intermediate;
  for (auto& l : c->L) {
    RC(refresh_weight(c, l.qkv_w, 3 * H, H, l.qkv16, l.qkvT16, s));
    RC(refresh_weight(c, l.ao_w, H, H, l.ao16, l.aoT16, s));
    RC(refresh_weight(c, l.ff1_w, I, H, l.ff116, l.ff1T16, s));
    RC(refresh_weight(c, l.ff2_w, H, I, l.ff216, l.ff2T16, s));
  }
  RC(refresh_weight(c, c->head_w, H, H, c->head16, c->headT16, s));
  RC(refresh_weight(c, c->word, c->Vp, H, c->E16, c->ET16, s));
  return PLLB_OK;
}

// dX = dY . W  and  dW = dY^T . X  of one linear layer, plus the bias gradient.
//   dy32 [R, N] fp32 (gradient of the layer's output); x16 [R, K] the layer's bf16 input
//   dx32 [R, K] fp32 out (may be null: the input needs no gradient)
int linear_bwd(pllb_trainer_ctx* c, const float* dy32, const __nv_bfloat16* x16, const __nv_bfloat16* wT16, int R, int N, int K,
               Span w, Span b, float* dx32, cudaStream_t s) {
  const int Rp = (int)round_up(R, 64);
  RC(launch_train_cast_transpose(dy32, false, R, N, Rp, c->d16, c->dT16, s));
  RC(launch_train_colsum(dy32, false, nullptr, R, N, c->G + b.off, nullptr, c->colsum_scratch, s));
  RC(launch_train_cast_transpose(x16, true, R, K, Rp, nullptr, c->xT16, s));
  RC(gemm(c->dT16, c->xT16, c->zeros, c->G + w.off, N, K, Rp, EPI_BIAS_F32, s));            // dW [N, K]
  if (dx32) RC(gemm(c->d16, wT16, c->zeros, dx32, R, K, N, EPI_BIAS_F32, s));                // dX [R, K]
  return PLLB_OK;
}

int copy_in(pllb_trainer_ctx* c, Span dst, const float* src, int64_t n, int64_t dst_off = 0) {
  if (!src) return fail(PLLB_ERR_INVALID, "pllb_train_create: a weight pointer is null (the MLM head is required)");
  PLLB_CUDA(cudaMemcpyAsync(c->P + dst.off + dst_off, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, 0));
  return PLLB_OK;
}
int copy_out(const float* buf, Span src, const float* dst, int64_t n, int64_t src_off = 0) {
  if (!dst) return PLLB_OK;
  PLLB_CUDA(cudaMemcpyAsync(const_cast<float*>(dst), buf + src.off + src_off, sizeof(float) * n, cudaMemcpyDeviceToDevice, 0));
  return PLLB_OK;
}

// state_dict tensors <- flat buffer `buf` (parameters or gradients)
int export_flat(pllb_trainer_ctx* c, const float* buf, const pllb_weights* w) {
  const int H = c->d.hidden, I = c->d.intermediate, V = c->d.vocab;
  const int64_t HH = (int64_t)H * H;
  RC(copy_out(buf, c->word, w->word_emb, (int64_t)V * H));
  RC(copy_out(buf, c->pos, w->pos_emb, (int64_t)c->d.max_position * H));
  RC(copy_out(buf, c->type, w->type_emb, 2 * H));
  RC(copy_out(buf, c->emb_g, w->emb_ln_g, H));
  RC(copy_out(buf, c->emb_b, w->emb_ln_b, H));
  for (int l = 0; l < c->d.num_layers; ++l) {
    const pllb_layer_weights& lw = w->layers[l];
    const TLayer& t = c->L[l];
    RC(copy_out(buf, t.qkv_w, lw.q_w, HH, 0));
    RC(copy_out(buf, t.qkv_w, lw.k_w, HH, HH));
    RC(copy_out(buf, t.qkv_w, lw.v_w, HH, 2 * HH));
    RC(copy_out(buf, t.qkv_b, lw.q_b, H, 0));
    RC(copy_out(buf, t.qkv_b, lw.k_b, H, H));
    RC(copy_out(buf, t.qkv_b, lw.v_b, H, 2 * H));
    RC(copy_out(buf, t.ao_w, lw.ao_w, HH));
    RC(copy_out(buf, t.ao_b, lw.ao_b, H));
    RC(copy_out(buf, t.ao_g, lw.ao_ln_g, H));
    RC(copy_out(buf, t.ao_be, lw.ao_ln_b, H));
    RC(copy_out(buf, t.ff1_w, lw.ff1_w, (int64_t)I * H));
    RC(copy_out(buf, t.ff1_b, lw.ff1_b, I));
    RC(copy_out(buf, t.ff2_w, lw.ff2_w, (int64_t)H * I));
    RC(copy_out(buf, t.ff2_b, lw.ff2_b, H));
    RC(copy_out(buf, t.out_g, lw.out_ln_g, H));
    RC(copy_out(buf, t.out_be, lw.out_ln_b, H));
  }
  RC(copy_out(buf, c->head_w, w->head_w, HH));
  RC(copy_out(buf, c->head_b, w->head_b, H));
  RC(copy_out(buf, c->head_g, w->head_ln_g, H));
  RC(copy_out(buf, c->head_be, w->head_ln_b, H));
  // the decoder weight is the word-embedding matrix (tied): written only when the caller keeps a separate tensor
  if (w->decoder_w && w->decoder_w != w->word_emb) RC(copy_out(buf, c->word, w->decoder_w, (int64_t)V * H));
  RC(copy_out(buf, c->dec_b, w->decoder_b, V));
  PLLB_CUDA(cudaStreamSynchronize(0));
  return PLLB_OK;
}

TrainDrop make_drop(const pllb_trainer_ctx* c, float p, bool train) {
  TrainDrop d{};
  d.seed = &c->params_dev->seed;
  if (train && p > 0.f) {
    d.thresh = (uint32_t)std::min<double>(4294967295.0, (double)p * 4294967296.0);
    d.inv_keep = 1.f / (1.f - p);
  } else {
    d.thresh = 0;
    d.inv_keep = 1.f;
  }
  return d;
}

// Every launch of one step (shape-static arguments only; the per-step scalars are read from params_dev).
// mode 0: loss only (model.eval()); 1: forward + backward + AdamW step; 2: forward + backward, no update
int enqueue_step(pllb_trainer_ctx* c, int B, int T, int mode, cudaStream_t s) {
  const pllb_model_desc& d = c->d;
  const int H = d.hidden, I = d.intermediate, NH = d.num_heads, NL = d.num_layers, Vp = c->Vp, V = d.vocab;
  const int R = B * T, Rp = (int)round_up(R, 64);
  const bool train = mode != 0;
  const TrainDrop hd = make_drop(c, c->t.hidden_dropout, train), ad = make_drop(c, c->t.attention_dropout, train);
  float* P = c->P;
  // ---------------- forward
  RC(launch_train_embed(c->ids, T, P + c->word.off, P + c->pos.off, P + c->type.off, P + c->emb_g.off, P + c->emb_b.off,
                        d.ln_eps, R, H, hd, SITE_EMB, c->h32, NL > 0 ? (void*)c->L[0].x16 : (void*)c->xlast16, c->xhat_e,
                        c->rstd_e, s));
  for (int l = 0; l < NL; ++l) {
    TLayer& t = c->L[l];
    __nv_bfloat16* x_next = l + 1 < NL ? c->L[l + 1].x16 : c->xlast16;
    RC(gemm(t.x16, t.qkv16, P + t.qkv_b.off, t.qkv, R, 3 * H, H, EPI_BIAS_BF16, s));
    RC(launch_train_attn_fwd(t.qkv, c->n_valid, R, T, H, NH, ad, SITE_ATT + 8 * l, t.ctx, t.lse, s));
    RC(gemm(t.ctx, t.ao16, P + t.ao_b.off, c->y32, R, H, H, EPI_BIAS_F32, s));
    RC(launch_train_ln_fwd(c->y32, c->h32, P + t.ao_g.off, P + t.ao_be.off, d.ln_eps, R, H, hd, SITE_AO + 8 * l, c->h32,
                           t.h1_16, t.xhat1, t.rstd1, s));
    RC(gemm(t.h1_16, t.ff116, P + t.ff1_b.off, t.f, R, I, H, EPI_BIAS_F32, s));
    RC(launch_train_gelu_fwd(t.f, (int64_t)R * I, t.g16, nullptr, s));
    RC(gemm(t.g16, t.ff216, P + t.ff2_b.off, c->y32, R, H, I, EPI_BIAS_F32, s));
    RC(launch_train_ln_fwd(c->y32, c->h32, P + t.out_g.off, P + t.out_be.off, d.ln_eps, R, H, hd, SITE_FF2 + 8 * l, c->h32,
                           x_next, t.xhat2, t.rstd2, s));
  }
  // MLM head on EVERY row (labels are the whole sequence, MLM_PLL/preprocess.py:24-28)
  RC(gemm(c->xlast16, c->head16, P + c->head_b.off, c->t_f32, R, H, H, EPI_BIAS_F32, s));
  RC(launch_train_gelu_fwd(c->t_f32, (int64_t)R * H, nullptr, c->y32, s));
  TrainDrop none{&c->params_dev->seed, 0, 1.f};
  RC(launch_train_ln_fwd(c->y32, nullptr, P + c->head_g.off, P + c->head_be.off, d.ln_eps, R, H, none, 0, nullptr, c->tn16,
                         c->xhat_h, c->rstd_h, s));
  RC(gemm(c->tn16, c->E16, P + c->dec_b.off, c->logits, R, Vp, H, EPI_BIAS_F32, s));
  RC(launch_train_ce(c->logits, c->labels, R, V, Vp, c->loss_rows, train ? c->dlogits16 : nullptr, c->loss_dev, s));
  if (train) {
    float* G = c->G;
    // ---------------- backward: head
    RC(launch_train_cast_transpose(c->dlogits16, true, R, Vp, Rp, nullptr, c->dT16, s));
    const float inv_r = 1.f / (float)R;        // the mean over the B*T positions, applied in fp32 (ce_kernel)
    RC(launch_train_colsum(c->dlogits16, true, nullptr, R, Vp, G + c->dec_b.off, nullptr, c->colsum_scratch, s));
    RC(launch_train_scale(G + c->dec_b.off, Vp, inv_r, s));
    RC(launch_train_cast_transpose(c->tn16, true, R, H, Rp, nullptr, c->xT16, s));
    RC(gemm(c->dT16, c->xT16, c->zeros, G + c->word.off, Vp, H, Rp, EPI_BIAS_F32, s));       // decoder part of dE
    RC(launch_train_scale(G + c->word.off, (int64_t)Vp * H, inv_r, s));
    RC(gemm(c->dlogits16, c->ET16, c->zeros, c->dh32, R, H, Vp, EPI_BIAS_F32, s));           // d(transform LayerNorm output)
    RC(launch_train_scale(c->dh32, (int64_t)R * H, inv_r, s));
    RC(launch_train_ln_bwd(c->dh32, nullptr, P + c->head_g.off, c->xhat_h, c->rstd_h, R, H, none, -1, -1, c->dz32, nullptr, s));
    RC(launch_train_colsum(c->dh32, false, c->xhat_h, R, H, G + c->head_be.off, G + c->head_g.off, c->colsum_scratch, s));
    RC(launch_train_gelu_bwd(c->dz32, c->t_f32, (int64_t)R * H, s));
    RC(linear_bwd(c, c->dz32, c->xlast16, c->headT16, R, H, H, c->head_w, c->head_b, c->dh32, s));
    // ---------------- backward: encoder layers.  dh32 = gradient of the layer output through the
    // layer above; dz32 = the part that arrives over the residual connection (added inside ln_bwd)
    bool have_res = false;
    for (int l = NL - 1; l >= 0; --l) {
      TLayer& t = c->L[l];
      const bool dr = hd.thresh != 0;
      // BertOutput: LN(dropout(FFN2(g)) + h1)
      RC(launch_train_ln_bwd(c->dh32, have_res ? c->dz32 : nullptr, P + t.out_g.off, t.xhat2, t.rstd2, R, H, hd, -1,
                             dr ? (int)(SITE_FF2 + 8 * l) : -1, c->dz32, dr ? c->dzd32 : nullptr, s));
      RC(launch_train_colsum(c->dh32, false, t.xhat2, R, H, G + t.out_be.off, G + t.out_g.off, c->colsum_scratch, s));
      RC(linear_bwd(c, dr ? c->dzd32 : c->dz32, t.g16, t.ff2T16, R, H, I, t.ff2_w, t.ff2_b, c->y32, s));   // y32 = dg [R, I]
      RC(launch_train_gelu_bwd(c->y32, t.f, (int64_t)R * I, s));
      RC(linear_bwd(c, c->y32, t.h1_16, t.ff1T16, R, I, H, t.ff1_w, t.ff1_b, c->dh32, s));                 // dh32 = dh1 via the FFN
      // BertSelfOutput: LN(dropout(AO(ctx)) + x)
      RC(launch_train_ln_bwd(c->dh32, c->dz32, P + t.ao_g.off, t.xhat1, t.rstd1, R, H, hd, -1,
                             dr ? (int)(SITE_AO + 8 * l) : -1, c->dz32, dr ? c->dzd32 : nullptr, s));
      RC(launch_train_colsum(c->dh32, false, t.xhat1, R, H, G + t.ao_be.off, G + t.ao_g.off, c->colsum_scratch, s));
      RC(linear_bwd(c, dr ? c->dzd32 : c->dz32, t.ctx, t.aoT16, R, H, H, t.ao_w, t.ao_b, c->dh32, s));     // dh32 = dctx
      RC(launch_train_attn_bwd(t.qkv, c->dh32, t.lse, c->n_valid, R, T, H, NH, ad, SITE_ATT + 8 * l, c->y32, c->Dq, s));
      RC(linear_bwd(c, c->y32, t.x16, t.qkvT16, R, 3 * H, H, t.qkv_w, t.qkv_b, c->dh32, s));               // dh32 = dx via QKV
      have_res = true;
    }
    // ---------------- backward: embeddings (dropout sits AFTER the LayerNorm here)
    RC(launch_train_ln_bwd(c->dh32, have_res ? c->dz32 : nullptr, P + c->emb_g.off, c->xhat_e, c->rstd_e, R, H, hd,
                           hd.thresh != 0 ? (int)SITE_EMB : -1, -1, c->dz32, nullptr, s));
    RC(launch_train_colsum(c->dh32, false, c->xhat_e, R, H, G + c->emb_b.off, G + c->emb_g.off, c->colsum_scratch, s));
    RC(launch_train_embed_bwd(c->dz32, c->ids, B, T, H, d.max_position, c->t.pad_id, G + c->word.off, G + c->pos.off,
                              G + c->type.off, c->colsum_scratch, s));
    if (mode == 1) {
      RC(launch_train_adamw(c->P, c->G, c->M, c->V, c->n_flat, c->t.lr, c->t.beta1, c->t.beta2, c->t.adam_eps,
                            c->t.weight_decay, c->params_dev, s));
      RC(refresh_all(c, s));
    }
  }
  return PLLB_OK;
}

int step_impl(pllb_trainer_ctx* c, int B, int T, int mode, float* out_loss_host, cudaStream_t s) {
  // per-step scalars -> device (the previous step ended with a stream synchronise, so the pinned copy is free)
  c->calls += 1;
  if (mode == 1) c->step += 1;
  const int64_t t = std::max<int64_t>(c->step, 1);
  const double bc1 = 1.0 - std::pow((double)c->t.beta1, (double)t), bc2 = 1.0 - std::pow((double)c->t.beta2, (double)t);
  c->params_host->seed = c->t.seed * 0x9E3779B97F4A7C15ull + (uint64_t)c->calls * 0xC2B2AE3D27D4EB4Full;
  c->params_host->adam_step_size = (float)((double)c->t.lr / bc1);
  c->params_host->adam_bc2_sqrt = (float)std::sqrt(bc2);
  PLLB_CUDA(cudaMemcpyAsync(c->params_dev, c->params_host, sizeof(TrainStepParams), cudaMemcpyHostToDevice, s));
  const uint64_t key = ((uint64_t)(uint32_t)B << 34) | ((uint64_t)(uint32_t)T << 2) | (uint64_t)mode;
  pllb_trainer_ctx::GraphEntry* ge = c->use_graph ? &c->graphs[key] : nullptr;
  if (ge && ge->exec) {
    PLLB_CUDA(cudaGraphLaunch(ge->exec, s));
    c->launches += ge->launches;
    c->graph_replays += 1;
  } else if (ge && ge->seen >= 1) {
    if (c->graphs.size() > 256) {              // bound the cache (a caller that never repeats a shape): start over
      for (auto& kv : c->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      c->graphs.clear();
      ge = &c->graphs[key];
      ge->seen = 1;
    }
    const int64_t before = g_launch_counter;
    PLLB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_step(c, B, T, mode, s);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess || !g) return fail(PLLB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) return fail(PLLB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei));
    ge->exec = exec;
    ge->launches = g_launch_counter - before;
    PLLB_CUDA(cudaGraphLaunch(ge->exec, s));
    c->launches += ge->launches;
    c->graph_replays += 1;
  } else {
    const int64_t before = g_launch_counter;
    RC(enqueue_step(c, B, T, mode, s));
    c->launches += g_launch_counter - before;
    if (ge) ge->seen += 1;
  }
  if (out_loss_host) {
    PLLB_CUDA(cudaMemcpyAsync(out_loss_host, c->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, s));
    PLLB_CUDA(cudaStreamSynchronize(s));
  }
  return PLLB_OK;
}

}  // namespace

extern "C" {

int pllb_train_create(pllb_trainer* out, const pllb_model_desc* desc, const pllb_weights* w, const pllb_train_desc* td,
                      int device) {
  if (!out || !desc || !w || !w->layers || !td) return fail(PLLB_ERR_INVALID, "pllb_train_create: null argument");
  *out = nullptr;
  if (pllb_device_count() < 1) return fail(PLLB_ERR_NO_DEVICE, "no sm_100 (B200) device visible; libpllb200 has no CPU fallback");
  const pllb_model_desc& d = *desc;
  if (d.hidden % 256 != 0 || d.hidden < 256 || d.hidden > 1024 || d.num_heads * 64 != d.hidden || d.intermediate % 256 != 0 ||
      d.num_layers < 1 || d.vocab < 1 || d.max_position < 3)
    return fail(PLLB_ERR_INVALID, "unsupported model shape: need num_layers >= 1, hidden in {256,512,768,1024}, head dim 64, "
                                  "intermediate % 256 == 0");
  if (td->max_rows < 1 || td->max_seq < 1 || td->max_seq > d.max_position || td->max_seq > 512 ||
      td->hidden_dropout < 0.f || td->hidden_dropout >= 1.f || td->attention_dropout < 0.f || td->attention_dropout >= 1.f)
    return fail(PLLB_ERR_INVALID, "pllb_train_create: need max_rows >= 1, 1 <= max_seq <= min(max_position, 512), dropout in [0, 1)");
  PLLB_CUDA(cudaSetDevice(device));
  int major = 0;
  PLLB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(PLLB_ERR_NO_DEVICE, "device is not sm_100 (B200); libpllb200 has no other code path");
  pllb_trainer_ctx* c = new pllb_trainer_ctx();
  c->d = d;
  c->t = *td;
  c->device = device;
  c->Vp = (int)round_up(d.vocab, 256);
  const int H = d.hidden, I = d.intermediate, V = d.vocab, NL = d.num_layers, NH = d.num_heads, Vp = c->Vp;
  const int64_t HH = (int64_t)H * H;
  int rc = PLLB_OK;
#define TRY(expr)                        \
  do {                                   \
    rc = (expr);                         \
    if (rc) { pllb_train_destroy(c); return rc; } \
  } while (0)
  // ---- flat layout
  int64_t cur = 0;
  c->word = take(cur, (int64_t)Vp * H);       // rows >= vocab stay zero (zero gradient, zero moments)
  c->pos = take(cur, (int64_t)d.max_position * H);
  c->type = take(cur, 2 * H);
  c->emb_g = take(cur, H);
  c->emb_b = take(cur, H);
  c->L.resize(NL);
  for (auto& l : c->L) {
    l.qkv_w = take(cur, 3 * HH); l.qkv_b = take(cur, 3 * H);
    l.ao_w = take(cur, HH); l.ao_b = take(cur, H); l.ao_g = take(cur, H); l.ao_be = take(cur, H);
    l.ff1_w = take(cur, (int64_t)I * H); l.ff1_b = take(cur, I);
    l.ff2_w = take(cur, (int64_t)H * I); l.ff2_b = take(cur, H); l.out_g = take(cur, H); l.out_be = take(cur, H);
  }
  c->head_w = take(cur, HH); c->head_b = take(cur, H); c->head_g = take(cur, H); c->head_be = take(cur, H);
  c->dec_b = take(cur, Vp);
  c->n_flat = cur;
  TRY(talloc(c, &c->P, cur, true));
  TRY(talloc(c, &c->G, cur, true));
  TRY(talloc(c, &c->M, cur, true));
  TRY(talloc(c, &c->V, cur, true));
  // ---- parameters in
  TRY(copy_in(c, c->word, w->word_emb, (int64_t)V * H));
  TRY(copy_in(c, c->pos, w->pos_emb, (int64_t)d.max_position * H));
  TRY(copy_in(c, c->type, w->type_emb, 2 * H));
  TRY(copy_in(c, c->emb_g, w->emb_ln_g, H));
  TRY(copy_in(c, c->emb_b, w->emb_ln_b, H));
  for (int l = 0; l < NL; ++l) {
    const pllb_layer_weights& lw = w->layers[l];
    TLayer& t = c->L[l];
    TRY(copy_in(c, t.qkv_w, lw.q_w, HH, 0)); TRY(copy_in(c, t.qkv_w, lw.k_w, HH, HH)); TRY(copy_in(c, t.qkv_w, lw.v_w, HH, 2 * HH));
    TRY(copy_in(c, t.qkv_b, lw.q_b, H, 0)); TRY(copy_in(c, t.qkv_b, lw.k_b, H, H)); TRY(copy_in(c, t.qkv_b, lw.v_b, H, 2 * H));
    TRY(copy_in(c, t.ao_w, lw.ao_w, HH)); TRY(copy_in(c, t.ao_b, lw.ao_b, H));
    TRY(copy_in(c, t.ao_g, lw.ao_ln_g, H)); TRY(copy_in(c, t.ao_be, lw.ao_ln_b, H));
    TRY(copy_in(c, t.ff1_w, lw.ff1_w, (int64_t)I * H)); TRY(copy_in(c, t.ff1_b, lw.ff1_b, I));
    TRY(copy_in(c, t.ff2_w, lw.ff2_w, (int64_t)H * I)); TRY(copy_in(c, t.ff2_b, lw.ff2_b, H));
    TRY(copy_in(c, t.out_g, lw.out_ln_g, H)); TRY(copy_in(c, t.out_be, lw.out_ln_b, H));
  }
  TRY(copy_in(c, c->head_w, w->head_w, HH)); TRY(copy_in(c, c->head_b, w->head_b, H));
  TRY(copy_in(c, c->head_g, w->head_ln_g, H)); TRY(copy_in(c, c->head_be, w->head_ln_b, H));
  TRY(copy_in(c, c->dec_b, w->decoder_b, V));
  // ---- operand copies, stash, scratch
  c->cap_rows = td->max_rows;
  c->cap_rows_pad = round_up(c->cap_rows, 64);
  const int64_t R = c->cap_rows, Rp = c->cap_rows_pad;
  const int wide = std::max(3 * H, I);
  for (auto& l : c->L) {
    TRY(talloc(c, &l.qkv16, 3 * HH)); TRY(talloc(c, &l.qkvT16, 3 * HH));
    TRY(talloc(c, &l.ao16, HH)); TRY(talloc(c, &l.aoT16, HH));
    TRY(talloc(c, &l.ff116, (int64_t)I * H)); TRY(talloc(c, &l.ff1T16, (int64_t)I * H));
    TRY(talloc(c, &l.ff216, (int64_t)I * H)); TRY(talloc(c, &l.ff2T16, (int64_t)I * H));
    TRY(talloc(c, &l.x16, R * H)); TRY(talloc(c, &l.qkv, R * 3 * H)); TRY(talloc(c, &l.ctx, R * H));
    TRY(talloc(c, &l.h1_16, R * H)); TRY(talloc(c, &l.g16, R * I));
    TRY(talloc(c, &l.lse, R * NH)); TRY(talloc(c, &l.xhat1, R * H)); TRY(talloc(c, &l.rstd1, R));
    TRY(talloc(c, &l.f, R * I)); TRY(talloc(c, &l.xhat2, R * H)); TRY(talloc(c, &l.rstd2, R));
  }
  TRY(talloc(c, &c->head16, HH)); TRY(talloc(c, &c->headT16, HH));
  TRY(talloc(c, &c->E16, (int64_t)Vp * H)); TRY(talloc(c, &c->ET16, (int64_t)Vp * H));
  TRY(talloc(c, &c->ids, R)); TRY(talloc(c, &c->labels, R)); TRY(talloc(c, &c->n_valid, R));
  TRY(talloc(c, &c->xhat_e, R * H)); TRY(talloc(c, &c->rstd_e, R)); TRY(talloc(c, &c->t_f32, R * H));
  TRY(talloc(c, &c->xhat_h, R * H)); TRY(talloc(c, &c->rstd_h, R));
  TRY(talloc(c, &c->xlast16, R * H)); TRY(talloc(c, &c->tn16, R * H)); TRY(talloc(c, &c->dlogits16, R * Vp));
  TRY(talloc(c, &c->logits, R * Vp)); TRY(talloc(c, &c->loss_rows, R)); TRY(talloc(c, &c->loss_dev, 1));
  TRY(talloc(c, &c->h32, R * H)); TRY(talloc(c, &c->y32, R * wide)); TRY(talloc(c, &c->dh32, R * H));
  TRY(talloc(c, &c->dz32, R * H)); TRY(talloc(c, &c->dzd32, R * H)); TRY(talloc(c, &c->Dq, R * NH));
  TRY(talloc(c, &c->zeros, std::max(Vp, wide), true));
  TRY(talloc(c, &c->colsum_scratch, (int64_t)2 * TRAIN_COLSUM_SPLITS * std::max(Vp, wide)));
  TRY(talloc(c, &c->d16, R * std::max(wide, H))); TRY(talloc(c, &c->dT16, (int64_t)std::max(wide, Vp) * Rp));
  TRY(talloc(c, &c->xT16, (int64_t)std::max(I, H) * Rp));
  TRY(talloc(c, &c->params_dev, 1, true));
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    pllb_train_destroy(c);
    return fail(PLLB_ERR_CUDA, "pllb_train_create: cudaStreamCreate failed");
  }
  if (const char* e = getenv("PLLB_TRAIN_GRAPH")) c->use_graph = atoi(e) != 0;
  if (cudaHostAlloc(&c->params_host, sizeof(TrainStepParams), cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc(&c->host_stage, sizeof(int32_t) * (size_t)(3 * R), cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    pllb_train_destroy(c);
    return fail(PLLB_ERR_OOM, "pllb_train_create: pinned staging buffer");
  }
  TRY(refresh_all(c, 0));
#undef TRY
  cudaError_t e = cudaStreamSynchronize(0);
  if (e != cudaSuccess) {
    pllb_train_destroy(c);
    return fail(PLLB_ERR_CUDA, std::string("pllb_train_create: ") + cudaGetErrorString(e));
  }
  *out = c;
  return PLLB_OK;
}

int pllb_train_destroy(pllb_trainer c) {
  if (!c) return PLLB_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : c->owned) cudaFree(p);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->host_stage) cudaFreeHost(c->host_stage);
  if (c->params_host) cudaFreeHost(c->params_host);
  delete c;
  return PLLB_OK;
}

int64_t pllb_train_workspace_bytes(pllb_trainer c) { return c ? c->owned_bytes : 0; }
int64_t pllb_train_kernel_launches(pllb_trainer c) { return c ? c->launches : 0; }
int64_t pllb_train_graph_replays(pllb_trainer c) { return c ? c->graph_replays : 0; }

int pllb_train_reset_optimizer(pllb_trainer c, float lr) {
  if (!c) return fail(PLLB_ERR_INVALID, "null trainer");
  PLLB_CUDA(cudaSetDevice(c->device));
  c->t.lr = lr;
  c->step = 0;
  PLLB_CUDA(cudaMemsetAsync(c->M, 0, sizeof(float) * (size_t)c->n_flat, c->stream));
  PLLB_CUDA(cudaMemsetAsync(c->V, 0, sizeof(float) * (size_t)c->n_flat, c->stream));
  return PLLB_OK;
}

int pllb_train_step_host(pllb_trainer c, const int32_t* input_ids, const int32_t* n_valid, const int32_t* labels, int32_t B,
                         int32_t T, int32_t mode, float* out_loss) {
  if (!c) return fail(PLLB_ERR_INVALID, "null trainer");
  if (!input_ids || !n_valid || !labels || B < 1 || T < 1 || mode < 0 || mode > 2)
    return fail(PLLB_ERR_INVALID, "pllb_train_step_host: bad argument");
  if (T > c->t.max_seq)
    return fail(PLLB_ERR_TOO_LONG, "batch of " + std::to_string(T) + " positions; the trainer was created for " +
                                       std::to_string(c->t.max_seq));
  if ((int64_t)B * T > c->cap_rows)
    return fail(PLLB_ERR_OOM, "batch of " + std::to_string((int64_t)B * T) + " rows; the trainer was created for " +
                                  std::to_string(c->cap_rows));
  const int R = B * T;
  for (int i = 0; i < R; ++i) {
    // the reference's embedding lookup / CrossEntropyLoss raise on an index outside the vocabulary
    if (input_ids[i] < 0 || input_ids[i] >= c->d.vocab || labels[i] < 0 || labels[i] >= c->d.vocab)
      return fail(PLLB_ERR_INVALID, "token id or label outside the vocabulary at row " + std::to_string(i));
  }
  for (int b = 0; b < B; ++b)
    if (n_valid[b] < 1 || n_valid[b] > T) return fail(PLLB_ERR_INVALID, "n_valid must be in [1, T]");
  PLLB_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  PLLB_CUDA(cudaStreamSynchronize(s));       // the pinned staging buffer of the previous call is free
  std::memcpy(c->host_stage, input_ids, sizeof(int32_t) * R);
  std::memcpy(c->host_stage + R, labels, sizeof(int32_t) * R);
  std::memcpy(c->host_stage + 2 * R, n_valid, sizeof(int32_t) * B);
  PLLB_CUDA(cudaMemcpyAsync(c->ids, c->host_stage, sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
  PLLB_CUDA(cudaMemcpyAsync(c->labels, c->host_stage + R, sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
  PLLB_CUDA(cudaMemcpyAsync(c->n_valid, c->host_stage + 2 * R, sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
  float loss = 0.f;
  RC(step_impl(c, B, T, mode, &loss, s));
  c->last_rows = R;
  if (out_loss) *out_loss = loss;
  return PLLB_OK;
}

int pllb_train_row_losses_host(pllb_trainer c, float* out, int32_t n) {
  if (!c || !out || n < 0 || n > c->last_rows) return fail(PLLB_ERR_INVALID, "pllb_train_row_losses_host: bad argument (n exceeds the rows of the last step)");
  PLLB_CUDA(cudaSetDevice(c->device));
  PLLB_CUDA(cudaMemcpyAsync(out, c->loss_rows, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  PLLB_CUDA(cudaStreamSynchronize(c->stream));
  return PLLB_OK;
}

int pllb_train_export(pllb_trainer c, const pllb_weights* dst) {
  if (!c || !dst || !dst->layers) return fail(PLLB_ERR_INVALID, "pllb_train_export: null argument");
  PLLB_CUDA(cudaSetDevice(c->device));
  PLLB_CUDA(cudaStreamSynchronize(c->stream));
  return export_flat(c, c->P, dst);
}

int pllb_train_export_grads(pllb_trainer c, const pllb_weights* dst) {
  if (!c || !dst || !dst->layers) return fail(PLLB_ERR_INVALID, "pllb_train_export_grads: null argument");
  PLLB_CUDA(cudaSetDevice(c->device));
  PLLB_CUDA(cudaStreamSynchronize(c->stream));
  return export_flat(c, c->G, dst);
}

}  // extern "C"
