/*
 * oracle.c — CPU restatement (plain C) of the integer / fp64 parts of the
 * MLM_PLL N-best scoring path of ishine/ASR-Rescoring.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may call this file:
 * it is the checker used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Pinning: the Levenshtein restatement is checked bit-exactly against the
 * 7 176 {ref,pred,cer} triples of the reference's Nbest_Align/cer.json, the
 * docstring examples of espnet_data/preprocess/align.py:13-18 and the integer
 * identities of the logged corpus CERs (tests/test_oracle.py, fixtures under
 * tests/golden/).  The combiner restatement is checked against the reference's
 * own rescore.py functions imported in the build container
 * (oracle/make_golden.py -> tests/golden/combiner_golden.npz).
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off, no fast-math)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------
 * Edit distance between two code-point strings, unit costs.
 * Follows: jiwer.cer as called at rescore.py:40 / rescore.py:118 and
 * espnet_data/preprocess/main.py:59-60 (third-party `jiwer`, unpinned by the
 * reference, absent from the tree: characters = Python code points after
 * strip(); S+D+I of the minimum-cost alignment), and the in-tree DP of
 * espnet_data/preprocess/align.py:27-50 (distance table without backtrace).
 * ------------------------------------------------------------------------- */
int32_t oracle_levenshtein(const int32_t* a, int32_t na, const int32_t* b, int32_t nb) {
  int32_t* row = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nb + 1));
  for (int32_t j = 0; j <= nb; ++j) row[j] = j;
  for (int32_t i = 1; i <= na; ++i) {
    int32_t diag = row[0];
    row[0] = i;
    for (int32_t j = 1; j <= nb; ++j) {
      int32_t up = row[j];
      int32_t sub = diag + (a[i - 1] != b[j - 1]);
      int32_t del = up + 1;
      int32_t ins = row[j - 1] + 1;
      int32_t v = sub < del ? sub : del;
      row[j] = v < ins ? v : ins;
      diag = up;
    }
  }
  int32_t d = row[nb];
  free(row);
  return d;
}

/* pair i: ref[pair_ref[i]] vs hyp i (same packing as pllb_levenshtein). */
void oracle_levenshtein_batch(const int32_t* ref_cp, const int64_t* ref_off,
                              const int32_t* hyp_cp, const int64_t* hyp_off,
                              const int32_t* pair_ref, int32_t n_pairs, int32_t* out_dist) {
  for (int32_t i = 0; i < n_pairs; ++i) {
    int32_t r = pair_ref[i];
    out_dist[i] = oracle_levenshtein(ref_cp + ref_off[r], (int32_t)(ref_off[r + 1] - ref_off[r]),
                                     hyp_cp + hyp_off[i], (int32_t)(hyp_off[i + 1] - hyp_off[i]));
  }
}

/* ---------------------------------------------------------------------------
 * rescore() — rescore.py:47-53.  numpy evaluates
 *     (1-weight)*(am)/hyps_len + weight*(lm)/hyps_len
 * left to right as  (((1-w)*am)/len) + ((w*lm)/len)  in float64, with the
 * int64 lengths converted to float64.  variant 1 / 2 are the formulas logged
 * at rescore_result/MLM_PLL/rescore.log:28 and
 * rescore_result/RMBR/BertScore/rescore_mbr_normalize.log:29.
 * ------------------------------------------------------------------------- */
static double combine(double w, double am, double lm, double len, int variant) {
  volatile double one_minus_w = 1.0 - w; /* python: (1-weight) evaluated first */
  volatile double a = one_minus_w * am;
  volatile double l = w * lm;
  if (variant == 0) { a = a / len; l = l / len; }
  else if (variant == 2) { a = a / len; }
  return a + l;
}

void oracle_rescore_scores(const double* am, const double* lm, const int64_t* len,
                           int32_t N, int32_t n_best, double weight, int32_t variant,
                           double* out) {
  for (int64_t i = 0; i < (int64_t)N * n_best; ++i)
    out[i] = combine(weight, am[i], lm[i], (double)len[i], variant);
}

/* get_highest_score_hyp() — rescore.py:55-58: np.argmax(axis=-1) returns the
 * first maximum and treats NaN as the maximum (first NaN wins). */
static int32_t argmax_first(const double* v, int32_t n) {
  int32_t best = 0;
  double bv = v[0];
  if (isnan(bv)) return 0;
  for (int32_t k = 1; k < n; ++k) {
    if (isnan(v[k])) return k;
    if (v[k] > bv) { bv = v[k]; best = k; }
  }
  return best;
}

/* The lambda sweep of find_best_weight — rescore.py:37-43 — with jiwer.cer's
 * numerator expressed through the precomputed per-pair distances:
 * out_edit_sum[w] = sum_u dist[u, argmax_k score_w[u,k]]. */
void oracle_rescore_sweep(const double* am, const double* lm, const int64_t* len,
                          const int32_t* dist, int32_t N, int32_t n_best,
                          const double* weights, int32_t W, int32_t variant,
                          int32_t* out_argmax, int64_t* out_edit_sum) {
  double* s = (double*)malloc(sizeof(double) * (size_t)n_best);
  for (int32_t wi = 0; wi < W; ++wi) {
    int64_t sum = 0;
    for (int32_t u = 0; u < N; ++u) {
      const int64_t base = (int64_t)u * n_best;
      for (int32_t k = 0; k < n_best; ++k)
        s[k] = combine(weights[wi], am[base + k], lm[base + k], (double)len[base + k], variant);
      int32_t a = argmax_first(s, n_best);
      out_argmax[(int64_t)wi * N + u] = a;
      sum += dist[base + a];
    }
    out_edit_sum[wi] = sum;
  }
  free(s);
}

/* ---------------------------------------------------------------------------
 * Masked-copy expansion — MLM_PLL/preprocess.py:9-30 (do_job) on token ids:
 * for a hypothesis of L wordpieces emit L rows
 *   [CLS] t[:m] [MASK] t[m+1:] [SEP],  mask_pos = m+1,  label = t[m]
 * packed back to back (the reference pads them at MLM_PLL/main.py:50-52).
 * ------------------------------------------------------------------------- */
void oracle_expand(const int32_t* hyp_tokens, const int64_t* hyp_off, int32_t n_hyp,
                   int32_t cls_id, int32_t sep_id, int32_t mask_id,
                   int32_t* out_ids, int32_t* out_mask_pos, int32_t* out_labels) {
  int64_t t = 0, c = 0;
  for (int32_t h = 0; h < n_hyp; ++h) {
    const int32_t* tok = hyp_tokens + hyp_off[h];
    int32_t L = (int32_t)(hyp_off[h + 1] - hyp_off[h]);
    for (int32_t m = 0; m < L; ++m) {
      out_ids[t++] = cls_id;
      for (int32_t p = 0; p < L; ++p) out_ids[t++] = (p == m) ? mask_id : tok[p];
      out_ids[t++] = sep_id;
      out_mask_pos[c] = m + 1;
      out_labels[c] = tok[m];
      ++c;
    }
  }
}
