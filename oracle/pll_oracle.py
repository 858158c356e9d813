"""CPU restatement (plain torch fp32) of the reference's MLM-PLL scoring path.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py; never by the product path.

What it restates (file:line relative to the reference tree):
  * masked-copy expansion ............ MLM_PLL/preprocess.py:9-30  (do_job)
  * zero-padded batches of 32 rows ... MLM_PLL/main.py:28-54,57-70 (collate, set_dataloader)
  * the scoring loop ................. MLM_PLL/main.py:73-114     (run_one_epoch, do_scoring)
  * BertForMaskedLM.forward .......... third-party `transformers` (unpinned by the
    reference; 5.5.0 in this image): models/bert/modeling_bert.py:72-112
    (embeddings), :168-207 (self-attention, scaling dh**-0.5, additive padding
    mask), :294-298, :339-342, :352-356 (post-LN blocks, erf GELU), :481-501
    (MLM head: dense+GELU+LayerNorm, decoder tied to the word embeddings).

Pinning: oracle/make_golden.py runs the UNMODIFIED reference
(/root/reference/MLM_PLL/main.py set_dataloader + run_one_epoch on a
transformers.BertForMaskedLM carrying the same state_dict) in the build
container, checks this restatement against it (max |dPLL| < 2e-4) and commits
the reference's outputs as tests/golden/pll_golden.json.  The reference tree
itself holds no PLL golden values, so that live run is the pin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def expand_rows(tokens: Sequence[int], utt_id, hyp_id, cls_id=101, sep_id=102, mask_id=103) -> List[dict]:
    """MLM_PLL/preprocess.py:9-30 on token ids (the tokenizer call is upstream)."""
    rows = []
    toks = list(tokens)
    for m in range(len(toks)):
        rows.append({
            "utt_id": utt_id,
            "hyp_id": hyp_id,
            "input_ids": [cls_id] + toks[:m] + [mask_id] + toks[m + 1:] + [sep_id],
            "attention_masks": [1] * (len(toks) + 2),
            "mask_pos": m + 1,
            "labels": [cls_id] + toks + [sep_id],
        })
    return rows


def _layer_norm(x, g, b, eps):
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


def bert_mlm_logits(sd: Dict[str, torch.Tensor], cfg: dict, input_ids: torch.Tensor,
                    attention_mask: torch.Tensor, upto_layer: int | None = None,
                    return_hidden: bool = False, padding_idx: int | None = None) -> torch.Tensor:
    """fp32 BertForMaskedLM.forward(...).logits — modeling_bert.py:944-987.
    padding_idx: nn.Embedding(..., padding_idx=pad_token_id) (:75) — same forward; under autograd the
    lookup contributes no gradient to that row (training oracle only)."""
    B, T = input_ids.shape
    H, NH = cfg["hidden"], cfg["num_heads"]
    dh = H // NH
    eps = cfg.get("ln_eps", 1e-12)
    pos = torch.arange(T)
    # BertEmbeddings (:72-112): word + token_type(0) + position -> LayerNorm
    x = (F.embedding(input_ids, sd["bert.embeddings.word_embeddings.weight"], padding_idx=padding_idx)
         + sd["bert.embeddings.token_type_embeddings.weight"][0]
         + sd["bert.embeddings.position_embeddings.weight"][pos])
    x = _layer_norm(x, sd["bert.embeddings.LayerNorm.weight"], sd["bert.embeddings.LayerNorm.bias"], eps)
    add_mask = (1.0 - attention_mask.to(x.dtype))[:, None, None, :] * torch.finfo(x.dtype).min
    n_layers = cfg["num_layers"] if upto_layer is None else upto_layer
    for i in range(n_layers):
        p = f"bert.encoder.layer.{i}."
        q = F.linear(x, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"])
        k = F.linear(x, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"])
        v = F.linear(x, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"])
        q = q.view(B, T, NH, dh).transpose(1, 2)
        k = k.view(B, T, NH, dh).transpose(1, 2)
        v = v.view(B, T, NH, dh).transpose(1, 2)
        s = q @ k.transpose(-1, -2) * (dh ** -0.5) + add_mask
        ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, H)
        a = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        x = _layer_norm(a + x, sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], eps)
        f = F.gelu(F.linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        o = F.linear(f, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
        x = _layer_norm(o + x, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)
    if return_hidden:
        return x
    # BertLMPredictionHead (:481-501)
    t = F.gelu(F.linear(x, sd["cls.predictions.transform.dense.weight"], sd["cls.predictions.transform.dense.bias"]))
    t = _layer_norm(t, sd["cls.predictions.transform.LayerNorm.weight"], sd["cls.predictions.transform.LayerNorm.bias"], eps)
    return F.linear(t, sd["cls.predictions.decoder.weight"], sd["cls.predictions.bias"])


def collate(batch: List[dict]):
    """MLM_PLL/main.py:28-54: right zero-padding of ids / mask / labels."""
    T = max(len(r["input_ids"]) for r in batch)
    ids = torch.zeros(len(batch), T, dtype=torch.long)
    am = torch.zeros(len(batch), T, dtype=torch.long)
    lab = torch.zeros(len(batch), T, dtype=torch.long)
    for i, r in enumerate(batch):
        n = len(r["input_ids"])
        ids[i, :n] = torch.tensor(r["input_ids"])
        am[i, :n] = torch.tensor(r["attention_masks"])
        lab[i, :n] = torch.tensor(r["labels"])
    return ids, am, lab, [r["utt_id"] for r in batch], [r["hyp_id"] for r in batch], [r["mask_pos"] for r in batch]


@torch.no_grad()
def score_rows(sd, cfg, rows: List[dict], output_score: dict, batch_size: int = 32,
               token_scores: list | None = None) -> dict:
    """run_one_epoch(train_mode=False, do_scoring=True) — MLM_PLL/main.py:73-114.

    Rows are consumed in order in batches of `batch_size` (shuffle=False,
    main.py:58-61); the masked row's logits go through a full-vocab
    log_softmax (:101-102), the label column is picked (:104-105) and the
    fp32 value is added to a Python float (:106-107).
    """
    for s in range(0, len(rows), batch_size):
        ids, am, lab, utt, hyp, mpos = collate(rows[s:s + batch_size])
        logits = bert_mlm_logits(sd, cfg, ids, am)
        tl = logits[range(len(logits)), mpos, :]
        ts = tl.log_softmax(dim=-1)
        mt = lab[range(len(lab)), mpos]
        vals = ts[range(len(ts)), mt].tolist()
        for u, h, v in zip(utt, hyp, vals):
            output_score[u][h] += v
        if token_scores is not None:
            token_scores.extend(vals)
    return output_score


def score_hyps(sd, cfg, hyps: Dict[str, Dict[str, Sequence[int]]], batch_size: int = 32,
               cls_id=101, sep_id=102, mask_id=103, token_scores: list | None = None) -> dict:
    """pll_bert_scoring's data flow (MLM_PLL/main.py:164-203) from token ids:
    expand -> skeleton {utt:{hyp:0}} (:189-193) -> score."""
    rows = []
    out = {}
    for u, hs in hyps.items():
        out[u] = {}
        for h, toks in hs.items():
            out[u][h] = 0
            rows += expand_rows(toks, u, h, cls_id, sep_id, mask_id)
    return score_rows(sd, cfg, rows, out, batch_size, token_scores)


@torch.no_grad()
def rescore_bert_scores(sd, cfg, token_lists, linear_w, linear_b, batch_size: int = 32, cls_id=101, sep_id=102):
    """RescoreBert scoring — RescoreBert/model.py:13-21 (BertModel -> last_hidden_state[:, 0, :] ->
    Linear(H, 1).squeeze) over batches padded by pad_sequence as in RescoreBert/main.py:31-75.
    `sd` carries the encoder under the same 'bert.' keys as BertForMaskedLM."""
    rows = [[cls_id] + list(t) + [sep_id] for t in token_lists]
    out = []
    w = torch.as_tensor(linear_w, dtype=torch.float32).reshape(1, -1)
    b = torch.as_tensor([linear_b], dtype=torch.float32)
    for s0 in range(0, len(rows), batch_size):
        batch = rows[s0:s0 + batch_size]
        T = max(len(r) for r in batch)
        ids = torch.zeros(len(batch), T, dtype=torch.long)
        am = torch.zeros(len(batch), T, dtype=torch.long)
        for i, r in enumerate(batch):
            ids[i, :len(r)] = torch.tensor(r)
            am[i, :len(r)] = 1
        hid = bert_mlm_logits(sd, cfg, ids, am, return_hidden=True)
        out += F.linear(hid[:, 0, :], w, b).squeeze(dim=-1).tolist()
    return out


def algorithmic_flops(lengths: Sequence[int], cfg: dict) -> float:
    """F(L) = L*[T*NL*(8H^2+4HI) + NL*4*T^2*H + 2H^2 + 2HV]."""
    H, I, NL, V = cfg["hidden"], cfg["intermediate"], cfg["num_layers"], cfg["vocab"]
    tot = 0.0
    for L in lengths:
        T = L + 2
        tot += L * (T * NL * (8 * H * H + 4 * H * I) + NL * 4 * T * T * H + 2 * H * H + 2 * H * V)
    return tot


def _self_check():  # pragma: no cover - tiny sanity run
    cfg = dict(num_layers=1, hidden=64, num_heads=1, intermediate=128, vocab=200, max_position=64)
    assert math.isfinite(algorithmic_flops([3, 4], cfg))


if __name__ == "__main__":  # pragma: no cover
    _self_check()
