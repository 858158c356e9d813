"""CPU restatement (plain torch fp32 + autograd) of the reference's MLM fine-tuning path.

TEST INFRASTRUCTURE ONLY — imported by tests/ (and oracle/make_golden_train.py); never by the
product path.

What it restates (file:line relative to the reference tree):
  * rows for training ............... MLM_PLL/preprocess.py:9-30,36-44 (do_job over ref_text.json:
    one masked copy per token; `labels` is the WHOLE unmasked sequence [CLS] t [SEP])
  * zero-padded batches ............. MLM_PLL/main.py:28-54 (collate; labels are padded with 0 =
    [PAD], not -100)
  * the training loop ............... MLM_PLL/main.py:73-99,109-114 (run_one_epoch, train_mode):
    a fresh torch.optim.AdamW(lr) per epoch (:76; betas 0.9/0.999, eps 1e-8, weight_decay 0.01 on
    EVERY parameter), per batch forward -> loss -> backward -> step -> zero_grad, epoch loss =
    mean of the batch losses
  * the loss ........................ transformers BertForMaskedLM.forward with labels
    (models/bert/modeling_bert.py:975-983): CrossEntropyLoss() over ALL B*T positions — [CLS],
    [SEP] and the pad positions (label 0) included, because nothing is set to -100
  * the epoch driver ................ MLM_PLL/main.py:117-161 (mlm_finetune_bert): train epoch,
    dev epoch (train_mode=False, do_scoring=False: the same loss, no update), checkpoint.

Dropout.  The reference trains in model.train() mode with the checkpoint's dropout (0.1 on the
hidden states and the attention probabilities), drawn from torch's global RNG; no independent
implementation can reproduce those draws, so parity is DEFINED at dropout 0 (the golden run of
oracle/make_golden_train.py builds the model with both probabilities set to 0).  The decoder weight
is tied to the word embeddings and the decoder bias to cls.predictions.bias (one parameter each).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import pll_oracle

TIED = {"cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight",
        "cls.predictions.decoder.bias": "cls.predictions.bias"}


def training_rows(token_lists, utt_ids=None, cls_id=101, sep_id=102, mask_id=103) -> List[dict]:
    """MLM_PLL/preprocess.py:58-60: for_training rows of reference sentences (hyp_id None)."""
    rows = []
    for i, toks in enumerate(token_lists):
        rows += pll_oracle.expand_rows(toks, utt_ids[i] if utt_ids else f"utt{i}", None, cls_id, sep_id, mask_id)
    return rows


def parameters(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The distinct trainable tensors of a BertForMaskedLM state_dict (tied entries and the
    position_ids buffer dropped), cloned as fp32 leaves."""
    out = {}
    for k, v in sd.items():
        if k in TIED or k.endswith("position_ids"):
            continue
        out[k] = v.detach().clone().float().requires_grad_(True)
    return out


def _full(params):
    sd = dict(params)
    for k, src in TIED.items():
        sd[k] = params[src]
    return sd


def batch_loss(params, cfg, ids, am, lab) -> torch.Tensor:
    """CrossEntropyLoss()(logits.view(-1, V), labels.view(-1)) — modeling_bert.py:979-983."""
    # padding_idx: the word-embedding LOOKUP sends no gradient to row 0 ([PAD], BertConfig.pad_token_id);
    # the tied decoder still does
    logits = pll_oracle.bert_mlm_logits(_full(params), cfg, ids, am, padding_idx=0)
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), lab.view(-1))


def loss_and_grads(sd, cfg, rows: List[dict]):
    """One batch: (loss, {name: grad}) with the parameters left untouched."""
    params = parameters(sd)
    ids, am, lab, *_ = pll_oracle.collate(rows)
    loss = batch_loss(params, cfg, ids, am, lab)
    loss.backward()
    return loss.item(), {k: p.grad.detach().clone() for k, p in params.items()}


def run_one_epoch(params, cfg, rows: List[dict], batch_size: int, lr: float, train_mode: bool,
                  batch_losses: list | None = None) -> float:
    """MLM_PLL/main.py:73-114 with do_scoring=False.  `params` (from parameters()) is updated in
    place when train_mode."""
    opt = torch.optim.AdamW(list(params.values()), lr=lr) if train_mode else None
    total, n = 0.0, 0
    for s in range(0, len(rows), batch_size):
        ids, am, lab, *_ = pll_oracle.collate(rows[s:s + batch_size])
        with torch.set_grad_enabled(train_mode):
            loss = batch_loss(params, cfg, ids, am, lab)
            if train_mode:
                loss.backward()
                opt.step()
                opt.zero_grad()
        total += loss.item()
        n += 1
        if batch_losses is not None:
            batch_losses.append(loss.item())
    return total / n


def state_dict_of(params) -> Dict[str, torch.Tensor]:
    return {k: v.detach().clone() for k, v in _full(params).items()}
