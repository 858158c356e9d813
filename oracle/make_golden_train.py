"""Generate tests/golden/train_golden.json by running the UNMODIFIED reference training loop.

    python oracle/make_golden_train.py        (needs /root/reference; not run on the GPU box)

MLM_PLL/main.py is imported by path (ruamel.yaml shimmed by PyYAML, the only change) and its own
set_dataloader + run_one_epoch(train_mode=True, do_scoring=False) (:57-114) drive a
transformers.BertForMaskedLM carrying synth.random_init_state_dict weights, exactly as
mlm_finetune_bert does (:117-161): per epoch one training pass (fresh AdamW, :76), one dev pass
(train_mode=False).  The model is built with hidden_dropout_prob = attention_probs_dropout_prob = 0:
the reference's dropout draws come from torch's global RNG and cannot be reproduced by any other
implementation, so parity is defined without them (oracle/train_oracle.py header).

Recorded per case: the epoch losses the reference returns, the loss and per-tensor gradient
summaries of the first batch (HF autograd), and per-tensor summaries of the final weights.  A tensor
summary is (L2 norm, dot with a fixed pseudo-random +-1 vector, first 4 values) — enough to pin an
implementation without committing 14 MB of weights.  Also asserts that oracle/train_oracle.py
(the restatement the GPU tests compare with) agrees with the reference run.
"""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import pll_oracle, train_oracle  # noqa: E402
from oracle.make_golden import REF, GOLD, _import_by_path, _install_shims  # noqa: E402
from asr_rescoring_b200 import synth  # noqa: E402


def summary(t: torch.Tensor) -> dict:
    v = t.detach().double().reshape(-1)
    g = torch.Generator().manual_seed(1234 + v.numel() % 977)
    sign = torch.randint(0, 2, (v.numel(),), generator=g, dtype=torch.int64).double() * 2 - 1
    return dict(norm=float(v.norm()), proj=float((v * sign).sum()), head=[float(x) for x in v[:4]])


def training_case(cfg, seed, perturb, n_train, n_dev, data_seed):
    """Rows as MLM_PLL/preprocess.py's for_training jobs make them, from synthetic reference sentences."""
    nb = synth.make_nbest(n_train + n_dev, 1, seed=data_seed)
    tok, off = nb.packed_tokens(cfg["vocab"])
    lists = [[int(t) for t in tok[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
    return lists[:n_train], lists[n_train:]


def hf_model(cfg, sd):
    from transformers import BertConfig, BertForMaskedLM
    hf = BertForMaskedLM(BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"],
                                    num_hidden_layers=cfg["num_layers"], num_attention_heads=cfg["num_heads"],
                                    intermediate_size=cfg["intermediate"], max_position_embeddings=cfg["max_position"],
                                    type_vocab_size=cfg["type_vocab"], layer_norm_eps=cfg["ln_eps"],
                                    pad_token_id=0, hidden_act="gelu", hidden_dropout_prob=0.0,
                                    attention_probs_dropout_prob=0.0))
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    return hf


def main():
    _install_shims()
    ref_main = _import_by_path("ref_mlm_pll_main", os.path.join(REF, "MLM_PLL", "main.py"), os.path.join(REF, "MLM_PLL"))
    cases = []
    specs = [
        # name, cfg, weight seed, perturb, train sentences, dev sentences, data seed, lr, epochs, batch
        ("tiny_perturbed", synth.BERT_TINY, 10, True, 12, 4, 3, 1e-3, 2, 32),
        ("tiny_lr1e-5", synth.BERT_TINY, 11, True, 8, 3, 4, 1e-5, 1, 32),
    ]
    for name, cfg, seed, perturb, n_train, n_dev, data_seed, lr, epochs, bs in specs:
        sd = synth.random_init_state_dict(cfg, seed, perturb)
        train_lists, dev_lists = training_case(cfg, seed, perturb, n_train, n_dev, data_seed)
        train_rows = train_oracle.training_rows(train_lists)
        dev_rows = train_oracle.training_rows(dev_lists)
        # ---- first-batch loss + gradients (HF autograd; the forward the reference calls, main.py:89-94)
        hf = hf_model(cfg, sd).train()
        ids, am, lab, *_ = ref_main.collate(train_rows[:bs])
        out = hf(input_ids=ids, attention_mask=am, labels=lab, return_dict=True)
        out.loss.backward()
        grads = {k: p.grad for k, p in hf.named_parameters()}
        first_loss = float(out.loss)
        o_loss, o_grads = train_oracle.loss_and_grads(sd, cfg, train_rows[:bs])
        assert abs(o_loss - first_loss) < 1e-5, (o_loss, first_loss)
        for k, g in grads.items():
            # (the key-bias gradient is identically zero in exact arithmetic — softmax ignores a per-row
            # constant — so both sides hold rounding noise there: absolute floor)
            diff = float((o_grads[k] - g).norm())
            assert diff < 2e-4 * float(g.norm()) + 1e-7, (k, diff, float(g.norm()))
        # ---- the reference's own epochs
        hf = hf_model(cfg, sd)
        conf = SimpleNamespace(device="cpu", lr=lr)
        dl_conf = SimpleNamespace(shuffle=False, batch_size=bs, num_worker=0)
        train_loader = ref_main.set_dataloader(dl_conf, ref_main.MyDataset(train_rows), False)
        dev_loader = ref_main.set_dataloader(dl_conf, ref_main.MyDataset(dev_rows), True)
        train_losses, dev_losses = [], []
        params = train_oracle.parameters(sd)
        for _ in range(epochs):
            train_losses.append(ref_main.run_one_epoch(config=conf, model=hf, dataloader=train_loader, output_score=None,
                                                       train_mode=True, do_scoring=False))
            dev_losses.append(ref_main.run_one_epoch(config=conf, model=hf, dataloader=dev_loader, output_score=None,
                                                     train_mode=False, do_scoring=False))
            o_tr = train_oracle.run_one_epoch(params, cfg, train_rows, bs, lr, True)
            o_dev = train_oracle.run_one_epoch(params, cfg, dev_rows, bs, lr, False)
            assert abs(o_tr - train_losses[-1]) < 2e-4 and abs(o_dev - dev_losses[-1]) < 2e-4, \
                (o_tr, train_losses[-1], o_dev, dev_losses[-1])
        final = hf.state_dict()
        o_final = train_oracle.state_dict_of(params)
        worst = max(float((o_final[k] - v).abs().max()) for k, v in final.items() if k in o_final)
        print(f"train[{name}]: {len(train_rows)} rows, first loss {first_loss:.5f}, epochs {train_losses} dev {dev_losses}; "
              f"restatement vs reference: max |dw| {worst:.2e}")
        # Adam moves every weight by ~lr per step whatever the gradient's size, so fp32 summation-order noise
        # on near-zero gradients shows up as a fraction of lr * steps
        n_steps = epochs * ((len(train_rows) + bs - 1) // bs)
        assert worst < 0.05 * lr * n_steps, (worst, lr, n_steps)
        keys = [k for k in final if not k.endswith("position_ids")]
        cases.append(dict(name=name, cfg=cfg, seed=seed, perturb=perturb, lr=lr, epochs=epochs, batch_size=bs,
                          train_tokens=train_lists, dev_tokens=dev_lists, first_batch_loss=first_loss,
                          first_batch_grads={k: summary(g) for k, g in grads.items()},
                          train_loss=train_losses, dev_loss=dev_losses,
                          final_weights={k: summary(final[k]) for k in keys},
                          final_minus_init={k: summary(final[k] - sd[k]) for k in keys if k in sd},
                          restatement_vs_reference_max_abs=worst))
    json.dump(dict(generator="oracle/make_golden_train.py",
                   reference="MLM_PLL/main.py set_dataloader + run_one_epoch(train_mode=True) (unmodified), dropout 0, "
                             f"transformers {__import__('transformers').__version__}, torch {torch.__version__}",
                   cases=cases),
              open(os.path.join(GOLD, "train_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
