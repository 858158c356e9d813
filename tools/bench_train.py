"""MLM fine-tuning throughput: masked-copy rows per second of a training step on
one B200 next to the reference's own loop on the box's host cores.  Not the driver's bench (bench.py
measures the north-star scoring metric); one JSON line on stdout.

    python tools/bench_train.py [--batch 32] [--steps 20] [--warmup 3] [--layers 12] [--cpu-batches 2]

Workload: bert-base-chinese shape, random init (seed 10), for_training rows (MLM_PLL/preprocess.py:36-44)
of AISHELL-1-test-shaped reference sentences, batches of `--batch` consecutive rows (train.yaml:19-21:
shuffle False, batch_size 32), dropout 0.1 on both arms, lr 1e-5.  GPU arm: engine.MlmTrainer.step per
batch through pllb_train_step_host (H2D of ids / labels and the D2H of the loss inside, one host
sync per batch like the reference's loss.item()).  CPU arm: baseline/_ref/MLM_PLL/main.py (unmodified)
run_one_epoch(train_mode=True) on transformers.BertForMaskedLM, all host threads, `--cpu-batches`
batches.  FLOPs per row-token: 6 x (dense parameters touched) = 3 x the forward's 2*(NL*(4H^2+2HI) + H^2 + HV).
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--cpu-batches", type=int, default=2)
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import importlib
    from asr_rescoring_b200 import engine, synth
    from asr_rescoring_b200.synth import CLS_ID, MASK_ID, SEP_ID
    import bench
    drop_in = importlib.import_module("asr_rescoring_b200.MLM_PLL.main")

    def rows_of(tokens, utt):          # the for_training rows of MLM_PLL/preprocess.py:9-30 (token ids instead of text)
        return [{"utt_id": utt, "hyp_id": None, "input_ids": [CLS_ID] + tokens[:m] + [MASK_ID] + tokens[m + 1:] + [SEP_ID],
                 "attention_masks": [1] * (len(tokens) + 2), "mask_pos": m + 1, "labels": [CLS_ID] + tokens + [SEP_ID]}
                for m in range(len(tokens))]
    cfg = dict(synth.BERT_BASE_CHINESE, num_layers=args.layers)
    sd = synth.random_init_state_dict(cfg, 10)
    need = (args.steps + args.warmup) * args.batch
    nb = synth.make_nbest(max(need // 10, 8), 1, seed=0)
    tok, off = nb.packed_tokens()
    rows = [r for i in range(len(off) - 1) for r in rows_of([int(t) for t in tok[off[i]:off[i + 1]]], f"utt{i}")]
    assert len(rows) >= need, (len(rows), need)
    batches = [rows[i * args.batch:(i + 1) * args.batch] for i in range(args.steps + args.warmup)]
    arrays = []
    for b in batches:
        arrays.append(drop_in.pad_batch(b))
    max_rows = max(a[0].size for a in arrays)
    max_seq = max(a[0].shape[1] for a in arrays)
    H, I, V, NL = cfg["hidden"], cfg["intermediate"], cfg["vocab"], cfg["num_layers"]
    flops_per_row_token = 3 * 2 * (NL * (4 * H * H + 2 * H * I) + H * H + H * V)
    with engine.MlmTrainer(sd, cfg, lr=1e-5, hidden_dropout=0.1, attention_dropout=0.1, seed=10, max_rows=max_rows,
                           max_seq=max_seq) as tr:
        tr.reset_optimizer(1e-5)
        losses = []
        for a in arrays[:args.warmup]:
            tr.step(*a, mode=1)
        l0 = tr.kernel_launches()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in arrays[args.warmup:]:
            losses.append(tr.step(*a, mode=1))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = (tr.kernel_launches() - l0) / args.steps
        replays = tr.graph_replays()
        ws = int(tr._lib.pllb_train_workspace_bytes(tr._h))
    tokens = sum(a[0].size for a in arrays[args.warmup:])
    gpu = dict(rows_per_s=args.steps * args.batch / dt, ms_per_step=1e3 * dt / args.steps, row_tokens_per_step=tokens / args.steps,
               tflops=flops_per_row_token * tokens / dt / 1e12, launches_per_step=launches, first_loss=losses[0], last_loss=losses[-1],
               workspace_gb=ws / 1e9, graph_replays_of_all_steps=f"{replays} of {args.steps + args.warmup}",
               distinct_batch_shapes=len({a[0].shape for a in arrays}))
    cpu = None
    if args.cpu_batches > 0:
        ref_main = bench._load_reference_module("ref_mlm_pll_main", os.path.join("MLM_PLL", "main.py"), "MLM_PLL")
        threads = bench.host_threads()
        torch.set_num_threads(threads)
        if ref_main is not None:
            from transformers import BertConfig, BertForMaskedLM
            hf = BertForMaskedLM(BertConfig(vocab_size=V, hidden_size=H, num_hidden_layers=NL, num_attention_heads=cfg["num_heads"],
                                            intermediate_size=I, max_position_embeddings=cfg["max_position"],
                                            type_vocab_size=cfg["type_vocab"], layer_norm_eps=cfg["ln_eps"], pad_token_id=0,
                                            hidden_act="gelu"))
            hf.load_state_dict(sd, strict=False)
            sample = [r for b in batches[args.warmup:args.warmup + args.cpu_batches] for r in b]
            loader = ref_main.set_dataloader(SimpleNamespace(shuffle=False, batch_size=args.batch, num_worker=0),
                                             ref_main.MyDataset(sample), False)
            t0 = time.perf_counter()
            loss = ref_main.run_one_epoch(config=SimpleNamespace(device="cpu", lr=1e-5), model=hf, dataloader=loader,
                                          output_score=None, train_mode=True, do_scoring=False)
            cdt = time.perf_counter() - t0
            cpu = dict(rows_per_s=len(sample) / cdt, seconds=cdt, cores=threads, kind="reference", epoch_loss=loss,
                       sample=f"{args.cpu_batches} batches of {args.batch} rows through baseline/_ref/MLM_PLL/main.py "
                              f"run_one_epoch(train_mode=True) (unmodified; includes the AdamW construction of :76)")
    os.dup2(real_stdout, 1)
    print(json.dumps({"metric": "MLM fine-tuning masked-copy rows/sec", "unit": "rows/s", "value": gpu["rows_per_s"],
                      "config": {"workload": f"bert-base-chinese shape, {NL} layers, batch {args.batch} rows, AISHELL-1-shaped sentences, "
                                             "dropout 0.1, AdamW lr 1e-5", "steps": args.steps, "warmup": args.warmup},
                      "dtype": "bf16 operands, fp32 master weights / gradients / moments", "gpu": gpu, "cpu_baseline": cpu,
                      "ratio": gpu["rows_per_s"] / cpu["rows_per_s"] if cpu else None}), flush=True)


if __name__ == "__main__":
    main()
