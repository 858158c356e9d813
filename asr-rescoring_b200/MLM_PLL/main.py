"""Drop-in for the reference's ``MLM_PLL/main.py``: PLL scoring and MLM fine-tuning.

    cd MLM_PLL && python main.py --config config/score.yaml
    cd MLM_PLL && python main.py --config config/train.yaml

Same ``--config`` flag, same YAML keys (MLM_PLL/config/score.yaml:1-20, train.yaml:1-25), same output
files (``<output_path>{train,dev,test}_lm.json``: {utt_id: {hyp_id: float}}, indent=4,
ensure_ascii=False; ``<output_path>/checkpoint_<epoch>.pth`` and ``<output_path>/loss.json`` for
training).  Differences, each deliberate (SURVEY.md §8a "quirks"):

* all three splits are scored — the reference's dev/test blocks sit inside a string
  literal (MLM_PLL/main.py:205-238) although rescore.py consumes their outputs;
* the arithmetic runs in libpllb200.so on a B200 (no padding, no [copies x T x V] logits,
  no per-batch host syncs); there is no CPU fallback;
* ``task: training`` (MLM_PLL/main.py:117-161) runs on the same library (engine.MlmTrainer: forward,
  backward and AdamW as CUDA kernels).  The reference starts from the hub checkpoint
  (``from_pretrained(config.model.bert)``, :128); offline the starting weights come from
  ``model.pretrained_path`` (a state_dict file) or ``model.random_init_seed``.  Dropout masks come
  from the library's own counter-based generator, not torch's RNG;
* data files may be the reference's row-list JSON (MLM_PLL/preprocess.py:9-30), a
  ``hyps_text.json`` ({utt: {hyp: str}}, tokenised here) or the compact packed JSON that
  our ``preprocess.py`` writes — the O(sum L^2) row list is never needed;
* optional keys (absent from the reference YAMLs, all defaulted): ``model.vocab_path``,
  ``model.random_init_seed``, ``model.operand_dtype`` ("bf16+fp16head" default: bf16 encoder,
  fp16 MLM head | "bf16" | "fp16": 8x smaller rounding error, ~5 % slower, needed for 24-layer
  encoders at L > ~40), ``max_chunk_tokens``; training: ``model.pretrained_path``,
  ``model.hidden_dropout_prob`` / ``model.attention_probs_dropout_prob`` (0.1 = BertConfig default).
"""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_PARENT = os.path.dirname(os.path.dirname(_HERE))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

from asr_rescoring_b200 import shard, synth  # noqa: E402
from asr_rescoring_b200.engine import DEFAULT_OPERAND_DTYPE, MlmTrainer, PllScorer  # noqa: E402
from asr_rescoring_b200.tokenizer import BertCharTokenizer, SyntheticCharTokenizer, encode_batch  # noqa: E402
from asr_rescoring_b200.util.arg_parser import ArgParser  # noqa: E402
from asr_rescoring_b200.util.saving import json_saving, model_saving  # noqa: E402

MODEL_SHAPES = {"bert-base-chinese": synth.BERT_BASE_CHINESE, "bert-large-shaped": synth.BERT_LARGE_SHAPED,
                "bert-tiny-test": synth.BERT_TINY}


class MyDataset:
    """List wrapper with slice support — MLM_PLL/main.py:17-25."""

    def __init__(self, data_set: List):
        self.data_set = data_set

    def __len__(self):
        return len(self.data_set)

    def __getitem__(self, idx):
        return self.data_set[idx]


def collate(batch: List[dict]):
    """The reference pads rows into [B, Tmax] tensors here (MLM_PLL/main.py:28-54); the
    varlen kernels need no padding, so a batch is just the list of rows."""
    return list(batch)


def pad_batch(batch: List[dict]):
    """The arrays the reference's collate builds (MLM_PLL/main.py:28-54, pad_sequence(batch_first=True)):
    input_ids, attention_mask, labels as int32 [B, Tmax], right-padded with 0."""
    T = max(len(r["input_ids"]) for r in batch)
    ids = np.zeros((len(batch), T), np.int32)
    am = np.zeros((len(batch), T), np.int32)
    lab = np.zeros((len(batch), T), np.int32)
    for i, r in enumerate(batch):
        ids[i, :len(r["input_ids"])] = r["input_ids"]
        am[i, :len(r["attention_masks"])] = r["attention_masks"]
        lab[i, :len(r["labels"])] = r["labels"]
    return ids, am, lab


class RowLoader:
    """Batches of rows in DataLoader order (MLM_PLL/main.py:57-70): sequential, or — shuffle=True — a
    fresh permutation per epoch drawn the way torch's RandomSampler draws it (a generator seeded from
    torch's global RNG, then randperm), so ``torch.manual_seed(config.seed)`` (MLM_PLL/main.py:245-247)
    fixes the order exactly as it does for the reference."""

    def __init__(self, dataset, batch_size: int, shuffle: bool = False):
        self.dataset = dataset
        self.batch_size = max(int(batch_size), 1)
        self.shuffle = bool(shuffle)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        if not self.shuffle:
            for s in range(0, len(self.dataset), self.batch_size):
                yield collate(self.dataset[s:s + self.batch_size])
            return
        import torch
        # torch 2.x DataLoader: the iterator first draws its worker base seed from the global RNG
        # (_BaseDataLoaderIter.__init__), then RandomSampler draws the seed of its own generator
        torch.empty((), dtype=torch.int64).random_()
        g = torch.Generator()
        g.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
        order = torch.randperm(len(self.dataset), generator=g).tolist()
        for s in range(0, len(order), self.batch_size):
            yield collate([self.dataset[i] for i in order[s:s + self.batch_size]])


def set_dataloader(config, dataset, for_scoring=False):
    """MLM_PLL/main.py:57-70.  ``num_worker`` host processes are not needed (no tensors are
    built on the host); the key is still read so a YAML without it fails as before."""
    _ = config.num_worker
    shuffle = False if for_scoring else config.shuffle
    return RowLoader(dataset, config.batch_size, shuffle)


def _iter_rows(dataloader) -> Iterable[dict]:
    for item in dataloader:
        if isinstance(item, dict):
            yield item
        else:
            for row in item:
                yield row


def _run_loss_epoch(config, model: MlmTrainer, dataloader, train_mode: bool, output_score=None, do_scoring=False):
    """run_one_epoch on a trainer (MLM_PLL/main.py:73-114): the mean of the batch losses; train_mode adds
    backward + AdamW per batch, with a fresh optimizer per epoch (:76).  do_scoring additionally adds, for
    every row, log_softmax(logits[mask_pos])[labels[mask_pos]] of THIS forward pass (dropout active when
    train_mode, as in the reference) to output_score and returns it instead of the loss (:100-107,111-114)."""
    if not isinstance(model, MlmTrainer):
        raise TypeError("the training / loss pass needs an engine.MlmTrainer (build_trainer), not a PllScorer")
    if train_mode:
        model.reset_optimizer(config.lr)
    epoch_loss, n_batches = 0.0, 0
    for batch in dataloader:
        ids, am, lab = pad_batch(batch)
        epoch_loss += model.step(ids, am, lab, mode=1 if train_mode else 0)
        n_batches += 1
        if do_scoring:
            T = ids.shape[1]
            nll = model.row_losses(ids.size)
            for b, row in enumerate(batch):
                output_score[row["utt_id"]][row["hyp_id"]] += float(-nll[b * T + int(row["mask_pos"])])
    if do_scoring:
        return output_score
    return epoch_loss / len(dataloader) if n_batches else 0.0


def run_one_epoch(config, model, dataloader, output_score=None, train_mode=True, do_scoring=False):
    """MLM_PLL/main.py:73-114.  do_scoring=False: the fine-tuning / dev-loss pass (model is an
    engine.MlmTrainer), returns the epoch loss.  Scoring branch: for every row (one masked copy) add
    log_softmax(logits[mask_pos])[labels[mask_pos]] to output_score[utt_id][hyp_id].
    Rows of one hypothesis are consecutive (preprocess.py emits them so); a row list cut in
    the middle of a hypothesis by ``num_of_data`` adds only the rows present, like the
    reference."""
    if not do_scoring or train_mode or isinstance(model, MlmTrainer):
        # (scoring while training, and scoring through a trainer, go row by row through the padded-batch forward)
        return _run_loss_epoch(config, model, dataloader, train_mode, output_score, do_scoring)
    groups: List[Tuple[str, str, List[int], List[int]]] = []   # utt, hyp, tokens, mask positions present
    last_key = None
    for row in _iter_rows(dataloader):
        key = (row["utt_id"], row["hyp_id"], tuple(row["labels"]))
        if key != last_key:
            groups.append((row["utt_id"], row["hyp_id"], list(row["labels"][1:-1]), []))
            last_key = key
        groups[-1][3].append(int(row["mask_pos"]) - 1)
    if not groups:
        return output_score
    off = np.zeros(len(groups) + 1, np.int64)
    np.cumsum([len(g[2]) for g in groups], out=off[1:])
    tokens = np.fromiter((t for g in groups for t in g[2]), np.int32, int(off[-1]))
    pll, tok_logp = model.score_packed(tokens, off, return_token_logp=True)
    for i, (u, h, toks, present) in enumerate(groups):
        if len(present) == len(toks) and present == list(range(len(toks))):
            output_score[u][h] += float(pll[i])          # device sum, same order, in double
        else:
            for m in present:
                output_score[u][h] += float(tok_logp[off[i] + m])
    return output_score


# ---------------------------------------------------------------------------- data
def _tokenizer(config):
    """The reference always tokenises with BertTokenizer.from_pretrained("bert-base-chinese")
    (MLM_PLL/preprocess.py:34).  Offline that vocabulary comes from ``model.vocab_path`` (or
    $PLLB_VOCAB).  The synthetic char->id map is only meaningful for random-init weights: it is
    allowed with ``model.random_init_seed`` and no checkpoint, and refused otherwise — a real
    checkpoint scored with made-up ids would still write plausible-looking *_lm.json files."""
    vocab_path = getattr(config.model, "vocab_path", None) or os.environ.get("PLLB_VOCAB")
    if vocab_path:
        return BertCharTokenizer(vocab_path)
    ckpt = getattr(config, "checkpoint_path", None)
    if (ckpt and os.path.exists(ckpt)) or getattr(config.model, "random_init_seed", None) is None:
        raise FileNotFoundError("text input (hyps_text.json) needs the checkpoint's vocabulary: set model.vocab_path "
                                "(or $PLLB_VOCAB) to its vocab.txt; the synthetic tokenizer is only allowed with "
                                "model.random_init_seed and no checkpoint")
    return SyntheticCharTokenizer()


def load_split(path: str, num_of_data: int, config) -> Tuple[Optional[list], Dict[str, Dict[str, List[int]]]]:
    """Returns (rows or None, {utt: {hyp: token ids}}) for any of the three file formats."""
    data = json.load(open(path, "r", encoding="utf-8"))
    if isinstance(data, list):                       # reference row list
        return data[:num_of_data], {}
    if isinstance(data, dict) and data.get("format") == "pllb-packed-v1":
        kind = data.get("tokenizer", "unknown")
        ckpt = getattr(config, "checkpoint_path", None)
        if kind == "synthetic" and ckpt and os.path.exists(ckpt):
            raise ValueError(f"{path} was tokenised with the synthetic char->id map; it cannot be scored with the "
                             f"checkpoint {ckpt!r} (re-run preprocess.py with PLLB_VOCAB=<vocab.txt>)")
        hyps: Dict[str, Dict[str, List[int]]] = {}
        off, tok = data["offsets"], data["tokens"]
        for i, (u, h) in enumerate(zip(data["utt_id"], data["hyp_id"])):
            hyps.setdefault(u, {})[h] = tok[off[i]:off[i + 1]]
        return None, hyps
    tk = _tokenizer(config)                          # hyps_text.json: one batched tokenizer call
    ids, off = encode_batch(tk, [s for hs in data.values() for s in hs.values()])
    ids, off = ids.tolist(), off.tolist()
    hyps, i = {}, 0
    for u, hs in data.items():
        hyps[u] = {}
        for h in hs:
            hyps[u][h] = ids[off[i]:off[i + 1]]
            i += 1
    return None, hyps


def skeleton_from_rows(rows: list) -> dict:
    """{utt: {hyp: 0}} exactly as MLM_PLL/main.py:189-193 (keys on hyp_id == "hyp_1")."""
    out: dict = {}
    for data in rows:
        if data["hyp_id"] == "hyp_1":
            out[data["utt_id"]] = {}
        out[data["utt_id"]][data["hyp_id"]] = 0
    return out


def build_scorer(config, device: int = 0) -> PllScorer:
    """Replaces BertForMaskedLM.from_pretrained + load_state_dict + .to(device)
    (MLM_PLL/main.py:184-187): the checkpoint is the bare state_dict written by
    util/saving.py:7-11; the model shape is inferred from it."""
    import torch
    seed = getattr(config.model, "random_init_seed", None)
    ckpt = config.checkpoint_path
    if ckpt and os.path.exists(ckpt):
        sd = torch.load(ckpt, map_location="cpu")
        cfg = synth.config_from_state_dict(sd)
    elif seed is not None:
        cfg = MODEL_SHAPES[config.model.bert]
        sd = synth.random_init_state_dict(cfg, int(seed))
    else:
        raise FileNotFoundError(f"checkpoint_path {ckpt!r} not found and model.random_init_seed not set")
    return PllScorer(sd, cfg, device=device, max_chunk_tokens=int(getattr(config, "max_chunk_tokens", 0) or 0),
                     operand_dtype=str(getattr(config.model, "operand_dtype", DEFAULT_OPERAND_DTYPE)))


def score_split(config, model: PllScorer, path: str) -> dict:
    rows, hyps = load_split(path, config.num_of_data, config)
    rank, world, _ = shard.dist_env()
    if rows is not None:
        output_score = skeleton_from_rows(rows)
        if world == 1:
            loader = set_dataloader(config.dataloader, MyDataset(rows), True)
            return run_one_epoch(config=config, model=model, dataloader=loader, output_score=output_score,
                                 train_mode=False, do_scoring=True)
        return _score_rows_sharded(config, model, rows, output_score, rank, world)
    utts = list(hyps.keys())
    parts = shard.lpt_partition(shard.utterance_costs([[len(t) for t in hyps[u].values()] for u in utts]), world)
    mine = [utts[i] for i in parts[rank]]
    base = np.zeros(len(utts) + 1, np.int64)
    np.cumsum([len(hyps[u]) for u in utts], out=base[1:])
    local = model.score_hyps({u: hyps[u] for u in mine})
    vals = np.array([float(v) for u in mine for v in local[u].values()], np.float64)
    idx = np.array([base[i] + k for i in parts[rank] for k in range(len(hyps[utts[i]]))], np.int64)
    full = shard.gather_scores(vals, idx, int(base[-1]))
    out, j = {}, 0
    for u in utts:
        out[u] = {}
        for h, toks in hyps[u].items():
            out[u][h] = float(full[j]) if len(toks) > 0 else 0
            j += 1
    return out


def _score_rows_sharded(config, model: PllScorer, rows: list, output_score: dict, rank: int, world: int) -> dict:
    """Row-list input on several GPUs: rows of one utterance are consecutive (preprocess.py emits
    them so); the utterance segments are LPT-partitioned by their token count, every rank runs
    run_one_epoch on its own rows, and the per-(utt, hyp) sums are gathered.  A hypothesis with
    no rows keeps the int 0 of the skeleton (MLM_PLL/main.py:189-193), as on one GPU."""
    seg_start, seg_utt = [], []
    for i, row in enumerate(rows):
        if not seg_utt or row["utt_id"] != seg_utt[-1]:
            seg_start.append(i)
            seg_utt.append(row["utt_id"])
    seg_start.append(len(rows))
    if len(set(seg_utt)) != len(seg_utt):
        raise ValueError("row list has non-consecutive rows of one utterance; cannot shard it by utterance")
    costs = [sum(len(r["input_ids"]) for r in rows[seg_start[j]:seg_start[j + 1]]) for j in range(len(seg_utt))]
    mine = shard.lpt_partition(costs, world)[rank]
    my_rows = [r for j in mine for r in rows[seg_start[j]:seg_start[j + 1]]]
    local = {seg_utt[j]: dict(output_score[seg_utt[j]]) for j in mine}
    loader = set_dataloader(config.dataloader, MyDataset(my_rows), True)
    run_one_epoch(config=config, model=model, dataloader=loader, output_score=local, train_mode=False, do_scoring=True)
    keys = [(u, h) for u, hs in output_score.items() for h in hs]
    index = {k: i for i, k in enumerate(keys)}
    idx = np.array([index[(u, h)] for u, hs in local.items() for h in hs], np.int64)
    vals = np.array([float(v) for hs in local.values() for v in hs.values()], np.float64)
    full = shard.gather_scores(vals, idx, len(keys))
    scored = {(r["utt_id"], r["hyp_id"]) for r in rows}
    for (u, h), v in zip(keys, full):
        output_score[u][h] = float(v) if (u, h) in scored else 0
    return output_score


def pll_bert_scoring(config):
    """MLM_PLL/main.py:164-203, for all three splits."""
    rank, world, local = shard.init_process_group() if shard.dist_env()[1] > 1 else (0, 1, 0)
    dev = config.device
    device = local if world > 1 else (int(str(dev).split(":")[1]) if ":" in str(dev) else 0)
    model = build_scorer(config, device)
    for split, path in (("train", config.train_data_path), ("dev", config.dev_data_path), ("test", config.test_data_path)):
        output_score = score_split(config, model, path)
        if rank == 0:
            json_saving(config.output_path + f"{split}_lm.json", output_score)   # string concat as main.py:203
    model.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def build_trainer(config, device: int = 0, max_rows: int = 0, max_seq: int = 0) -> MlmTrainer:
    """Replaces BertForMaskedLM.from_pretrained(config.model.bert).to(device) (MLM_PLL/main.py:128-129).
    The hub is not reachable offline: the starting state_dict comes from ``model.pretrained_path`` or,
    for synthetic runs, ``model.random_init_seed``."""
    import torch
    path = getattr(config.model, "pretrained_path", None)
    seed = getattr(config.model, "random_init_seed", None)
    if path and os.path.exists(path):
        sd = torch.load(path, map_location="cpu")
        cfg = synth.config_from_state_dict(sd)
    elif seed is not None:
        cfg = MODEL_SHAPES[config.model.bert]
        sd = synth.random_init_state_dict(cfg, int(seed))
    else:
        raise FileNotFoundError(f"model.pretrained_path {path!r} not found and model.random_init_seed not set "
                                f"(the reference downloads {config.model.bert!r} from the hub)")
    return MlmTrainer(sd, cfg, device=device, lr=float(config.lr),
                      hidden_dropout=float(getattr(config.model, "hidden_dropout_prob", 0.1)),
                      attention_dropout=float(getattr(config.model, "attention_probs_dropout_prob", 0.1)),
                      seed=int(config.seed or 0), max_rows=max_rows or 4096, max_seq=max_seq or 128)


def mlm_finetune_bert(config):
    """MLM_PLL/main.py:117-161: per epoch one training pass, one dev-loss pass, a checkpoint
    (<output_path>/checkpoint_<epoch>.pth, the bare state_dict) and <output_path>/loss.json."""
    train_set = MyDataset(json.load(open(config.train_data_path, "r", encoding="utf-8")))[:config.num_of_data]
    dev_set = MyDataset(json.load(open(config.dev_data_path, "r", encoding="utf-8")))[:config.num_of_data]
    train_loader = set_dataloader(config.dataloader, train_set, False)
    dev_loader = set_dataloader(config.dataloader, dev_set, True)
    dev = config.device
    longest = max([len(r["input_ids"]) for r in train_set] + [len(r["input_ids"]) for r in dev_set] + [1])
    model = build_trainer(config, int(str(dev).split(":")[1]) if ":" in str(dev) else 0,
                          max_rows=int(config.dataloader.batch_size) * longest, max_seq=longest)
    train_loss_record = [0] * config.epoch
    dev_loss_record = [0] * config.epoch
    for epoch_id in range(1, config.epoch + 1):
        print("Epoch {}/{}".format(epoch_id, config.epoch))
        train_loss_record[epoch_id - 1] = run_one_epoch(config=config, model=model, output_score=None,
                                                        dataloader=train_loader, train_mode=True, do_scoring=False)
        print("epoch ", epoch_id, " train loss: ", train_loss_record[epoch_id - 1])
        dev_loss_record[epoch_id - 1] = run_one_epoch(config=config, model=model, output_score=None,
                                                      dataloader=dev_loader, train_mode=False, do_scoring=False)
        print("epoch ", epoch_id, " dev loss: ", dev_loss_record[epoch_id - 1], "\n")
        model_saving(config.output_path, model.state_dict(), epoch_id)
        json_saving(config.output_path + "/loss.json", {"train": train_loss_record, "dev": dev_loss_record})
    model.close()


if __name__ == "__main__":
    arg_parser = ArgParser()
    config = arg_parser.parse()

    if config.seed is not None:
        import torch
        torch.manual_seed(config.seed)
        torch.cuda.manual_seed(config.seed)

    if config.task == "training":
        mlm_finetune_bert(config)
    elif config.task == "scoring":
        pll_bert_scoring(config)
