"""Host-side driver of libpllb200.so: weight handoff, PLL scoring, Levenshtein, λ-sweep.

PyTorch is used only to own device memory (weights on their way in, caller-visible
outputs) and for the current CUDA stream; all arithmetic happens in the library.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import LayerWeights, ModelDesc, Stats, Weights, check

VARIANTS = {"B": 0, "A": 1, "C": 2, 0: 0, 1: 1, 2: 2}
OPERAND_DTYPES = {"bf16": 0, "fp16": 1, "bf16+fp16head": 2, "bf16+fp16tail": 3}   # pllb_model_desc.operand_dtype


def operand_dtype_code(name: str) -> int:
    """"bf16" | "fp16" | "bf16+fp16head" (bf16 encoder, fp16 MLM head) | "bf16+fp16tail" (bf16 for the
    first half of the encoder layers, fp16 for the second half and the head) | "fp16from:K" (bf16 for
    layers < K, fp16 for layers >= K and the head)."""
    if name.startswith("fp16from:"):
        return 16 + int(name.split(":", 1)[1])
    return OPERAND_DTYPES[name]
# bf16 operands on every encoder GEMM and attention MMA (98.9 % of the FLOPs), IEEE fp16 operands on the
# two MLM-head GEMMs: same speed as all-bf16 (measured), and the largest single rounding site is gone
DEFAULT_OPERAND_DTYPE = "bf16+fp16head"
GEMM_KINDS = ("qkv", "attn_out", "ffn1", "ffn2", "head_transform", "decoder_lse")


def _np_ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def weights_struct(get, num_layers: int, keys):
    """pllb_weights over the tensors ``get(key)`` returns (CUDA fp32, contiguous) for the HF
    BertForMaskedLM state_dict names.  Returns (struct, list keeping the tensors and the layer array
    alive).  The MLM head is optional: a RescoreBert checkpoint (BertModel + Linear(H, 1)) has none."""
    keys = set(keys)
    keep = []

    def dp(key):
        t = get(key)
        keep.append(t)
        return ctypes.c_void_p(t.data_ptr())

    layers = (LayerWeights * max(num_layers, 1))()
    keep.append(layers)
    for i in range(num_layers):
        p = f"bert.encoder.layer.{i}."
        lw = layers[i]
        lw.q_w, lw.q_b = dp(p + "attention.self.query.weight"), dp(p + "attention.self.query.bias")
        lw.k_w, lw.k_b = dp(p + "attention.self.key.weight"), dp(p + "attention.self.key.bias")
        lw.v_w, lw.v_b = dp(p + "attention.self.value.weight"), dp(p + "attention.self.value.bias")
        lw.ao_w, lw.ao_b = dp(p + "attention.output.dense.weight"), dp(p + "attention.output.dense.bias")
        lw.ao_ln_g, lw.ao_ln_b = dp(p + "attention.output.LayerNorm.weight"), dp(p + "attention.output.LayerNorm.bias")
        lw.ff1_w, lw.ff1_b = dp(p + "intermediate.dense.weight"), dp(p + "intermediate.dense.bias")
        lw.ff2_w, lw.ff2_b = dp(p + "output.dense.weight"), dp(p + "output.dense.bias")
        lw.out_ln_g, lw.out_ln_b = dp(p + "output.LayerNorm.weight"), dp(p + "output.LayerNorm.bias")
    w = Weights()
    w.word_emb = dp("bert.embeddings.word_embeddings.weight")
    w.pos_emb = dp("bert.embeddings.position_embeddings.weight")
    w.type_emb = dp("bert.embeddings.token_type_embeddings.weight")
    w.emb_ln_g = dp("bert.embeddings.LayerNorm.weight")
    w.emb_ln_b = dp("bert.embeddings.LayerNorm.bias")
    w.layers = ctypes.cast(layers, ctypes.POINTER(LayerWeights))
    if "cls.predictions.transform.dense.weight" in keys:
        w.head_w = dp("cls.predictions.transform.dense.weight")
        w.head_b = dp("cls.predictions.transform.dense.bias")
        w.head_ln_g = dp("cls.predictions.transform.LayerNorm.weight")
        w.head_ln_b = dp("cls.predictions.transform.LayerNorm.bias")
        w.decoder_w = dp("cls.predictions.decoder.weight" if "cls.predictions.decoder.weight" in keys
                         else "bert.embeddings.word_embeddings.weight")
        w.decoder_b = dp("cls.predictions.bias" if "cls.predictions.bias" in keys else "cls.predictions.decoder.bias")
    return w, keep


class PllScorer:
    """BERT masked-LM pseudo-log-likelihood scorer; stands in for the
    ``BertForMaskedLM`` + ``run_one_epoch(do_scoring=True)`` pair of
    MLM_PLL/main.py:73-114,184-187."""

    def __init__(self, state_dict: Dict[str, "torch.Tensor"], cfg: Optional[dict] = None, device: int = 0,
                 max_chunk_tokens: int = 0, cls_id: int = 101, sep_id: int = 102, mask_id: int = 103,
                 operand_dtype: str = DEFAULT_OPERAND_DTYPE):
        import torch
        from .synth import config_from_state_dict

        self._lib = _lib.load()
        _lib.require_device()
        self._h = ctypes.c_void_p()
        self.cfg = dict(cfg) if cfg is not None else config_from_state_dict(state_dict)
        self.device = device
        dev = torch.device("cuda", device)
        nl = self.cfg["num_layers"]
        w, keep = weights_struct(lambda key: state_dict[key].detach().to(device=dev, dtype=torch.float32).contiguous(),
                                 nl, state_dict.keys())
        self.has_mlm_head = bool(w.head_w)
        d = ModelDesc(nl, self.cfg["hidden"], self.cfg["num_heads"], self.cfg["intermediate"], self.cfg["vocab"],
                      self.cfg["max_position"], float(self.cfg.get("ln_eps", 1e-12)), cls_id, sep_id, mask_id,
                      operand_dtype_code(operand_dtype))
        self.operand_dtype = operand_dtype
        torch.cuda.synchronize(dev)
        check(self._lib.pllb_create(ctypes.byref(self._h), ctypes.byref(d), ctypes.byref(w), int(max_chunk_tokens), device))
        del keep   # the library made its own (bf16 / fp32) copies

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pllb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ scoring
    @staticmethod
    def _check_packed(tokens, offsets):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if offsets.ndim != 1 or len(offsets) < 1:
            raise ValueError("offsets must be a 1-D array of n_hyp+1 entries")
        return offsets

    def score_packed(self, tokens: np.ndarray, offsets: np.ndarray, return_token_logp: bool = False):
        """HOST arrays in, HOST arrays out (H2D + score + D2H inside one C call).
        tokens: int32 wordpiece ids of all hypotheses back to back (no specials);
        offsets: int64[n_hyp+1].  Returns float64[n_hyp] PLLs (and float32 per-token terms)."""
        offsets = self._check_packed(tokens, offsets)
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        n = len(offsets) - 1
        out = np.zeros(n, np.float64)
        tl = np.zeros(int(offsets[-1]), np.float32) if return_token_logp else None
        check(self._lib.pllb_score_host(self._h, _np_ptr(tokens), _np_ptr(offsets), n, _np_ptr(out),
                                        _np_ptr(tl) if tl is not None else None))
        return (out, tl) if return_token_logp else out

    def score_device(self, tokens_cuda, offsets: np.ndarray, out=None, token_logp=None):
        """tokens_cuda: torch int32 CUDA tensor; offsets: host int64.  Asynchronous on the
        current torch stream; returns a torch float64 CUDA tensor [n_hyp]."""
        import torch
        offsets = self._check_packed(None, offsets)
        n = len(offsets) - 1
        assert tokens_cuda.is_cuda and tokens_cuda.dtype == torch.int32 and tokens_cuda.is_contiguous()
        if out is None:
            out = torch.zeros(n, dtype=torch.float64, device=tokens_cuda.device)
        stream = torch.cuda.current_stream(tokens_cuda.device).cuda_stream
        check(self._lib.pllb_score(self._h, ctypes.c_void_p(tokens_cuda.data_ptr()), _np_ptr(offsets), n,
                                   ctypes.c_void_p(out.data_ptr()),
                                   ctypes.c_void_p(token_logp.data_ptr()) if token_logp is not None else None,
                                   ctypes.c_void_p(stream)))
        return out

    def score_cls_packed(self, tokens: np.ndarray, offsets: np.ndarray, linear_w, linear_b: float) -> np.ndarray:
        """Sequence-level scores (RescoreBert/model.py:13-21): every hypothesis runs through the
        encoder once as [CLS] t [SEP]; returns float32[n_hyp] = Linear(H,1)([CLS] state)."""
        offsets = self._check_packed(tokens, offsets)
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        lw = np.ascontiguousarray(np.asarray(linear_w, np.float32).reshape(-1))
        if lw.shape[0] != self.cfg["hidden"]:
            raise ValueError("linear_w must have `hidden` elements")
        n = len(offsets) - 1
        out = np.zeros(n, np.float32)
        check(self._lib.pllb_score_cls_host(self._h, _np_ptr(tokens), _np_ptr(offsets), n, _np_ptr(lw),
                                            float(linear_b), _np_ptr(out)))
        return out

    def score_hyps(self, hyps: Dict[str, Dict[str, Sequence[int]]]) -> Dict[str, Dict[str, float]]:
        """{utt: {hyp: [token ids]}} -> {utt: {hyp: PLL}}; a hypothesis with no tokens keeps
        the int 0 of the reference's skeleton (MLM_PLL/main.py:189-193)."""
        flat, off = [], [0]
        for hs in hyps.values():
            for toks in hs.values():
                flat.extend(toks)
                off.append(len(flat))
        pll = self.score_packed(np.asarray(flat, np.int32), np.asarray(off, np.int64))
        out, i = {}, 0
        for u, hs in hyps.items():
            out[u] = {}
            for h, toks in hs.items():
                out[u][h] = float(pll[i]) if len(toks) > 0 else 0
                i += 1
        return out

    # ------------------------------------------------------------------ parity hooks
    def expand(self, tokens: np.ndarray, offsets: np.ndarray):
        import torch
        offsets = self._check_packed(tokens, offsets)
        L = np.diff(offsets)
        dev = torch.device("cuda", self.device)
        t = torch.from_numpy(np.ascontiguousarray(tokens, np.int32)).to(dev)
        ids = torch.zeros(max(int((L * (L + 2)).sum()), 1), dtype=torch.int32, device=dev)
        mp = torch.zeros(max(int(L.sum()), 1), dtype=torch.int32, device=dev)
        lab = torch.zeros_like(mp)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(self._lib.pllb_expand(self._h, ctypes.c_void_p(t.data_ptr()), _np_ptr(offsets), len(L),
                                    ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(mp.data_ptr()),
                                    ctypes.c_void_p(lab.data_ptr()), ctypes.c_void_p(stream)))
        torch.cuda.synchronize(dev)
        return (ids.cpu().numpy()[:int((L * (L + 2)).sum())], mp.cpu().numpy()[:int(L.sum())],
                lab.cpu().numpy()[:int(L.sum())])

    def hidden(self, tokens: np.ndarray, offsets: np.ndarray, upto_layer: int):
        import torch
        offsets = self._check_packed(tokens, offsets)
        L = np.diff(offsets)
        rows = int((L * (L + 2)).sum())
        dev = torch.device("cuda", self.device)
        t = torch.from_numpy(np.ascontiguousarray(tokens, np.int32)).to(dev)
        out = torch.zeros(max(rows, 1), self.cfg["hidden"], dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(self._lib.pllb_debug_hidden(self._h, ctypes.c_void_p(t.data_ptr()), _np_ptr(offsets), len(L), upto_layer,
                                          ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)))
        torch.cuda.synchronize(dev)
        return out[:rows].cpu()

    # ------------------------------------------------------------------ stats
    def set_timing(self, enable: bool):
        check(self._lib.pllb_set_timing(self._h, int(enable)))

    def reset_stats(self):
        check(self._lib.pllb_reset_stats(self._h))

    def stats(self) -> dict:
        s = Stats()
        check(self._lib.pllb_get_stats(self._h, ctypes.byref(s)))
        ms = (ctypes.c_float * 6)()
        fl = (ctypes.c_double * 6)()
        check(self._lib.pllb_get_gemm_breakdown(self._h, ms, fl))
        d = {n: getattr(s, n) for n, _ in Stats._fields_}
        d["gemm_ms_by_kind"] = dict(zip(GEMM_KINDS, [float(x) for x in ms]))
        d["gemm_flops_by_kind"] = dict(zip(GEMM_KINDS, [float(x) for x in fl]))
        d["workspace_bytes"] = int(self._lib.pllb_workspace_bytes(self._h))
        return d


class MlmTrainer:
    """BERT masked-LM fine-tuning on the device; stands in for ``BertForMaskedLM`` (train mode) +
    ``torch.optim.AdamW`` inside ``run_one_epoch(train_mode=True)`` (MLM_PLL/main.py:73-99) and for
    the loss-only dev pass (train_mode=False, do_scoring=False).  fp32 master weights, bf16 GEMM
    operands, fp32 accumulation.  ``hidden_dropout`` / ``attention_dropout`` default to the
    BertConfig values of bert-base-chinese (0.1); the masks come from a stateless hash of
    (seed, step, site, element), not from torch's RNG, so only runs with dropout 0 are comparable
    value by value with the reference."""

    TIED = {"cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight",
            "cls.predictions.decoder.bias": "cls.predictions.bias"}

    def __init__(self, state_dict: Dict[str, "torch.Tensor"], cfg: Optional[dict] = None, device: int = 0,
                 lr: float = 1e-5, hidden_dropout: float = 0.1, attention_dropout: float = 0.1, seed: int = 0,
                 max_rows: int = 4096, max_seq: int = 128, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, pad_id: int = 0):
        import torch
        from ._lib import TrainDesc
        from .synth import config_from_state_dict

        self._lib = _lib.load()
        _lib.require_device()
        self._h = ctypes.c_void_p()
        self.cfg = dict(cfg) if cfg is not None else config_from_state_dict(state_dict)
        self.device = device
        self._dev = torch.device("cuda", device)
        if "cls.predictions.transform.dense.weight" not in state_dict:
            raise ValueError("MLM fine-tuning needs a BertForMaskedLM state_dict (cls.predictions.* missing)")
        self._keys = [k for k in state_dict.keys()]
        self._passthrough = {k: v for k, v in state_dict.items() if k.endswith("position_ids")}
        self._shapes = {k: tuple(v.shape) for k, v in state_dict.items()}
        nl = self.cfg["num_layers"]
        self.max_seq = min(int(max_seq), int(self.cfg["max_position"]), 512)
        w, keep = weights_struct(lambda key: state_dict[key].detach().to(device=self._dev, dtype=torch.float32).contiguous(),
                                 nl, state_dict.keys())
        d = ModelDesc(nl, self.cfg["hidden"], self.cfg["num_heads"], self.cfg["intermediate"], self.cfg["vocab"],
                      self.cfg["max_position"], float(self.cfg.get("ln_eps", 1e-12)), 101, 102, 103, 0)
        td = TrainDesc(float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(hidden_dropout),
                       float(attention_dropout), int(seed) & (2 ** 64 - 1), int(pad_id), int(max_rows), self.max_seq)
        torch.cuda.synchronize(self._dev)
        check(self._lib.pllb_train_create(ctypes.byref(self._h), ctypes.byref(d), ctypes.byref(w), ctypes.byref(td), device))
        del keep

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pllb_train_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset_optimizer(self, lr: float):
        """A fresh AdamW, as the reference creates one per epoch (MLM_PLL/main.py:76)."""
        check(self._lib.pllb_train_reset_optimizer(self._h, float(lr)))

    def step(self, input_ids, attention_mask, labels, mode: int = 1) -> float:
        """One zero-padded batch (int arrays [B, T], what collate builds).  mode 1: forward + backward
        + AdamW step; 0: loss only (eval); 2: forward + backward without the update.  Returns the loss."""
        ids = np.ascontiguousarray(input_ids, np.int32)
        lab = np.ascontiguousarray(labels, np.int32)
        am = np.asarray(attention_mask)
        if ids.ndim != 2 or lab.shape != ids.shape or am.shape != ids.shape:
            raise ValueError("input_ids, attention_mask and labels must be equal [B, T] arrays")
        nv = np.ascontiguousarray((am != 0).sum(1), np.int32)
        if ((am != 0) != (np.arange(ids.shape[1])[None, :] < nv[:, None])).any():
            raise ValueError("attention_mask must be a prefix mask (ones, then the zero padding of collate)")
        loss = ctypes.c_float(0.0)
        check(self._lib.pllb_train_step_host(self._h, _np_ptr(ids), _np_ptr(nv), _np_ptr(lab), ids.shape[0], ids.shape[1],
                                             int(mode), ctypes.byref(loss)))
        return float(loss.value)

    def row_losses(self, n_rows: int) -> np.ndarray:
        """-log_softmax(logits)[labels] of every position of the last step (float32 [B*T], row-major)."""
        out = np.zeros(int(n_rows), np.float32)
        check(self._lib.pllb_train_row_losses_host(self._h, _np_ptr(out), int(n_rows)))
        return out

    def _export(self, fn) -> Dict[str, "torch.Tensor"]:
        import torch
        bufs = {}

        def get(key):
            key = self.TIED.get(key, key) if self.TIED.get(key, key) in self._shapes else key
            if key not in bufs:
                bufs[key] = torch.zeros(self._shapes[key], dtype=torch.float32, device=self._dev)
            return bufs[key]

        w, keep = weights_struct(get, self.cfg["num_layers"], self._keys)
        torch.cuda.synchronize(self._dev)
        check(fn(self._h, ctypes.byref(w)))
        torch.cuda.synchronize(self._dev)
        out = {}
        for k in self._keys:
            if k in self._passthrough:
                out[k] = self._passthrough[k]
            else:
                src = self.TIED.get(k, k)
                out[k] = bufs[src if src in bufs else k].cpu()
        for k, src in self.TIED.items():       # tied entries share storage, as in model.state_dict()
            if k in out and src in out:
                out[k] = out[src]
        del keep
        return out

    def state_dict(self) -> Dict[str, "torch.Tensor"]:
        """fp32 CPU tensors under the keys of the state_dict given at construction (model.state_dict(),
        MLM_PLL/main.py:155)."""
        return self._export(self._lib.pllb_train_export)

    def grads(self) -> Dict[str, "torch.Tensor"]:
        """Gradients of the last mode-1 / mode-2 step, same keys (parity tests)."""
        return self._export(self._lib.pllb_train_export_grads)

    def kernel_launches(self) -> int:
        return int(self._lib.pllb_train_kernel_launches(self._h))

    def graph_replays(self) -> int:
        """Steps that ran as replays of a captured CUDA graph (PLLB_TRAIN_GRAPH=0 disables capture)."""
        return int(self._lib.pllb_train_graph_replays(self._h))


# ---------------------------------------------------------------------- stage 4
def pack_strings(strings: Sequence[str]):
    """Strings -> (code points int32[sum len], offsets int64[n+1]); len = Python code points."""
    off = np.zeros(len(strings) + 1, np.int64)
    if len(strings):
        np.cumsum(np.fromiter(map(len, strings), np.int64, len(strings)), out=off[1:])
    cp = np.frombuffer("".join(strings).encode("utf-32-le", "surrogatepass"), np.int32).copy()
    assert len(cp) == off[-1]
    return cp, off


_WS = None


def _whitespace_code_points() -> np.ndarray:
    """Code points str.strip() removes (str.isspace()), for the vectorised strip check."""
    global _WS
    if _WS is None:
        _WS = np.array([c for c in range(0x3001) if chr(c).isspace()], np.int32)   # none above U+3000
    return _WS


def pack_stripped(strings: Sequence[str]):
    """pack_strings([s.strip() for s in strings]) plus the RAW lengths, in one pass over the list:
    the strings are packed unstripped, and only if some string starts or ends with a whitespace
    code point (checked on the packed array) is the per-string strip() done.  Returns
    (code points, offsets, raw lengths int64)."""
    n = len(strings)
    raw_len = np.fromiter(map(len, strings), np.int64, n) if n else np.zeros(0, np.int64)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(raw_len, out=off[1:])
    cp = np.frombuffer("".join(strings).encode("utf-32-le", "surrogatepass"), np.int32)
    assert len(cp) == off[-1]
    nz = raw_len > 0
    ws = _whitespace_code_points()
    if nz.any() and (np.isin(cp[off[:-1][nz]], ws).any() or np.isin(cp[off[1:][nz] - 1], ws).any()):
        cp, off = pack_strings([s.strip() for s in strings])
        return cp, off, raw_len
    return cp.copy(), off, raw_len


def tokenize_packed(table: np.ndarray, cp: np.ndarray, cp_off: np.ndarray):
    """pllb_tokenize_host: packed code points -> (ids int32, offsets int64[n+1], needs_host uint8[n])."""
    lib = _lib.load()
    _lib.require_device()
    table = np.ascontiguousarray(table, np.int32)
    cp = np.ascontiguousarray(cp, np.int32)
    cp_off = np.ascontiguousarray(cp_off, np.int64)
    n = len(cp_off) - 1
    ids = np.zeros(max(int(cp_off[-1] - cp_off[0]), 1), np.int32)
    off = np.zeros(n + 1, np.int64)
    flag = np.zeros(max(n, 1), np.uint8)
    check(lib.pllb_tokenize_host(_np_ptr(table), len(table), _np_ptr(cp), _np_ptr(cp_off), n, _np_ptr(ids), _np_ptr(off),
                                 _np_ptr(flag)))
    return ids[:int(off[-1])], off, flag[:n]


def levenshtein_packed(ref_cp, ref_off, hyp_cp, hyp_off, pair_ref) -> np.ndarray:
    """Edit distances of (ref[pair_ref[i]], hyp[i]) pairs on the GPU (host arrays in/out)."""
    lib = _lib.load()
    _lib.require_device()
    ref_cp = np.ascontiguousarray(ref_cp, np.int32)
    hyp_cp = np.ascontiguousarray(hyp_cp, np.int32)
    ref_off = np.ascontiguousarray(ref_off, np.int64)
    hyp_off = np.ascontiguousarray(hyp_off, np.int64)
    pair_ref = np.ascontiguousarray(pair_ref, np.int32)
    out = np.zeros(len(pair_ref), np.int32)
    check(lib.pllb_levenshtein_host(_np_ptr(ref_cp), _np_ptr(ref_off), len(ref_off) - 1, _np_ptr(hyp_cp),
                                    _np_ptr(hyp_off), _np_ptr(pair_ref), len(pair_ref), _np_ptr(out)))
    return out


def levenshtein(refs: Sequence[str], hyps: Sequence[str], pair_ref: Optional[Sequence[int]] = None) -> np.ndarray:
    """jiwer-style character edit distance: strings are stripped, characters = code points."""
    refs = [r.strip() for r in refs]
    hyps = [h.strip() for h in hyps]
    rc, ro = pack_strings(refs)
    hc, ho = pack_strings(hyps)
    if pair_ref is None:
        if len(refs) != len(hyps):
            raise ValueError("reference and hypothesis lists differ in length")
        pair_ref = np.arange(len(hyps), dtype=np.int32)
    return levenshtein_packed(rc, ro, hc, ho, pair_ref)


def rescore_sweep(am, lm, lens, dist, weights, variant="B"):
    """For every weight: per-utterance argmax of the interpolated score and the summed edit
    distance of the chosen hypotheses.  Returns (argmax int32 [W,N], edit_sum int64 [W])."""
    lib = _lib.load()
    _lib.require_device()
    am = np.ascontiguousarray(am, np.float64)
    lm = np.ascontiguousarray(lm, np.float64)
    lens = np.ascontiguousarray(lens, np.int64)
    weights = np.ascontiguousarray(weights, np.float64)
    if am.shape != lm.shape or am.shape != lens.shape or am.ndim != 2:
        raise ValueError(f"am {am.shape}, lm {lm.shape}, len {lens.shape} must be equal 2-D shapes")
    N, nb = am.shape
    W = len(weights)
    arg = np.zeros((W, N), np.int32)
    es = np.zeros(W, np.int64)
    d = None
    if dist is not None:
        d = np.ascontiguousarray(dist, np.int32)
        if d.shape != am.shape:
            raise ValueError("dist shape mismatch")
    check(lib.pllb_rescore_sweep_host(_np_ptr(am), _np_ptr(lm), _np_ptr(lens), _np_ptr(d) if d is not None else None,
                                      N, nb, _np_ptr(weights), W, VARIANTS[variant], _np_ptr(arg), _np_ptr(es)))
    return arg, es


def rescore_scores(am, lm, lens, weight, variant="B") -> np.ndarray:
    """The [N, n_best] interpolated score matrix for one weight (rescore.py:47-53)."""
    import torch
    lib = _lib.load()
    _lib.require_device()
    am = np.ascontiguousarray(am, np.float64)
    lm = np.ascontiguousarray(lm, np.float64)
    lens = np.ascontiguousarray(lens, np.int64)
    if am.shape != lm.shape or am.shape != lens.shape or am.ndim != 2:
        raise ValueError(f"am {am.shape}, lm {lm.shape}, len {lens.shape} must be equal 2-D shapes")
    N, nb = am.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    d_am, d_lm, d_len = (torch.from_numpy(x).to(dev) for x in (am, lm, lens))
    out = torch.empty(N, nb, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(lib.pllb_rescore_scores(ctypes.c_void_p(d_am.data_ptr()), ctypes.c_void_p(d_lm.data_ptr()),
                                  ctypes.c_void_p(d_len.data_ptr()), N, nb, float(weight), VARIANTS[variant],
                                  ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)))
    torch.cuda.synchronize(dev)
    return out.cpu().numpy()


def debug_gemm(A_bf16, W_bf16, bias_f32, epilogue: int, simt: bool = False, operand_dtype: str = "bf16"):
    """Test hook: C = A @ W^T + bias through the tcgen05 (or SIMT validation) kernel.
    operand_dtype "fp16": A and W are torch.float16."""
    import torch
    lib = _lib.load()
    _lib.require_device()
    M, K = A_bf16.shape
    N = W_bf16.shape[0]
    out = torch.empty(M, N, dtype=A_bf16.dtype if epilogue in (0, 1) else torch.float32, device=A_bf16.device)
    stream = torch.cuda.current_stream(A_bf16.device).cuda_stream
    args = (ctypes.c_void_p(A_bf16.data_ptr()), ctypes.c_void_p(W_bf16.data_ptr()), ctypes.c_void_p(bias_f32.data_ptr()),
            ctypes.c_void_p(out.data_ptr()), M, N, K, epilogue)
    if simt:
        check(lib.pllb_debug_gemm_simt(*args, ctypes.c_void_p(stream)))
    elif operand_dtype == "bf16":
        check(lib.pllb_debug_gemm(*args, ctypes.c_void_p(stream)))
    else:
        check(lib.pllb_debug_gemm_dt(*args, OPERAND_DTYPES[operand_dtype], ctypes.c_void_p(stream)))
    return out
