"""Generate tests/golden/* by running the UNMODIFIED reference in the build container.

    python oracle/make_golden.py            (needs /root/reference; not run on the GPU box)

What it does
  1. Levenshtein KATs: converts the reference's Nbest_Align/cer.json (7 176
     {ref,pred,cer} triples) to packed code points + the implied integer
     distance round(cer*len(ref)) -> tests/golden/levenshtein_kat.npz, plus the
     two docstring examples of espnet_data/preprocess/align.py:13-18 (executed
     through the reference's own function) and the 17 logged corpus CERs.
  2. PLL: imports /root/reference/MLM_PLL/main.py by path (ruamel.yaml shimmed by
     PyYAML, the only change), builds transformers.BertForMaskedLM with the
     state_dict from asr_rescoring_b200.synth.random_init_state_dict, and runs the
     reference's set_dataloader + run_one_epoch(train_mode=False, do_scoring=True)
     on CPU.  Outputs -> tests/golden/pll_golden.json.  Also asserts that the
     oracle restatement (oracle/pll_oracle.py) agrees to < 2e-4 nats.
  3. Combiner: imports /root/reference/rescore.py (jiwer shimmed by the oracle's
     cer, validated in step 1) and records rescore / argmax / find_best_weight
     outputs on synthetic inputs -> tests/golden/combiner_golden.npz.
"""
from __future__ import annotations

import glob
import importlib.util
import json
import os
import re
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import pll_oracle, rescore_oracle  # noqa: E402
from asr_rescoring_b200 import synth  # noqa: E402


def _install_shims():
    import yaml
    ruamel = types.ModuleType("ruamel")
    ruamel_yaml = types.ModuleType("ruamel.yaml")
    ruamel_yaml.load = yaml.load
    ruamel_yaml.Loader = yaml.Loader
    ruamel.yaml = ruamel_yaml
    sys.modules["ruamel"] = ruamel
    sys.modules["ruamel.yaml"] = ruamel_yaml
    jiwer = types.ModuleType("jiwer")
    jiwer.cer = rescore_oracle.cer
    sys.modules["jiwer"] = jiwer


def _import_by_path(name, path, extra_sys_path):
    sys.path.insert(0, extra_sys_path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def levenshtein_golden():
    triples = json.load(open(os.path.join(REF, "Nbest_Align", "cer.json"), encoding="utf-8"))
    refs = [t["ref"] for t in triples]
    preds = [t["pred"] for t in triples]
    dist = np.array([round(t["cer"] * len(t["ref"])) for t in triples], np.int32)
    for t, d in zip(triples, dist):
        assert abs(d / len(t["ref"]) - t["cer"]) < 1e-12
    rc, ro = oracle.pack_strings(refs)
    pc, po = oracle.pack_strings(preds)
    # restatement pinned: all 7 176 exact
    got = oracle.levenshtein_batch(rc, ro, pc, po, np.arange(len(refs), dtype=np.int32))
    assert np.array_equal(got, dist), "C oracle disagrees with Nbest_Align/cer.json"
    py = np.array([rescore_oracle.levenshtein(r, p) for r, p in zip(refs[:500], preds[:500])])
    assert np.array_equal(py, dist[:500])
    # docstring examples through the reference's own alignment function
    align = _import_by_path("ref_align", os.path.join(REF, "espnet_data", "preprocess", "align.py"),
                            os.path.join(REF, "espnet_data", "preprocess"))
    ex = []
    for r, h in ((["how", "are", "you"], ["how", "are", "you", "doing"]), (["你", "好", "嗎"], ["你", "好", "不", "好"])):
        out = align.levenshtein_distance_alignment(list(r), list(h))
        ex.append(dict(ref=r, hyp=h, ops=out[2], distance=sum(o != "U" for o in out[2])))
    logged = []
    for p in sorted(glob.glob(os.path.join(REF, "rescore_result", "**", "*.log"), recursive=True)):
        for line in open(p, encoding="utf-8"):
            m = re.search(r"test cer: ([0-9.eE+-]+)", line)
            if m:
                logged.append(dict(log=os.path.relpath(p, REF), test_cer=float(m.group(1))))
    ref_text = json.load(open(os.path.join(REF, "espnet_data/alfred/test/ref_text.json"), encoding="utf-8"))
    total = sum(len(v) for v in ref_text.values())
    np.savez_compressed(os.path.join(GOLD, "levenshtein_kat.npz"), ref_cp=rc, ref_off=ro, hyp_cp=pc, hyp_off=po, dist=dist)
    json.dump(dict(source="Nbest_Align/cer.json", n=len(refs), sum_dist=int(dist.sum()), sum_ref_len=int(ro[-1]),
                   docstring_examples=ex, logged_test_cer=logged, ref_text_total_chars=total),
              open(os.path.join(GOLD, "levenshtein_meta.json"), "w", encoding="utf-8"), ensure_ascii=False, indent=1)
    print(f"levenshtein: {len(refs)} KATs, sum dist {dist.sum()}, sum len {ro[-1]}, {len(logged)} logged CERs")


def _hf_model(cfg, sd):
    from transformers import BertConfig, BertForMaskedLM
    hf = BertForMaskedLM(BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"],
                                    num_hidden_layers=cfg["num_layers"], num_attention_heads=cfg["num_heads"],
                                    intermediate_size=cfg["intermediate"], max_position_embeddings=cfg["max_position"],
                                    type_vocab_size=cfg["type_vocab"], layer_norm_eps=cfg["ln_eps"],
                                    pad_token_id=0, hidden_act="gelu"))
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    return hf.eval()


def pll_golden():
    ref_main = _import_by_path("ref_mlm_pll_main", os.path.join(REF, "MLM_PLL", "main.py"), os.path.join(REF, "MLM_PLL"))
    cases = []
    specs = [
        ("tiny_perturbed", synth.BERT_TINY, 10, True, dict(n_utts=6, n_best=4, seed=3)),
        ("base_chinese_seed10", synth.BERT_BASE_CHINESE, 10, False, dict(n_utts=3, n_best=4, seed=5)),
        ("base_chinese_perturbed", synth.BERT_BASE_CHINESE, 11, True, dict(n_utts=2, n_best=3, seed=7)),
    ]
    for name, cfg, seed, perturb, nbk in specs:
        sd = synth.random_init_state_dict(cfg, seed, perturb)
        nb = synth.make_nbest(**nbk)
        if name == "tiny_perturbed":   # edge cases: 1-token and empty hypotheses
            nb.hyps[0][1] = nb.hyps[0][1][:1]
            nb.hyps[1][2] = ""
        tok, off = nb.packed_tokens(cfg["vocab"])
        hyps = {}
        i = 0
        for u, hs in zip(nb.utt_ids, nb.hyps):
            hyps[u] = {}
            for k in range(len(hs)):
                hyps[u][f"hyp_{k + 1}"] = [int(t) for t in tok[off[i]:off[i + 1]]]
                i += 1
        rows, skel = [], {}
        for u, hs in hyps.items():
            skel[u] = {}
            for h, toks in hs.items():
                skel[u][h] = 0
                rows += pll_oracle.expand_rows(toks, u, h)
        hf = _hf_model(cfg, sd)
        loader = ref_main.set_dataloader(SimpleNamespace(batch_size=32, num_worker=0), ref_main.MyDataset(rows), True)
        with torch.no_grad():
            ref_out = ref_main.run_one_epoch(config=SimpleNamespace(device="cpu"), model=hf, dataloader=loader,
                                             output_score={u: dict(v) for u, v in skel.items()},
                                             train_mode=False, do_scoring=True)
        mine = pll_oracle.score_hyps(sd, cfg, hyps)
        worst = max(abs(ref_out[u][h] - mine[u][h]) for u in hyps for h in hyps[u])
        print(f"pll[{name}]: {len(rows)} copies, oracle-vs-reference max |dPLL| = {worst:.2e}")
        assert worst < 2e-4
        cases.append(dict(name=name, cfg=cfg, seed=seed, perturb=perturb, hyps=hyps, pll=ref_out,
                          oracle_vs_reference_max_abs=worst))
    json.dump(dict(generator="oracle/make_golden.py", reference="MLM_PLL/main.py run_one_epoch (unmodified), "
                   f"transformers {__import__('transformers').__version__}, torch {torch.__version__}", cases=cases),
              open(os.path.join(GOLD, "pll_golden.json"), "w"), indent=1)


def rescore_bert_golden():
    """RescoreBert scoring through the reference's own forward (RescoreBert/model.py:13-21) on a
    transformers.BertModel carrying our encoder weights; BertModel.from_pretrained needs the HF
    hub, so the module object is assembled by hand and only its forward() comes from the reference."""
    from transformers import BertConfig, BertModel
    ref_model = _import_by_path("ref_rescorebert_model", os.path.join(REF, "RescoreBert", "model.py"),
                                os.path.join(REF, "RescoreBert"))
    cases = []
    for name, cfg, seed in (("tiny", synth.BERT_TINY, 31), ("base_chinese", synth.BERT_BASE_CHINESE, 10)):
        sd = synth.random_init_state_dict(cfg, seed, perturb=(name == "tiny"))
        g = torch.Generator().manual_seed(seed + 1)
        lin_w = torch.empty(1, cfg["hidden"]).normal_(0, 0.05, generator=g)
        lin_b = float(torch.empty(1).normal_(0, 0.5, generator=g))
        bert = BertModel(BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["num_layers"],
                                    num_attention_heads=cfg["num_heads"], intermediate_size=cfg["intermediate"],
                                    max_position_embeddings=cfg["max_position"], type_vocab_size=cfg["type_vocab"],
                                    layer_norm_eps=cfg["ln_eps"], pad_token_id=0, hidden_act="gelu"))
        enc = {k[len("bert."):]: v for k, v in sd.items() if k.startswith("bert.")}
        missing, unexpected = bert.load_state_dict(enc, strict=False)
        assert not unexpected and all("pooler" in m or "position_ids" in m for m in missing), (missing, unexpected)
        m = object.__new__(ref_model.RescoreBert)
        torch.nn.Module.__init__(m)
        m.bert = bert.eval()
        m.linear = torch.nn.Linear(cfg["hidden"], 1)
        with torch.no_grad():
            m.linear.weight.copy_(lin_w)
            m.linear.bias.fill_(lin_b)
        nb = synth.make_nbest(4 if name == "tiny" else 2, 4, seed=seed)
        if name == "tiny":
            nb.hyps[0][1] = ""
        tok, off = nb.packed_tokens(cfg["vocab"])
        lists = [[int(t) for t in tok[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
        rows = [[101] + t + [102] for t in lists]
        T = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), T, dtype=torch.long)
        am = torch.zeros(len(rows), T, dtype=torch.long)
        for i, r in enumerate(rows):                      # pad_sequence(batch_first=True), RescoreBert/main.py:59-60
            ids[i, :len(r)] = torch.tensor(r)
            am[i, :len(r)] = 1
        with torch.no_grad():
            ref_scores = m(ids, am).tolist()
        mine = pll_oracle.rescore_bert_scores(sd, cfg, lists, lin_w, lin_b)
        worst = max(abs(a - b) for a, b in zip(ref_scores, mine))
        print(f"rescorebert[{name}]: {len(rows)} hyps, oracle-vs-reference max |d| = {worst:.2e}")
        assert worst < 1e-4
        cases.append(dict(name=name, cfg=cfg, seed=seed, perturb=(name == "tiny"), tokens=lists,
                          linear_w=lin_w.reshape(-1).tolist(), linear_b=lin_b, scores=ref_scores))
    json.dump(dict(generator="oracle/make_golden.py", reference="RescoreBert/model.py forward (unmodified)", cases=cases),
              open(os.path.join(GOLD, "rescorebert_golden.json"), "w"), indent=1)


SCORE_IDX = [0, 1, 29, 33, 50, 100]   # grid points whose full score matrix is stored


def combiner_golden():
    ref_rescore = _import_by_path("ref_rescore", os.path.join(REF, "rescore.py"), REF)
    nb = synth.make_nbest(300, 10, seed=11)
    lm = synth.synthetic_lm_scores(nb, seed=2)
    # exact ties and a NaN-producing empty hypothesis (rescore.py:51 divides by len 0)
    lm[5, 3] = lm[5, 0]; nb.am[5, 3] = nb.am[5, 0]; nb.hyps[5][3] = nb.hyps[5][0]
    nb.hyps[7][2] = ""
    cfg = SimpleNamespace(n_best=10)
    am_l, lm_l = nb.am.tolist(), lm.tolist()
    lens = [[len(h) for h in hs] for hs in nb.hyps]
    weights = np.arange(0.0, 1.01, 0.01)
    scores = []
    argmax = []
    with np.errstate(all="ignore"):
        for w in weights:
            s = ref_rescore.rescore(w, lens, am_l, lm_l, cfg)
            scores.append(s)
            argmax.append(np.argmax(s, axis=-1))
        bw, bc = ref_rescore.find_best_weight(am_l, lm_l, nb.hyps, nb.refs, cfg)
    # 6-best slice behaviour (am is sliced, lm must already be n_best wide: rescore.py:48-49)
    cfg6 = SimpleNamespace(n_best=6)
    lm6 = [r[:6] for r in lm_l]
    with np.errstate(all="ignore"):
        bw6, bc6 = ref_rescore.find_best_weight(am_l, lm6, nb.hyps, nb.refs, cfg6)
    hc, ho = oracle.pack_strings([h for hs in nb.hyps for h in hs])
    rc, ro = oracle.pack_strings(nb.refs)
    np.savez_compressed(os.path.join(GOLD, "combiner_golden.npz"), am=nb.am, lm=lm, lens=np.array(lens, np.int64),
                        hyp_cp=hc, hyp_off=ho, ref_cp=rc, ref_off=ro, weights=weights,
                        score_idx=np.array(SCORE_IDX), scores=np.stack([scores[i] for i in SCORE_IDX]),
                        argmax=np.stack(argmax).astype(np.int32),
                        best_weight=np.float64(bw), best_cer=np.float64(bc),
                        best_weight6=np.float64(bw6), best_cer6=np.float64(bc6))
    print(f"combiner: best_weight {bw} cer {bc}; 6-best {bw6} {bc6}")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    oracle.build()
    _install_shims()
    levenshtein_golden()
    combiner_golden()
    pll_golden()
    rescore_bert_golden()
