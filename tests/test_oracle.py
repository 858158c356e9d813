"""The oracle against the reference's own fixtures and live outputs (CPU only).

Golden files under tests/golden/ were produced by oracle/make_golden.py from the
UNMODIFIED reference (Nbest_Align/cer.json, align.py docstring examples, rescore_result logs,
rescore.py functions, MLM_PLL/main.py run_one_epoch on transformers.BertForMaskedLM)."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import pll_oracle, rescore_oracle
from asr_rescoring_b200 import synth


@pytest.fixture(scope="module")
def kat(gold_dir):
    return np.load(os.path.join(gold_dir, "levenshtein_kat.npz")), json.load(open(os.path.join(gold_dir, "levenshtein_meta.json"), encoding="utf-8"))


def test_levenshtein_c_oracle_matches_all_reference_kats(kat):
    z, meta = kat
    got = oracle.levenshtein_batch(z["ref_cp"], z["ref_off"], z["hyp_cp"], z["hyp_off"], np.arange(len(z["dist"]), dtype=np.int32))
    assert len(got) == 7176 == meta["n"]
    assert np.array_equal(got, z["dist"])
    assert int(got.sum()) == 3230 and int(z["ref_off"][-1]) == 104765


def test_levenshtein_python_oracle_matches_kats(kat):
    z, _ = kat
    for i in list(range(0, 7176, 37)) + list(np.nonzero(z["dist"])[0][:100]):
        r = "".join(map(chr, z["ref_cp"][z["ref_off"][i]:z["ref_off"][i + 1]]))
        h = "".join(map(chr, z["hyp_cp"][z["hyp_off"][i]:z["hyp_off"][i + 1]]))
        assert rescore_oracle.levenshtein(r, h) == z["dist"][i]


def test_align_docstring_examples(kat):
    _, meta = kat
    for ex in meta["docstring_examples"]:      # espnet_data/preprocess/align.py:13-18
        d = oracle.levenshtein_strings(["".join(chr(0x100 + hash(w) % 5000) for w in [])], [""])  # smoke: empty vs empty
        assert d[0] == 0
        words = {w: chr(0x4E00 + i) for i, w in enumerate(dict.fromkeys(ex["ref"] + ex["hyp"]))}
        r = "".join(words[w] for w in ex["ref"])
        h = "".join(words[w] for w in ex["hyp"])
        assert oracle.levenshtein_strings([r], [h])[0] == ex["distance"]
    assert [e["distance"] for e in meta["docstring_examples"]] == [1, 2]


def test_logged_corpus_cers_are_integer_counts_over_ref_length(kat):
    _, meta = kat
    assert meta["ref_text_total_chars"] == 104765 and len(meta["logged_test_cer"]) == 17
    for e in meta["logged_test_cer"]:           # corpus CER = sum edits / sum len(ref)
        x = e["test_cer"] * 104765
        assert abs(x - round(x)) < 1e-6, e
    cers = {e["log"]: e["test_cer"] for e in meta["logged_test_cer"]}
    assert round(cers["rescore_result/MLM_PLL/rescore.log"] * 104765) == 5038


def test_cer_restatement_semantics():
    assert rescore_oracle.cer(["abc", "de"], ["abd", "de"]) == 1 / 5
    assert rescore_oracle.cer(" abc ", "abc") == 0.0
    with pytest.raises(ValueError):
        rescore_oracle.cer([""], ["a"])


@pytest.fixture(scope="module")
def comb(gold_dir):
    return np.load(os.path.join(gold_dir, "combiner_golden.npz"))


def _bits_equal(a, b):
    return bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())


def test_combiner_oracles_bit_exact_vs_reference(comb):
    z = comb
    cfg = rescore_oracle.config(10)
    for i, wi in enumerate(z["score_idx"]):
        w = z["weights"][wi]
        with np.errstate(all="ignore"):
            s_np = rescore_oracle.rescore(w, z["lens"], z["am"], z["lm"], cfg)
        s_c = oracle.rescore_scores(z["am"], z["lm"], z["lens"], w, 0)
        assert _bits_equal(s_np, z["scores"][i])
        assert _bits_equal(s_c, z["scores"][i])
    dist = np.zeros(z["am"].shape, np.int32)
    arg, _ = oracle.rescore_sweep(z["am"], z["lm"], z["lens"], dist, z["weights"], 0)
    assert np.array_equal(arg, z["argmax"])      # incl. the exact-tie row and the NaN row


def test_find_best_weight_restatement_vs_reference(comb):
    z = comb
    N, nb = z["am"].shape
    hyps = [["".join(map(chr, z["hyp_cp"][z["hyp_off"][u * nb + k]:z["hyp_off"][u * nb + k + 1]])) for k in range(nb)] for u in range(N)]
    refs = ["".join(map(chr, z["ref_cp"][z["ref_off"][u]:z["ref_off"][u + 1]])) for u in range(N)]
    bw, bc = rescore_oracle.find_best_weight(z["am"].tolist(), z["lm"].tolist(), hyps, refs, rescore_oracle.config(10))
    assert bw == float(z["best_weight"]) and bc == float(z["best_cer"])
    # the same through the C sweep + integer edit sums
    pair_ref = np.repeat(np.arange(N, dtype=np.int32), nb)
    dist = oracle.levenshtein_batch(z["ref_cp"], z["ref_off"], z["hyp_cp"], z["hyp_off"], pair_ref).reshape(N, nb)
    _, es = oracle.rescore_sweep(z["am"], z["lm"], z["lens"], dist, z["weights"], 0)
    cers = es / float(z["ref_off"][-1])
    assert z["weights"][int(np.argmin(cers))] == float(z["best_weight"]) and cers.min() == float(z["best_cer"])


def test_expand_oracle_matches_row_schema():
    nb = synth.make_nbest(5, 3, seed=4)
    nb.hyps[1][1] = ""
    tok, off = nb.packed_tokens()
    ids, mp, lab = oracle.expand(tok, off)
    rows = []
    for h in range(len(off) - 1):
        rows += pll_oracle.expand_rows([int(t) for t in tok[off[h]:off[h + 1]]], "u", "h")
    assert ids.tolist() == [t for r in rows for t in r["input_ids"]]
    assert mp.tolist() == [r["mask_pos"] for r in rows]
    assert lab.tolist() == [r["labels"][r["mask_pos"]] for r in rows]


@pytest.mark.parametrize("case_name", ["tiny_perturbed", "base_chinese_perturbed"])
def test_pll_oracle_reproduces_reference_run(gold_dir, case_name):
    gold = json.load(open(os.path.join(gold_dir, "pll_golden.json")))
    case = next(c for c in gold["cases"] if c["name"] == case_name)
    sd = synth.random_init_state_dict(case["cfg"], case["seed"], case["perturb"])
    got = pll_oracle.score_hyps(sd, case["cfg"], case["hyps"])
    for u, hs in case["pll"].items():
        for h, v in hs.items():
            assert abs(got[u][h] - v) < 2e-4, (u, h, got[u][h], v)
            if len(case["hyps"][u][h]) == 0:
                assert got[u][h] == 0 and isinstance(got[u][h], int)   # skeleton int 0 survives


def test_algorithmic_flops_formula():
    # bert-base-chinese, L=14 -> 38.65 GFLOP per hypothesis
    f = pll_oracle.algorithmic_flops([14], synth.BERT_BASE_CHINESE)
    assert abs(f / 1e9 - 38.65) < 0.05


# ---------------------------------------------------------------------- MLM fine-tuning oracle
def _summary(t):
    import torch
    v = t.detach().double().reshape(-1)
    g = torch.Generator().manual_seed(1234 + v.numel() % 977)
    sign = torch.randint(0, 2, (v.numel(),), generator=g, dtype=torch.int64).double() * 2 - 1
    return float(v.norm()), float((v * sign).sum())


def test_training_oracle_matches_the_reference_training_golden(gold_dir):
    """oracle/train_oracle.py (autograd restatement of run_one_epoch(train_mode=True)) against
    tests/golden/train_golden.json, which oracle/make_golden_train.py recorded from the UNMODIFIED
    MLM_PLL/main.py loop on transformers.BertForMaskedLM (dropout 0): first-batch loss and per-tensor
    gradient summaries, the epoch losses of train and dev passes, the final weights."""
    from oracle import train_oracle
    gold = json.load(open(os.path.join(gold_dir, "train_golden.json")))
    for case in gold["cases"]:
        cfg, bs, lr = case["cfg"], case["batch_size"], case["lr"]
        sd = synth.random_init_state_dict(cfg, case["seed"], case["perturb"])
        train_rows = train_oracle.training_rows(case["train_tokens"])
        dev_rows = train_oracle.training_rows(case["dev_tokens"])
        loss, grads = train_oracle.loss_and_grads(sd, cfg, train_rows[:bs])
        assert abs(loss - case["first_batch_loss"]) < 1e-5
        for k, ref in case["first_batch_grads"].items():
            name = train_oracle.TIED.get(k, k)
            norm, proj = _summary(grads[name])
            assert abs(norm - ref["norm"]) <= 2e-4 * ref["norm"] + 1e-7, (k, norm, ref["norm"])
            assert abs(proj - ref["proj"]) <= 2e-4 * ref["norm"] + 1e-7, (k, proj, ref["proj"])
        if case["name"] != "tiny_lr1e-5":          # one full case is enough for the CPU budget
            continue
        params = train_oracle.parameters(sd)
        for e in range(case["epochs"]):
            tr = train_oracle.run_one_epoch(params, cfg, train_rows, bs, lr, True)
            dv = train_oracle.run_one_epoch(params, cfg, dev_rows, bs, lr, False)
            assert abs(tr - case["train_loss"][e]) < 2e-4 and abs(dv - case["dev_loss"][e]) < 2e-4
        final = train_oracle.state_dict_of(params)
        for k, ref in case["final_minus_init"].items():
            if k.endswith("attention.self.key.bias"):
                continue        # its true gradient is 0 (softmax ignores a per-row constant): Adam amplifies pure rounding noise
            norm, proj = _summary(final[k] - sd[k])
            assert abs(norm - ref["norm"]) <= 0.02 * ref["norm"] + 1e-9, (k, norm, ref["norm"])
