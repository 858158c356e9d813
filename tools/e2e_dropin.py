"""End-to-end run of the drop-in CLIs on synthetic files: MLM_PLL/main.py (1 GPU, and N GPUs under
torchrun when --gpus N) then rescore.py, checked against the oracle.

    python tools/e2e_dropin.py [--gpus 2] [--model bert-tiny-test]
"""
import argparse, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from asr_rescoring_b200 import synth
from oracle import pll_oracle, rescore_oracle

PKG = os.path.join(ROOT, "asr-rescoring_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--model", default="bert-tiny-test")
    ap.add_argument("--utts", type=int, default=24)
    args = ap.parse_args()
    cfg = {"bert-tiny-test": synth.BERT_TINY, "bert-base-chinese": synth.BERT_BASE_CHINESE}[args.model]
    tmp = tempfile.mkdtemp(prefix="pllb_e2e_")
    os.makedirs(os.path.join(tmp, "result"))
    splits = {}
    for i, split in enumerate(("train", "dev", "test")):
        nb = synth.make_nbest(args.utts, 5, seed=100 + i)
        nb.hyps[1][2] = ""                                # empty hypothesis keeps the int 0
        splits[split] = nb
        json.dump(nb.hyps_text(), open(os.path.join(tmp, f"{split}_hyps_text.json"), "w"), ensure_ascii=False)
        json.dump(nb.hyps_score(), open(os.path.join(tmp, f"{split}_hyps_score.json"), "w"))
        json.dump(nb.ref_text(), open(os.path.join(tmp, f"{split}_ref_text.json"), "w"), ensure_ascii=False)
    yaml_txt = f"""task: scoring
seed: 10
device: "cuda:0"
train_data_path: "{tmp}/train_hyps_text.json"
dev_data_path: "{tmp}/dev_hyps_text.json"
test_data_path: "{tmp}/test_hyps_text.json"
checkpoint_path: "{tmp}/no_checkpoint.pth"
output_path: "{tmp}/result/"
num_of_data: 99999999
dataloader:
  batch_size: 32
  num_worker: 5
model:
  bert: "{args.model}"
  random_init_seed: 10
"""
    open(os.path.join(tmp, "score.yaml"), "w").write(yaml_txt)
    main_py = os.path.join(PKG, "MLM_PLL", "main.py")
    subprocess.check_call([sys.executable, main_py, "--config", os.path.join(tmp, "score.yaml")])
    one = {s: json.load(open(os.path.join(tmp, "result", f"{s}_lm.json"))) for s in splits}
    if args.gpus > 1:
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                               "--master-addr", "127.0.0.1", "--master-port", "29577", main_py, "--config",
                               os.path.join(tmp, "score.yaml")])
        multi = {s: json.load(open(os.path.join(tmp, "result", f"{s}_lm.json"))) for s in splits}
        assert multi == one, "multi-GPU scores differ from single-GPU scores"
        print(f"{args.gpus}-GPU *_lm.json identical to the 1-GPU files (bitwise, {sum(len(v) for v in one.values())} utterances)")
        # the reference's row-list input (MLM_PLL/preprocess.py schema), cut mid-hypothesis by num_of_data
        from asr_rescoring_b200.tokenizer import SyntheticCharTokenizer as _Tk
        tk_ = _Tk()
        for split, nb_ in splits.items():
            rows = []
            for u, hs in nb_.hyps_text().items():
                for h, sent in hs.items():
                    rows += pll_oracle.expand_rows(tk_.encode(sent), u, h)
            json.dump(rows, open(os.path.join(tmp, f"{split}_rows.json"), "w"))
        yaml_rows = yaml_txt.replace("_hyps_text.json", "_rows.json").replace("num_of_data: 99999999", "num_of_data: 777") \
                            .replace(f"{tmp}/result/", f"{tmp}/result_rows/")
        os.makedirs(os.path.join(tmp, "result_rows"))
        open(os.path.join(tmp, "score_rows.yaml"), "w").write(yaml_rows)
        subprocess.check_call([sys.executable, main_py, "--config", os.path.join(tmp, "score_rows.yaml")])
        one_rows = {s: json.load(open(os.path.join(tmp, "result_rows", f"{s}_lm.json"))) for s in splits}
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                               "--master-addr", "127.0.0.1", "--master-port", "29578", main_py, "--config",
                               os.path.join(tmp, "score_rows.yaml")])
        multi_rows = {s: json.load(open(os.path.join(tmp, "result_rows", f"{s}_lm.json"))) for s in splits}
        assert multi_rows == one_rows, "multi-GPU row-list scores differ from single-GPU scores"
        print(f"{args.gpus}-GPU row-list input (num_of_data cut mid-hypothesis) identical to 1 GPU (bitwise)")
    # oracle check of the dev split
    sd = synth.random_init_state_dict(cfg, 10)
    from asr_rescoring_b200.tokenizer import SyntheticCharTokenizer
    tk = SyntheticCharTokenizer()
    nb = splits["dev"]
    hyps = {u: {h: tk.encode(s) for h, s in hs.items()} for u, hs in nb.hyps_text().items()}
    exp = pll_oracle.score_hyps(sd, cfg, hyps)
    worst = max(abs(one["dev"][u][h] - exp[u][h]) for u in exp for h in exp[u])
    assert worst <= 0.05, worst
    assert one["dev"][nb.utt_ids[1]]["hyp_3"] == 0
    print(f"dev_lm.json vs oracle: max |dPLL| {worst:.4f}")
    # combiner CLI
    os.makedirs(os.path.join(tmp, "rescore_out"))
    open(os.path.join(tmp, "rescore.yaml"), "w").write(f"""dev_am_path: "{tmp}/dev_hyps_score.json"
dev_lm_path: "{tmp}/result/dev_lm.json"
dev_hyps_text_path: "{tmp}/dev_hyps_text.json"
dev_ref_text_path: "{tmp}/dev_ref_text.json"
test_am_path: "{tmp}/test_hyps_score.json"
test_lm_path: "{tmp}/result/test_lm.json"
test_hyps_text_path: "{tmp}/test_hyps_text.json"
test_ref_text_path: "{tmp}/test_ref_text.json"
n_best: 5
output_path: "{tmp}/rescore_out"
""")
    out = subprocess.check_output([sys.executable, os.path.join(PKG, "rescore.py"), "--config", os.path.join(tmp, "rescore.yaml")],
                                  text=True)
    print(out.strip())
    c = rescore_oracle.config(5)
    d = rescore_oracle.dict_to_list
    with np.errstate(all="ignore"):
        bw, bc = rescore_oracle.find_best_weight(d(splits["dev"].hyps_score()), d(one["dev"]), d(splits["dev"].hyps_text()),
                                                 d(splits["dev"].ref_text()), c)
    log = open(os.path.join(tmp, "rescore_out", "rescore.log")).read()
    assert f"best_weight: {bw}" in log and f"dev cer: {bc}" in log, (bw, bc, log[-400:])
    print("rescore.log matches the oracle's best weight and dev CER")
    if args.gpus > 1:
        tail = lambda t: [l.split(" INFO ")[-1] for l in t.splitlines() if "best_weight:" in l or " cer: " in l]
        single = tail(log)
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                               "--master-addr", "127.0.0.1", "--master-port", "29579", os.path.join(PKG, "rescore.py"),
                               "--config", os.path.join(tmp, "rescore.yaml")])
        sharded = tail(open(os.path.join(tmp, "rescore_out", "rescore.log")).read())
        assert sharded == single and len(single) == 3, (single, sharded)
        print(f"{args.gpus}-GPU rescore.py (sharded sweep, all_reduce of the CER counts): {sharded} — identical to 1 GPU")


if __name__ == "__main__":
    main()
