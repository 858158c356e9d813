"""profiles/r02_summary.md from the round-2 ncu exports and bench lines under profiles/."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_profiles as sp  # noqa: E402

P = os.path.join(ROOT, "profiles")


def raw_table(path, only=None, first_of_each=False):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if "__" in h]
    short = lambda h: h.replace(".avg.pct_of_peak_sustained_active", " %act").replace(".avg.pct_of_peak_sustained_elapsed", " %").replace(".sum", "")
    out = ["| kernel | grid | " + " | ".join(f"{short(hdr[i])} [{units[i]}]" if units[i] else short(hdr[i]) for i in keep) + " |",
           "|---|---|" + "---|" * len(keep)]
    seen = set()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("unnamed>::", "").replace("pllb::<", "")
        if only and not re.search(only, name):
            continue
        if first_of_each and name in seen:
            continue
        seen.add(name)
        out.append(f"| `{name}` | {r[hdr.index('Grid Size')]} | " + " | ".join(r[i][:9] for i in keep) + " |")
    return "\n".join(out)


def bench_row(name):
    d = json.load(open(os.path.join(P, name)))
    r = d.get("roofline", {})
    cb = d.get("cpu_baseline", {})
    return (f"| `{name}` | {d['config']['workload'].split(':')[0]} | {d['n_gpus']} | {d['dtype']} | {d['value']:.1f} | {d['ms_per_step']:.2f} | "
            f"{d['e2e']['value']:.1f} | {r.get('kind', r.get('kernel', ''))} {r.get('frac', 0):.3f} | "
            f"{r.get('gemm_family', {}).get('frac', float('nan')):.3f} | {cb.get('value', float('nan')):.2f} ({cb.get('kind', '-')}, {cb.get('cores', '-')} cores) | "
            f"{d.get('clocks', {}).get('sm_mhz')} MHz, {d.get('clocks', {}).get('power_w')} W, {d.get('clocks', {}).get('reasons')} |")


def main():
    out = ["# Round 2 — ncu on B200 (`--clock-control none`) and bench lines\n",
           "Raw exports: `r02_launches.csv` (every launch of `python bench.py --utts 400 --steps 1 --warmup 1 --no-cpu-baseline`: "
           "5 passes over 999 k packed rows), `r02_ncu_full_layer.csv` (five consecutive launches of one encoder layer, `--set full`), "
           "`r02_ncu_full_small_kernels.csv` (`--workload c1`: stage 1, head tail, stage 4), `r02_ncu_full_tokenizer.csv`, "
           "`r02_launches_att_tma.csv` / `r02_ncu_full_layer_att_tma.csv` (same with the opt-in TMA-fed attention kernel), "
           "`r02_sass_histogram.md`.  Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n",
           "## Bench lines (no profiler)\n",
           "| file | workload | GPUs | dtype | value hyps/s | ms/step | e2e hyps/s | roofline (dominant kernel) frac | GEMM family frac | CPU reference arm hyps/s | clocks |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    for f in sorted(os.listdir(P)):
        if (f.startswith("r02_bench_") or f.startswith("r02_final_bench_")) and f.endswith(".json") and "reference_arm" not in f:
            out.append(bench_row(f))
    out += ["", "## Launch list — default kernels (`r02_launches.csv`)\n", sp.launches(os.path.join(P, "r02_launches.csv")), "",
            "## One encoder layer, `--set full` (`r02_ncu_full_layer.csv`; M = 999 k rows)\n",
            raw_table(os.path.join(P, "r02_ncu_full_layer.csv")), "",
            "## Small kernels (`r02_ncu_full_small_kernels.csv`; c1 = 1 000 hypotheses, 14 k copies, 243 k rows — latency-bound sizes)\n",
            raw_table(os.path.join(P, "r02_ncu_full_small_kernels.csv"), first_of_each=True), "",
            "## Text front end (`r02_ncu_full_tokenizer.csv`; 71 760 hypotheses, 1.05 M code points)\n",
            raw_table(os.path.join(P, "r02_ncu_full_tokenizer.csv")), "",
            "## Opt-in TMA-fed attention kernel (`PLLB_ATT_TMA=1`)\n",
            sp.launches(os.path.join(P, "r02_launches_att_tma.csv")), "",
            raw_table(os.path.join(P, "r02_ncu_full_layer_att_tma.csv"), only="attention"), ""]
    fin = os.path.join(P, "r02_final_launches.csv")
    if os.path.exists(fin):
        out += ["## Launch list of the final build (`r02_final_launches.csv`, same command)\n", sp.launches(fin), ""]
    ln = os.path.join(P, "r02_ncu_full_gemm_ln_after_smem_params.md")
    if os.path.exists(ln):
        out += ["## Fused GEMM + LayerNorm kernels after the shared-memory parameters, `--set full`\n", open(ln).read(), ""]
    extra = os.path.join(P, "r02_notes.md")
    if os.path.exists(extra):
        out.append(open(extra).read())
    open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(out))
    print("wrote profiles/r02_summary.md")


if __name__ == "__main__":
    main()
