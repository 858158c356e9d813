"""Summarise an ncu launch list (gpu__time_duration) and a --set full capture into Markdown."""
import collections, csv, re, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for x in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", x["Kernel Name"]); name = re.sub(r".*::", "", name)
        v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    out = ["| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v / s < 0.0005: continue
        out.append(f"| `{k}` | {cnt[k]} | {v:.2f} | {100*v/s:.1f}% | {1000*v/cnt[k]:.1f} |")
    out.append(f"| total | {sum(cnt.values())} | {s:.2f} | | |")
    return "\n".join(out)

def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    out = ["| " + " | ".join(hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx) + " |", "|" + "---|" * len(idx)]
    for r in rows[2:]:
        cells = []
        for i in idx:
            v = r[i]
            if hdr[i] == "Kernel Name":
                m = re.search(r"(\w+<\d+>|\w+)\(", v)
                v = m.group(1) if m else v[:40]
            cells.append(v)
        out.append("| " + " | ".join(cells) + " |")
    return "\n".join(out)

if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    print(launches(path) if kind == "launches" else full(path))
