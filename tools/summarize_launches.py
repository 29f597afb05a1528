"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 1 --warmup 3`:
per-kernel share of ONE training step (the last step that starts with the input-statistics kernel).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import collections
import csv
import re
import sys


def main(path, title="UNet3D f32 d4, batch 2x1x128^3, fwd+BCEDice+bwd"):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    starts = [i for i, r in enumerate(rows) if "stats_ncdhw" in r["Kernel Name"]]
    s = starts[-1] if starts else len(rows) // 2  # tools/one_step_model.py runs exactly two identical steps
    e = len(rows)
    step = rows[s:e]
    agg = collections.OrderedDict()
    for r in step:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        v = float(r["Metric Value"].replace(",", ""))
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list, one training step ({title}), {len(step)} launches, "
          f"sum of device times {tot / 1e6:.2f} ms (cold-cache, serialised: read the SHARES)\n")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:80]}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% |")
    print("\n## tensor-core / direct conv launches in order (grid identifies the layer)\n")
    print("| us | grid | kernel |\n|---:|---|---|")
    for r in step:
        nm = r["Kernel Name"]
        if "igemm" in nm or "direct" in nm or "stem" in nm or "ring" in nm:
            short = nm.split("(")[0].replace("void ", "")[:60]
            print(f"| {float(r['Metric Value'].replace(',', '')) / 1e3:.1f} | {r['Grid Size']} | `{short}` |")


if __name__ == "__main__":
    main(*sys.argv[1:3])
