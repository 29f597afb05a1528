"""profiles/ncu_r02_full_summary.{md,json} from the raw page of tools/gpu_ncu_full_r02.sh's capture (every launch of every tensor-core
kernel family in one cfg-2 training step, `ncu --set full --clock-control none --import-source on`).
usage: python tools/summarize_ncu_r02.py [gpurun_out/prof_r02_cfg2_raw.csv | gpurun_out/prof_r02_cfg2.ncu-rep]"""
import csv, io, json, os, subprocess, sys
src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_r02_cfg2_raw.csv"
if src.endswith(".ncu-rep"):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
else:
    txt = open(src).read()
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg"]


def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")


def to_us(r):
    v, u = val(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]]
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)


def to_mb(r, k):
    v, u = val(r, k), units[col[k]]
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)


by = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
    by.setdefault(name, []).append(r)
out_json = {}
md = ["# ncu --set full captures, round 02: every launch of the tensor-core kernel families in ONE cfg-2 training step\n",
      "Command: `tools/gpu_ncu_full_r02.sh` (`ncu --set full --clock-control none --import-source on -k regex:<families> -c 34 python tools/one_step.py cfg2 1`).",
      "UNet3D f_maps=32 depth=4, batch 2x1x128^3, forward + BCEDice + backward.  Times under ncu are cold-cache and serialised (replayed",
      "passes): read tensor-pipe %, DRAM bytes and L2->SM bytes, not absolute durations (those are in `profiles/bench_r02_*.json`).\n",
      "## all captured launches\n",
      "| kernel | grid | us | tensor pipe % | DRAM read MB | DRAM write MB | L2->SM MB | DRAM % of peak |", "|---|---|---:|---:|---:|---:|---:|---:|"]
for name, rs in by.items():
    for r in rs:
        md.append(f"| `{name[:44]}` | {r[col['Grid Size']] if 'Grid Size' in col else val(r, 'launch__grid_size'):} | {to_us(r):.1f} | "
                  f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {to_mb(r, 'dram__bytes_read.sum'):.1f} | "
                  f"{to_mb(r, 'dram__bytes_write.sum'):.1f} | {to_mb(r, 'l1tex__m_xbar2l1tex_read_bytes.sum'):.1f} | "
                  f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
md.append("\n## the longest launch of every kernel (full metric set used by the roofline)\n")
for name, rs in by.items():
    r = max(rs, key=to_us)
    key = name.replace("b200::", "")   # full name with template arguments: conv3_zs_kernel<32, 2> != conv3_zs_kernel<64, 2>
    out_json.setdefault(key, {})
    md += [f"### {name}\n", "| metric | value | unit |", "|---|---:|---|"]
    for w in WANT:
        if w in col:
            out_json[key][w] = {"value": r[col[w]].replace(",", ""), "unit": units[col[w]]}
            md.append(f"| `{w}` | {r[col[w]]} | {units[col[w]]} |")
    md.append("")
json.dump(out_json, open("profiles/ncu_r02_full_summary.json", "w"), indent=1)
open("profiles/ncu_r02_full_summary.md", "w").write("\n".join(md) + "\n")
print("\n".join(md[:60]))
