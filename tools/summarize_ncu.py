"""profiles/ncu_rNN_full_summary.{md,json} from the .ncu-rep files of tools/gpu_ncu_full.sh (+ the virtual-concat captures of
tools/gpu_ncu_up.sh when present).  usage: python tools/summarize_ncu.py r01"""
import csv, io, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg"]
REPS = [("wgrad_halo_kernel", f"gpurun_out/prof_{tag}_wgrad_halo_kernel.ncu-rep", "dec0.conv2 weight gradient, 32->32 @ 2x128^3"),
        ("conv3_halo_kernel", f"gpurun_out/prof_{tag}_conv3_halo_kernel.ncu-rep", "dec0.conv2 fprop, 32->32 @ 2x128^3"),
        ("conv3_igemm_kernel", f"gpurun_out/prof_{tag}_conv3_igemm_kernel.ncu-rep", "tap-loop kernel, a decoder layer"),
        ("up_wgrad", "gpurun_out/prof_up_wgrad.ncu-rep", "virtual-concat weight gradient of decoder 0 (captured BEFORE the lean issue loop / tap stacking)"),
        ("up_fwd_dgrad", "gpurun_out/prof_up_fwd_dgrad.ncu-rep", "virtual-concat phase conv (8 accumulators per CTA) and its transpose (4x4x4 stride-2), decoder 0")]
out_json, md = {}, [f"# ncu --set full captures, round {tag[1:]} (one launch each)\n",
                    "Commands: `tools/gpu_ncu_full.sh`, `tools/gpu_ncu_up.sh` (`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1`).",
                    "The `.ncu-rep` files stay in gpurun_out/ (scratch); the metrics the roofline uses are copied here.\n"]
for key, path, what in REPS:
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for li, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        k = key if li == 0 else f"{key}_{li}"
        out_json[k] = {}
        md += [f"## {name.split('(')[0]} — {what}\n", "| metric | value | unit |", "|---|---:|---|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out_json[k][w] = {"value": r[i].replace(",", ""), "unit": units[i]}
                md.append(f"| `{w}` | {r[i]} | {units[i]} |")
        md.append("")
json.dump(out_json, open(f"profiles/ncu_{tag}_full_summary.json", "w"), indent=1)
open(f"profiles/ncu_{tag}_full_summary.md", "w").write("\n".join(md) + "\n")
print("\n".join(md))
