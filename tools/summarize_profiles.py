"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_profiles.py r1       # reads gpurun_out/launches_r1.csv, prof_conv_r1.ncu-rep, prof_corr_r1.ncu-rep

Runs here (no GPU): `ncu -i <rep> --page raw --csv` parses the reports.
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum",
]


def launches_summary(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.isfile(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    per_launch = []
    for r in rows[hdr + 1:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0].replace("void ", "").replace("tsnet::", "")
        ns = float(r[-1])
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ns
        a[1] += 1
        per_launch.append((r[0], name, r[7], r[8], ns))
    # tools/one_forward.py runs several identical forwards: keep the last one (the forward starts with the stem loader)
    starts = [i for i, pl in enumerate(per_launch) if pl[1].startswith("stem_taps_kernel")]
    if len(starts) >= 2:
        first = starts[-2]          # two stem loaders per forward: img_enc then lbl_enc
        per_launch = per_launch[first:]
        agg = collections.OrderedDict()
        for _, name, _, _, ns in per_launch:
            a = agg.setdefault(name, [0.0, 0])
            a[0] += ns
            a[1] += 1
    tot = sum(v[0] for v in agg.values())
    out = [f"# Launch list of one steady-state forward ({tag})", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` over one forward of `bench.py`'s workload "
           "(bs=32, n_source=3, n_blocks=4).  Times are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| `{k}` | {v[1]} | {v[0] / 1e6:.3f} | {v[0] / tot:.3f} |")
    out += [f"| **total** | {sum(v[1] for v in agg.values())} | {tot / 1e6:.3f} | 1.000 |", "",
            "<details><summary>every launch</summary>", "", "| id | kernel | block | grid | us |", "|---|---|---|---|---|"]
    for i, n, blk, grd, ns in per_launch:
        out.append(f"| {i} | `{n}` | {blk} | {grd} | {ns / 1e3:.1f} |")
    out += ["", "</details>", ""]
    return "\n".join(out)


def raw_metrics(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return None
    h, units, v = rows[0], rows[1], rows[2]
    d = {}
    for i, n in enumerate(h):
        d[n] = (v[i], units[i])
    return d


def kernel_summary(tag, which, title, algo_note):
    rep = os.path.join(OUT, f"prof_{which}_{tag}.ncu-rep")
    if not os.path.isfile(rep):
        return None
    d = raw_metrics(rep)
    if d is None:
        return None
    out = [f"# {title} ({tag})", "", f"`ncu --set full --clock-control none --import-source on` — report `prof_{which}_{tag}.ncu-rep` "
           "(kept in gpurun_out/, not tracked: binary).", "", f"Kernel: `{d.get('Kernel Name', ('?', ''))[0]}`", "",
           "| metric | value | unit |", "|---|---|---|"]
    for k in RAW_KEYS:
        if k in d:
            out.append(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
    try:
        rd = float(d["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
        out += ["", f"DRAM traffic per launch (read + write) = **{(rd + wr) / 1e6:.1f} MB**. {algo_note}"]
    except (KeyError, ValueError):
        pass
    out.append("")
    return "\n".join(out)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    for name, text in (
        (f"{tag}_launches.md", launches_summary(tag)),
        (f"{tag}_conv_gemm.md", kernel_summary(tag, "conv", "conv_gemm_kernel<256>, 512->512 3x3 @32x32, 96 samples",
                                               "Algorithmic bytes: A tap source 2 x 96*34*34*512*2 B = 227 MB + raw output "
                                               "96*1024*512*4 B = 201 MB + weights 9.4 MB.")),
        (f"{tag}_corr_tiles.md", kernel_summary(tag, "corr", "corr_tile_kernel (tcgen05 similarity tiles + softmax partial states)",
                                                "Algorithmic bytes of the whole correlation chain (SURVEY section 8d): "
                                                "10,502,144 B/frame x 32 frames = 336.1 MB; this kernel reads the 16-bit "
                                                "hi/lo operands (268 MB) through L2 several times and writes 12.6 MB of states.")),
        (f"{tag}_corr_finish.md", kernel_summary(tag, "finish", "warp_mean_taps_kernel<true> = corr_finish (state merge + "
                                                 "4-tap grid_sample + source mean -> map_conv operand)",
                                                 "Reads the fp32 source features (201 MB, 4 taps each: L2-amplified) and "
                                                 "writes the 67 MB hi/lo operand.")),
        (f"{tag}_stem_vr.md", kernel_summary(tag, "stem", "conv_gemm_vr_kernel (kw-folded 7x7 stem, 96 x 256 x 256, Cout 64)",
                                             "Algorithmic bytes: tap source 1.65 GB + raw output 1.61 GB + statistics 0.1 GB.")),
    ):
        if text:
            open(os.path.join(PROF, name), "w").write(text)
            print("wrote", name)


if __name__ == "__main__":
    main()
