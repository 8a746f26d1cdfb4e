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
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
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
    """every kernel of an ncu report as {metric: (value, unit)} dicts"""
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return []
    h, units = rows[0], rows[1]
    return [{n: (v[i], units[i]) for i, n in enumerate(h)} for v in rows[2:]]


_UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}


def kernel_summary(tag, which, title, algo_note):
    rep = os.path.join(OUT, f"prof_{which}_{tag}.ncu-rep")
    if not os.path.isfile(rep):
        return None
    ks = raw_metrics(rep)
    if not ks:
        return None
    out = [f"# {title} ({tag})", "", f"`ncu --set full --clock-control none --import-source on` -- report "
           f"`prof_{which}_{tag}.ncu-rep` (kept in gpurun_out/, not tracked: binary).", ""]
    total = 0.0
    for d in ks:
        out += [f"Kernel: `{d.get('Kernel Name', ('?', ''))[0]}`  grid {d.get('Grid Size', ('?', ''))[0]}", "",
                "| metric | value | unit |", "|---|---|---|"]
        for k in RAW_KEYS:
            if k in d:
                out.append(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
        try:
            rd = float(d["dram__bytes_read.sum"][0]) * _UNIT[d["dram__bytes_read.sum"][1]]
            wr = float(d["dram__bytes_write.sum"][0]) * _UNIT[d["dram__bytes_write.sum"][1]]
            total += rd + wr
            out += ["", f"DRAM traffic of this launch (read + write) = **{(rd + wr) / 1e6:.1f} MB**.", ""]
        except (KeyError, ValueError):
            out.append("")
    if len(ks) > 1:
        out += [f"DRAM traffic of all {len(ks)} launches = **{total / 1e6:.1f} MB**.", ""]
    out += [algo_note, ""]
    return "\n".join(out)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    os.makedirs(PROF, exist_ok=True)
    for name, text in (
        (f"{tag}_launches.md", launches_summary(tag)),
        (f"{tag}_wino_gemm.md", kernel_summary(
            tag, "winogemm", "conv_gemm2_kernel in batched-plane mode = the 16 Winograd plane GEMMs of one layer",
            "Algorithmic bytes of the 512->512 layer over 96 samples: V 96*16*256*512*4 B = 805 MB read (each A tile is read "
            "by two channel slabs: the second read is an L2 hit) + M 805 MB written + 16 MB of weights; the 1024->1024 layer "
            "is twice that in each direction.  Executed FLOPs = 16/36 of the 3x3 convolution's.")),
        (f"{tag}_bridge.md", kernel_summary(
            tag, "bridge", "wino_bridge_kernel (output transform + InstanceNorm + ReLU / residual + input transform)",
            "Algorithmic bytes per launch over X samples of C channels: M X*16*256*C*4 B read + V the same written "
            "(+ X*1024*C*4 B residual read and act_out written on the second convolution of a ResnetBlock).")),
        (f"{tag}_wino_input.md", kernel_summary(tag, "winoin", "wino_input_kernel (pass T: norm + pad + B^T d B + split)", "")),
        (f"{tag}_wino_output.md", kernel_summary(tag, "winoout", "wino_output_kernel (pass I: A^T M A + bias + statistics)", "")),
        (f"{tag}_corr_chain_dram.md", kernel_summary(
            tag, "corr", "every kernel of the correlation chain of one forward (bs=32, n_source=3)",
            "Algorithmic bytes of the chain (SURVEY section 8d): 10,502,144 B/frame x 32 frames = 336.1 MB.  The sources' "
            "operand rows (201 MB) are written by the last img_enc bridge pass and are not part of these launches.")),
        (f"{tag}_stem_direct.md", kernel_summary(
            tag, "stem", "conv_gemm_vr_kernel<true> (direct-input stem, 96 x 256 x 256, Cout 64; opt-in)",
            "Algorithmic bytes: raw inputs 96*5*65536*4 B = 126 MB + raw output 1.61 GB + statistics; the materialised "
            "path reads a 1.65 GB tap source on top (3.31 GB per launch, plus the 1.65 GB stem_taps write).")),
    ):
        if text:
            open(os.path.join(PROF, name), "w").write(text)
            print("wrote", name)


if __name__ == "__main__":
    main()
