"""Turn ncu outputs (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.md --steps 2 --warmup 1
    python profiles/summarize.py kernel   gpurun_out/r1_prof_conv3d.ncu-rep profiles/r1_conv3d_ncu.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst, steps, warmup):
    """Aggregate the last `steps` steps of the launch list.  A step is the cyclic segment between two consecutive launches of
    the loss kernel `fo::mse_kernel<0>` (exactly one per step), so the split does not depend on how many set-up, warm-up or
    instrumented steps the bench ran around the timed ones."""
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    marks = [i for i, r in enumerate(rows) if "mse_kernel<0>" in r["Kernel Name"] or "mse_kernel<false>" in r["Kernel Name"]]
    if len(marks) > steps:
        timed = rows[marks[-steps - 1]:marks[-1]]
    else:   # fallback: equal split
        per_step = len(rows) // (steps + warmup)
        timed = rows[len(rows) - steps * per_step:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in timed:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none), {src}\n\n")
        out.write(f"`bench.py`; the last {steps} steps; {len(timed) // steps} launches/step; "
                  f"sum of kernel durations {tot / steps:.2f} ms/step (serialised, cold-cache: compare SHARES).\n\n")
        out.write("| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| `{k[:90]}` | {v[0] / steps:.0f} | {v[1] / steps:.3f} | {100 * v[1] / tot:.1f}% |\n")


WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    with open(dst, "w") as out:
        out.write(f"# ncu --set full --clock-control none, {src}\n\nkernel: `{name}`\n\n| metric | unit | value |\n|---|---|---:|\n")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                out.write(f"| {h} | {u} | {v} |\n")


def census(so, dst, title):
    """cuobjdump -sass instruction census per kernel: the mnemonics that prove the tcgen05 / TMEM / TMA paths."""
    sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "ATOMS", "ATOMG", "RED", "HMMA", "FFMA2", "STG", "LDG"]
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("void ", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["instrs"] += 1
            base = op.split(".")[0]
            if base == "UTCHMMA":
                kernels[cur]["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
            elif base in cols:
                kernels[cur][base] += 1
    with open(dst, "w") as out:
        out.write(f"# {title}\n\n")
        out.write("Counts of the instructions that prove the tcgen05 / TMEM / TMA paths (B200_PROFILING.md mnemonics) per kernel.\n"
                  "`UTCHMMA` = tcgen05.mma (column `.2CTA` = the cta_group::2 form), `UTMALDG` = TMA tensor load, `LDTM` = tcgen05.ld,\n"
                  "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier operations, `FFMA2` = packed fp32x2 FMA.  `HMMA` (mma.sync) does not occur:\n"
                  "every GEMM-shaped op runs on the 5th-generation tensor cores.  `python profiles/summarize.py census <so> <md>`.\n\n")
        out.write("| kernel | instrs | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols) + "\n")
        for k, c in kernels.items():
            out.write(f"| `{k[:70]}` | {c['instrs']} | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |\n")


def hbm(src_csv, time_txt, dst):
    """ncu per-launch DRAM bytes of tests/gpu_profile_hbm.py (8 clips) + the CUDA-event bandwidth table (32 clips)."""
    with open(src_csv) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    by_id = collections.OrderedDict()
    for r in rows:
        d = by_id.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d[r["Metric Name"] + "_unit"] = r["Metric Unit"]
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    with open(dst, "w") as out:
        out.write("# HBM-bound kernels: ncu DRAM traffic vs algorithmic bytes\n\n"
                  "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput...` over "
                  "`python tests/gpu_profile_hbm.py 8` (8 clips = 240 frames, production per-frame shapes).  DRAM traffic near the\n"
                  "algorithmic bytes (second table, scaled by 8 / 32) = no wasted re-reads.  Peak = 6534 GB/s (MEASURED_PEAKS.json).\n\n")
        out.write("| # | kernel | time (us) | DRAM read (MB) | DRAM write (MB) | DRAM total (MB) | DRAM % of peak |\n|---:|---|---:|---:|---:|---:|---:|\n")
        for i, d in by_id.items():
            if not d["name"].startswith("fo::") or "gpu__time_duration.sum" not in d:
                continue
            t = d["gpu__time_duration.sum"] * tscale.get(d["gpu__time_duration.sum_unit"], 1.0)
            rd = d.get("dram__bytes_read.sum", 0.0) * scale.get(d.get("dram__bytes_read.sum_unit", "Mbyte"), 1.0)
            wr = d.get("dram__bytes_write.sum", 0.0) * scale.get(d.get("dram__bytes_write.sum_unit", "Mbyte"), 1.0)
            pct = d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
            if t < 20:
                continue
            out.write(f"| {i} | `{d['name'][:60]}` | {t:.1f} | {rd:.1f} | {wr:.1f} | {rd + wr:.1f} | {pct:.1f} |\n")
        out.write("\n## CUDA-event bandwidth at 32 clips (960 frames), `python tests/gpu_profile_hbm.py 32 --time`\n\n```\n")
        out.write(open(time_txt).read())
        out.write("```\n")


if __name__ == "__main__":
    if sys.argv[1] == "census":
        census(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "SASS census (cuobjdump -sass, sm_100a)")
    elif sys.argv[1] == "hbm":
        hbm(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "launches":
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
        warmup = int(sys.argv[sys.argv.index("--warmup") + 1])
        launches(sys.argv[2], sys.argv[3], steps, warmup)
    else:
        kernel(sys.argv[2], sys.argv[3])
