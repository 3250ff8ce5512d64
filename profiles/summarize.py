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
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
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
        out.write(f"`bench.py --steps {steps} --warmup {warmup}`; timed steps only; {len(timed) // steps} launches/step; "
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


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
        warmup = int(sys.argv[sys.argv.index("--warmup") + 1])
        launches(sys.argv[2], sys.argv[3], steps, warmup)
    else:
        kernel(sys.argv[2], sys.argv[3])
