"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run here, no GPU).

    python tools/summarize_profiles.py launches <csv> <out.md> <title>
    python tools/summarize_profiles.py ncu <out.md> <title> <rep> [<rep> ...]
    python tools/summarize_profiles.py traffic <workload> <obs.rep> <gram.rep>
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg.per_second",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        res.append(d)
    return res


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    names = [r[4] for r in rows]
    last = max(i for i, n in enumerate(names) if "k_prep" in n)
    items, tot = [], 0.0
    for r in rows[last:]:
        v, u = float(r[-1]), r[-2]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        items.append((r[4].split("(")[0], v))
        tot += v
    with open(out, "w") as f:
        f.write("# %s\n\n(last repetition; cold-cache, serialised under ncu: compare SHARES)\n\n" % title)
        f.write("| kernel | us | share |\n|---|---|---|\n")
        for n, v in items:
            f.write("| %s | %.1f | %.1f%% |\n" % (n, v, 100 * v / tot))
        f.write("| total | %.1f | |\n" % tot)


def launches_agg(path, out, title):
    """Whole-command launch list aggregated by kernel: count, mean duration, share of the summed
    duration of OUR kernels (lrvb::*)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg, order = {}, []
    for r in rows:
        v, u = float(r[-1]), r[-2]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        n = r[4].split("(")[0]
        if n not in agg:
            agg[n] = [0, 0.0]
            order.append(n)
        agg[n][0] += 1
        agg[n][1] += v
    ours = sum(t for n, (c, t) in agg.items() if "lrvb::" in n)
    with open(out, "w") as f:
        f.write("# %s\n\n(every launch of the command; cold-cache, serialised under ncu: compare SHARES)\n\n" % title)
        f.write("| kernel | launches | mean us | share of lrvb:: time |\n|---|---|---|---|\n")
        for n in order:
            c, t = agg[n]
            f.write("| %s | %d | %.1f | %s |\n" % (n, c, t / c, ("%.1f%%" % (100 * t / ours)) if "lrvb::" in n else "-"))


def ncu(out, title, reps):
    with open(out, "w") as f:
        f.write("# %s\n" % title)
        for rep in reps:
            for d in raw(rep):
                f.write("\n## %s  (%s)\n\n| metric | value |\n|---|---|\n" % (
                    d.get("Kernel Name", ("?", ""))[0], os.path.basename(rep)))
                for m in METRICS:
                    if m in d:
                        f.write("| %s | %s %s |\n" % (m, d[m][0], d[m][1]))


def dram(rep):
    d = raw(rep)[0]
    def b(key):
        v, u = d[key]
        s = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v) * s
    return int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))


def traffic(workload, obs_rep, gram_rep):
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    t[workload] = {"k_obs_dram_bytes": dram(obs_rep), "k_gram_dram_bytes": dram(gram_rep),
                   "source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full: %s, %s" % (
                       os.path.basename(obs_rep), os.path.basename(gram_rep))}
    json.dump(t, open(path, "w"), indent=1)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(*sys.argv[2:5])
    elif cmd == "launches_agg":
        launches_agg(*sys.argv[2:5])
    elif cmd == "ncu":
        ncu(sys.argv[2], sys.argv[3], sys.argv[4:])
    elif cmd == "traffic":
        traffic(*sys.argv[2:5])
