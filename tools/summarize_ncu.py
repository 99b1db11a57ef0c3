#!/usr/bin/env python
"""Turn the outputs of tools/ncu_bench.sh (gpurun_out/launches_<tag>.csv, gpurun_out/step_<tag>_full.ncu-rep and, when
present, batch_<tag>_full.ncu-rep / gemm_<tag>_full.ncu-rep) into the committed summaries under profiles/: the launch list
(shares), the metrics of the full captures, the executed-instruction and stall-sample shares per source function, and
step_kernel_traffic.json (read by bench.py for roofline.traffic).

    python tools/summarize_ncu.py [--tag r2]
"""
import argparse
import collections
import csv
import json
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
PROF = os.path.join(REPO, "profiles")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
           "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_issued.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_uniform.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__inst_executed.sum", "smsp__pcsamp_sample_count"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launch_list(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        per[r[4]][0] += 1
        per[r[4]][1] += float(r[-1]) / 1e6          # ns -> ms
    total = sum(v[1] for v in per.values())
    with open(os.path.join(PROF, f"{tag}_ncu_launch_list.txt"), "w") as f:
        f.write("# ncu launch list, kernels of libdia_b200.so only (-k regex:dia), of `python bench.py --steps 1 --warmup 1 "
                "--no-cpu-baseline --batch-utterances 8`\n# (cold-cache, serialised: compare SHARES).  One bench step = one "
                "3071-step generation = 24 launches of dia_step_kernel (128 steps each);\n# the batch field adds 8 utterances "
                "in one dia_batch_step_kernel stream; the one-time weight repacks are outside every timed\n# region.  "
                f"Total device time of these launches: {total:.1f} ms\nlaunches  ms_total  share  kernel\n")
        for k, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:8d} {ms:10.2f} {100 * ms / total:6.1f}%  {k}\n")


def metrics_only(tag, rep_name, title, out_name):
    rep = os.path.join(OUT, rep_name)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kv = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = [title]
    for m in METRICS:
        if m in kv:
            lines.append(f"{m:80s} {kv[m][1]:>16s} {kv[m][0]}")
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and kv[h][1] not in ("0", ""):
            lines.append(f"{h:80s} {kv[h][1]:>16s} {kv[h][0]}")
    with open(os.path.join(PROF, out_name), "w") as f:
        f.write("\n".join(lines) + "\n")


def full_capture(tag):
    rep = os.path.join(OUT, f"step_{tag}_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kv = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = ["# ncu --set full --clock-control none, dia_step_kernel, ONE launch of 64 decode steps from slot 1500 "
             "(tools/ncu_bench.sh; summary by tools/summarize_ncu.py)"]
    for m in METRICS:
        if m in kv:
            lines.append(f"{m:80s} {kv[m][1]:>16s} {kv[m][0]}")
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and kv[h][1] not in ("0", ""):
            lines.append(f"{h:80s} {kv[h][1]:>16s} {kv[h][0]}")
    rd = float(kv["dram__bytes_read.sum"][1]) * TO_BYTES[kv["dram__bytes_read.sum"][0]]
    wr = float(kv["dram__bytes_write.sum"][1]) * TO_BYTES[kv["dram__bytes_write.sum"][0]]
    # executed instructions / stall samples per source function (needs the same build in-tree for the line table)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    ia, iex, ismp = sh.index("Address"), sh.index("Instructions Executed"), sh.index("# Samples")
    data = [(int(r[ia], 16), int(r[iex]), int(r[ismp])) for r in srows[2:] if len(r) > iex and r[iex].isdigit()]
    try:
        off2line = line_table()
        cu = open(os.path.join(REPO, "dia_tts_prune_b200", "csrc", "step_kernel.cu")).read().split("\n")
        starts = []
        for i, l in enumerate(cu, 1):
            m = re.match(r"(?:static )?(?:template.*)?__device__.*?(\w+)\(", l)
            if m and not l.startswith(" "):
                starts.append((i, m.group(1)))
            if l.startswith('extern "C" __global__'):
                starts.append((i, "kernel body"))

        def bucket(k):
            if k is None:
                return "?"
            f, ln = k
            if f != "step_kernel.cu":
                return f
            name = "?"
            for s, n in starts:
                if s <= ln:
                    name = n
            return name
        ex, sm = collections.Counter(), collections.Counter()
        for a, e, n in data:
            b = bucket(off2line.get(a - data[0][0]))
            ex[b] += e
            sm[b] += n
        te, ts = sum(ex.values()), sum(sm.values())
        lines.append("# share of executed warp instructions / of stall samples per source function (inlined code is "
                     "attributed to the function it was written in)")
        for k, v in ex.most_common(12):
            lines.append(f"{100 * v / te:6.1f}% executed {100 * sm[k] / ts:6.1f}% samples   {k}")
    except Exception as e:                                      # pragma: no cover
        lines.append(f"# (no per-function table: {e})")
    with open(os.path.join(PROF, f"{tag}_ncu_step_kernel_full.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(PROF, "step_kernel_traffic.json"), "w") as f:
        json.dump({"dram_bytes_per_step": (rd + wr) / 64, "captured_steps_per_launch": 64, "first_slot": 1500,
                   "source": f"profiles/{tag}_ncu_step_kernel_full.txt"}, f)
        f.write("\n")


def line_table():
    import tempfile
    lib = os.path.join(REPO, "dia_tts_prune_b200", "csrc", "libdia_b200.so")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
        cub = [f for f in os.listdir(d) if f.startswith("step_kernel.")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    cur, infn, out = None, False, {}
    for line in txt.splitlines():
        if re.match(r"\s*\.text\.dia_step_kernel:", line):
            infn = True
            continue
        if re.match(r"\s*\.text\.", line):
            infn = False
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", line)
        if infn and m:
            out[int(m.group(1), 16)] = cur
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2")
    a = ap.parse_args()
    launch_list(a.tag)
    full_capture(a.tag)
    metrics_only(a.tag, f"batch_{a.tag}_full.ncu-rep",
                 "# ncu --set full --clock-control none, dia_batch_step_kernel, ONE launch of 16 decode steps of 8 utterances "
                 "from slot 1500 (tools/ncu_bench.sh)", f"{a.tag}_ncu_batch_kernel_full.txt")
    metrics_only(a.tag, f"gemm_{a.tag}_full.ncu-rep",
                 "# ncu --set full --clock-control none, dia_gemm_tcgen05_kernel, M = 2048, N = 16384, K = 2048 (tools/gemm_check.py)",
                 f"{a.tag}_ncu_tcgen05_gemm.txt")
    print("profiles updated")
