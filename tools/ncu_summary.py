"""Summarise an .ncu-rep (read on the CPU box): headline metrics, stall mix, per-region hot spots.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2 + li]
col = {h: i for i, h in enumerate(hdr)}
def get(name):
    return data[col[name]] if name in col else None
print("kernel:", get("Kernel Name")[:100])
for m in ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]:
    if m in col:
        print("  %-90s %s %s" % (m, data[col[m]], units[col[m]]))
pc = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(data[i] or 0) for h, i in col.items()
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
tot = sum(pc.values()) or 1
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(pc.items(), key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
if starts:
    h = srows[starts[li] if li < len(starts) else starts[0]]
    s0 = starts[li] if li < len(starts) else starts[0]
    ends = [i for i, r in enumerate(srows) if i > s0 and r and r[0] == "Kernel Name"]
    body = srows[s0 + 1:(ends[0] if ends else len(srows))]
    si, ii = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    def op(r):
        p = r[1].split()
        o = p[1] if p[0].startswith("@") else p[0]
        return o.split(".")[0]
    T = sum(int(r[si] or 0) for r in body) or 1
    ag, ex = collections.Counter(), collections.Counter()
    for r in body:
        ag[op(r)] += int(r[si] or 0); ex[op(r)] += int(r[ii] or 0)
    print("samples by opcode:", ", ".join("%s %.1f%% (x%d)" % (k, 100 * v / T, ex[k]) for k, v in ag.most_common(12)))
    N, B = len(body), 40
    step = N // B + 1
    for b in range(0, N, step):
        ch = body[b:b + step]
        s = sum(int(r[si] or 0) for r in ch)
        if 100 * s / T < 0.5:
            continue
        ops = collections.Counter(op(r) for r in ch)
        print("  instr %5d-%5d  %5.1f%%  exec/instr %9d  %s" % (b, b + len(ch), 100 * s / T, int(ch[len(ch) // 2][ii] or 0), ops.most_common(4)))
