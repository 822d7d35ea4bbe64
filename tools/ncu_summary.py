"""Summarises an .ncu-rep (read here, no GPU needed) into the small text files kept under profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_resident_v2"""
import csv, io, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "lts__t_sectors_op_write.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out + "_metrics.txt", "w") as f:
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        f.write(f"# kernel: {name}\n")
        for h, u, v in zip(hdr, units, r):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                f.write(f"{h} [{u}] = {v}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows[:10]):
    if "Source" in r:
        hdr, start = r, i + 1
        break
ie = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]
ws = [i for i, h in enumerate(hdr) if h.startswith("Warp Stall Sampling (All")][0]
si = hdr.index("Source")
data = [r for r in rows[start:] if len(r) > max(ie, ws)]
tot = sum(float(r[ie] or 0) for r in data) or 1.0
tots = sum(float(r[ws] or 0) for r in data) or 1.0
c, cs = collections.Counter(), collections.Counter()
for r in data:
    t = r[si].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    c[op] += float(r[ie] or 0)
    cs[op] += float(r[ws] or 0)
with open(out + "_sass_mix.txt", "w") as f:
    f.write(f"# warp instructions executed: {tot:.0f}; stall samples: {tots:.0f}\n# opcode  %executed  %stall-samples\n")
    for k, v in c.most_common(30):
        f.write(f"{k:12s} {100 * v / tot:6.2f} {100 * cs[k] / tots:6.2f}\n")
    f.write("# top stall lines: samples executed sass\n")
    for r in sorted(data, key=lambda r: -float(r[ws] or 0))[:30]:
        f.write(f"{float(r[ws] or 0):8.0f} {float(r[ie] or 0):12.0f}  {r[si][:100]}\n")
print("wrote", out + "_metrics.txt", out + "_sass_mix.txt")
