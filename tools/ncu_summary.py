"""Summarise an .ncu-rep (one kernel launch, --set full) into a small JSON + opcode table for profiles/."""
import collections, csv, io, json, re, subprocess, sys

rep, out_json, cells = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
summ = {"kernel": get.get("Kernel Name", ("?", ""))[0]}
for k in keys:
    if k in get:
        v, u = get[k]
        try:
            summ[k] = {"value": float(v), "unit": u}
        except ValueError:
            summ[k] = {"value": v, "unit": u}
stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(vals[i]) for i, h in enumerate(hdr)
          if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued") and vals[i].replace(".", "").isdigit()}
tot = sum(stalls.values()) or 1
summ["stall_pct"] = {k: round(100 * v / tot, 2) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]}
def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
if "dram__bytes_read.sum" in get:
    rd = float(get["dram__bytes_read.sum"][0]) * unit_scale(get["dram__bytes_read.sum"][1])
    wr = float(get["dram__bytes_write.sum"][0]) * unit_scale(get["dram__bytes_write.sum"][1])
    summ["dram_bytes_per_launch"] = rd + wr
    if cells:
        summ["cells_in_launch"] = cells
        summ["dram_bytes_per_cell"] = (rd + wr) / cells
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h2 = srows[1]; ia = h2.index("Instructions Executed"); isrc = h2.index("Source"); isamp = h2.index("# Samples")
    ops = collections.Counter(); samp = collections.Counter()
    for r in srows[2:]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        if not m:
            continue
        op = m.group(2).split(".")[0]
        ops[op] += float(r[ia]); samp[op] += float(r[isamp])
    ti = sum(ops.values()); ts = sum(samp.values()) or 1
    summ["warp_instructions_total"] = ti
    summ["opcode_share_pct"] = {op: {"inst": round(100 * c / ti, 2), "stall_samples": round(100 * samp[op] / ts, 2)} for op, c in ops.most_common(24)}
    if cells:
        summ["warp_instructions_per_cell"] = ti / cells
json.dump(summ, open(out_json, "w"), indent=1)
print(json.dumps(summ, indent=1)[:1500])
