#!/usr/bin/env python3
"""Reduce the ncu exports of tools/ncu_capture.sh (gpurun_out/prof_<tag>_<kernel>_{raw,src}.csv, launches_<tag>.csv) to the
summaries committed under profiles/:  <tag>_kernels.json (per kernel: duration, DRAM bytes, instructions, pipe
utilisation, issue, occupancy, registers, stall reasons, SASS opcode mix from the source page) and <tag>_kernels.md
(the same as a table + the launch-list shares).

    python tools/ncu_summary.py r2b [--src gpurun_out] [--out profiles]
"""
import argparse, collections, csv, glob, json, os, re

ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--src", default="gpurun_out")
ap.add_argument("--out", default="profiles")
a = ap.parse_args()

RAW = {
    "time_us": ("gpu__time_duration.sum", 1),
    "dram_read_bytes": ("dram__bytes_read.sum", 1), "dram_write_bytes": ("dram__bytes_write.sum", 1),
    "dram_pct_of_peak": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "warp_inst_executed": ("smsp__inst_executed.sum", 1),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "pipe_alu_pct": ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "pipe_fma_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "pipe_fmaheavy_pct": ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "registers_per_thread": ("launch__registers_per_thread", 1),
    "occupancy_limit_registers_blocks": ("launch__occupancy_limit_registers", 1),
    "grid": ("launch__grid_size", 1), "block": ("launch__block_size", 1),
    "dyn_smem_bytes": ("launch__shared_mem_per_block_dynamic", 1),
    "smem_wavefronts": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1),
    "smem_bank_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
}
STALLS = ["math_pipe_throttle", "long_scoreboard", "short_scoreboard", "barrier", "wait", "mio_throttle", "not_selected", "dispatch_stall", "lg_throttle", "no_instruction"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6, "Kbyte/block": 1e3, "byte/block": 1}


def num(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * UNIT.get(u, 1)


out = {}
for f in sorted(glob.glob(os.path.join(a.src, f"prof_{a.tag}_*_raw.csv"))):
    kern = re.match(rf"prof_{a.tag}_(.*)_raw\.csv", os.path.basename(f)).group(1)
    rows = list(csv.reader(open(f)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[h], rows[h + 1], rows[h + 2]
    col = {n: i for i, n in enumerate(names)}
    d = {"kernel": vals[col["Kernel Name"]]}
    for k, (m, _) in RAW.items():
        if m in col:
            d[k] = num(vals[col[m]], units[col[m]])
    d["stalls_per_issue"] = {s: num(vals[col[f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"]], "") for s in STALLS
                            if f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" in col}
    src = f.replace("_raw.csv", "_src.csv")
    if os.path.exists(src):
        rows = list(csv.reader(open(src)))
        hdr = rows[1]
        ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
        mix, tot = collections.Counter(), 0
        for r in rows[2:]:
            if len(r) <= ia or not r[ia].isdigit():
                continue
            s = r[isrc].strip()
            if s.startswith("@"):
                s = s.split(None, 1)[1] if len(s.split(None, 1)) > 1 else s
            op = s.split()[0] if s else "?"
            base = op.split(".")[0]
            if base == "IMAD":
                base = "IMAD.WIDE" if "WIDE" in op else ("IMAD.MOV" if any(t in op for t in ("MOV", "IADD", "SHL")) else "IMAD")
            mix[base] += int(r[ia]); tot += int(r[ia])
        d["sass_static_instructions"] = len(rows) - 2
        d["sass_mix_pct"] = {k: round(100.0 * v / tot, 1) for k, v in mix.most_common(12)} if tot else {}
    out[kern] = d

launches = os.path.join(a.src, f"launches_{a.tag}.csv")
shares = {}
if os.path.exists(launches):
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) > iv:
            k = re.sub(r"\(.*", "", r[ik]).replace("mk::", "").replace("void ", "")
            try:
                agg[k] += float(r[iv].replace(",", "")); cnt[k] += 1
            except ValueError:
                pass
    tot = sum(agg.values())
    shares = {k: {"launches": cnt[k], "total_ms": round(v / 1e6, 3), "share_pct": round(100 * v / tot, 2)} for k, v in agg.most_common()}

os.makedirs(a.out, exist_ok=True)
json.dump({"tag": a.tag, "kernels": out, "launch_list": shares,
           "note": "ncu --set full --clock-control none, one launch per kernel after two warm-up proofs (tools/ncu_capture.sh); launch list: serialised cold-cache device times of 3 proofs, compare shares"},
          open(os.path.join(a.out, f"{a.tag}_kernels.json"), "w"), indent=1)
with open(os.path.join(a.out, f"{a.tag}_kernels.md"), "w") as md:
    md.write(f"# ncu summary `{a.tag}` (tools/ncu_capture.sh + tools/ncu_summary.py)\n\n")
    md.write("| kernel | time | DRAM r+w | DRAM % | warp instr | issue % | ALU % | FMA-heavy % | warps active % | regs | top stalls (per issue) |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for k, d in out.items():
        st = sorted(d["stalls_per_issue"].items(), key=lambda kv: -(kv[1] or 0))[:3]
        md.write(f"| `{k}` | {d.get('time_us', 0) / 1e3:.3f} ms | {(d.get('dram_read_bytes', 0) + d.get('dram_write_bytes', 0)) / 1e6:.1f} MB | {d.get('dram_pct_of_peak', 0):.1f} | "
                 f"{d.get('warp_inst_executed', 0):.3g} | {d.get('issue_active_pct', 0):.1f} | {d.get('pipe_alu_pct', 0):.1f} | {d.get('pipe_fmaheavy_pct', 0):.1f} | "
                 f"{d.get('warps_active_pct', 0):.1f} | {int(d.get('registers_per_thread', 0))} | {', '.join(f'{n} {v:.2f}' for n, v in st)} |\n")
    md.write("\nSASS opcode mix of the executed instructions (source page):\n\n")
    for k, d in out.items():
        if d.get("sass_mix_pct"):
            md.write(f"* `{k}`: " + ", ".join(f"{o} {p}%" for o, p in d["sass_mix_pct"].items()) + "\n")
    if shares:
        md.write("\nLaunch list (3 proofs under ncu, serialised):\n\n| kernel | launches | total ms | share % |\n|---|---|---|---|\n")
        for k, v in shares.items():
            md.write(f"| `{k}` | {v['launches']} | {v['total_ms']} | {v['share_pct']} |\n")
print("wrote", os.path.join(a.out, f"{a.tag}_kernels.json"))
