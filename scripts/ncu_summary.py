#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into a small markdown summary for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_fwd_k32.ncu-rep "note" > profiles/r01_fwd.md
Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    start = out.find('"')
    return list(csv.reader(io.StringIO(out[start:])))


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary: {vals[col['Kernel Name']]}\n")
    print(f"source report: `{rep}` (ncu --set full --clock-control none --import-source on; one launch)\n")
    if note:
        print(note + "\n")
    print("| metric | value | unit |\n|---|---|---|")
    get = {}
    for m in METRICS:
        if m in col:
            get[m] = vals[col[m]]
            print(f"| {m} | {vals[col[m]]} | {units[col[m]]} |")

    def f(name):
        try:
            return float(get[name].replace(",", ""))
        except Exception:
            return None
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        tot = rd * scale.get(units[col["dram__bytes_read.sum"]], 1) + wr * scale.get(units[col["dram__bytes_write.sum"]], 1)
        print(f"\nDRAM traffic (read + write) per launch: **{tot / 1e9:.3f} GB**")
    # stall reasons / hottest instructions from the source page
    src = ncu_csv(rep, "source")
    h = None
    for i, r in enumerate(src):
        if r and r[0] == "Address":
            h = i
            break
    if h is not None:
        hd = src[h]
        c = {name: i for i, name in enumerate(hd)}
        body = []
        for r in src[h + 1:]:   # the first captured launch only (a report with several repeats the header)
            if r and r[0] == "Address":
                break
            if len(r) == len(hd):
                body.append(r)
        key = "Warp Stall Sampling (All Samples)"
        tot = sum(int(r[c[key]] or 0) for r in body) or 1
        stall_cols = [n for n in hd if n.startswith("stall_")]
        if stall_cols:
            agg = {n: sum(int(r[c[n]] or 0) for r in body) for n in stall_cols}
            top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
            print("\nwarp-stall samples by reason: " + ", ".join(f"{n[6:]} {100 * v / max(1, sum(agg.values())):.0f}%" for n, v in top))
        body.sort(key=lambda r: -int(r[c[key]] or 0))
        print("\n| hottest SASS (stall samples) | share |\n|---|---|")
        for r in body[:12]:
            print(f"| `{r[c['Source']].strip()}` | {100 * int(r[c[key]] or 0) / tot:.1f}% |")


if __name__ == "__main__":
    main()
