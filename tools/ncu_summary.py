"""Summarise an `ncu --set full` report: per captured launch, the metrics DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/x/full_orb.ncu-rep profiles/r1_ncu_full_orb.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__t_bytes.sum": "l1_bytes",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
}
UNIT_SCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    res = []
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]].split("(")[0], "id": int(r[col["ID"]])}
        for m, k in KEEP.items():
            if m not in col:
                continue
            raw = r[col[m]].replace(",", "")
            try:
                v = float(raw)
            except ValueError:
                continue
            v *= UNIT_SCALE.get(units[col[m]], 1.0)
            rec[k] = v
        if "duration" in rec and "dram_read" in rec:
            rec["dram_traffic"] = rec["dram_read"] + rec["dram_write"]
            rec["dram_gbs"] = rec["dram_traffic"] / rec["duration"] / 1e9
            rec["duration_us"] = rec.pop("duration") * 1e6
        res.append(rec)
    json.dump({"report": rep.split("/")[-1], "note": "ncu --set full --clock-control none; durations are serialised, "
               "cold-cache replays: compare shares", "launches": res}, open(out, "w"), indent=1)
    for r in res:
        print(f"{r['kernel'][:28]:28s} grid={int(r.get('grid', 0)):7d} {r.get('duration_us', 0):8.1f} us  dram {r.get('dram_traffic', 0) / 1e6:8.2f} MB "
              f"{r.get('dram_gbs', 0):7.1f} GB/s ({r.get('dram_pct', 0):4.1f}%)  issue {r.get('issue_active_pct', 0):4.1f}%  occ {r.get('occupancy_pct', 0):4.1f}%  "
              f"regs {int(r.get('registers', 0))}  winst {r.get('warp_instructions', 0) / 1e6:7.2f} M")


main()
