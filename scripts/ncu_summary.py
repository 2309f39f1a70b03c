"""Summarise one kernel of an ncu report (--page raw --csv) as JSON: the metrics DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/<name>.ncu-rep [messages_in_launch] > profiles/<name>_summary.json
"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gcc__cache_requests_type_instruction.sum",
        "sm__icc_request_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "sm__cycles_elapsed.avg"]


def main():
    rep = sys.argv[1]
    msgs = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in WANT or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or h == "Kernel Name":
                try:
                    d[h] = float(vals[i].replace(",", ""))
                except ValueError:
                    d[h] = vals[i]
                if h in WANT and units[i]:
                    d[h + "__unit"] = units[i]
        stalls = {k.split("stalled_")[1].split("_per_issue")[0]: v for k, v in d.items()
                  if k.startswith("smsp__average_warps_issue_stalled_") and isinstance(v, float) and v > 0.005}
        d = {k: v for k, v in d.items() if not k.startswith("smsp__average_warps_issue_stalled_")}
        d["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if msgs:
            d["messages_in_launch"] = msgs
            d["warp_instructions_per_message"] = d.get("smsp__inst_executed.sum", 0) / msgs
            d["dram_bytes_per_message"] = 1e9 * (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / msgs \
                if d.get("dram__bytes_read.sum__unit") == "Gbyte" else None
            d["gpc_icache_requests_per_message"] = d.get("gcc__cache_requests_type_instruction.sum", 0) / msgs
        res.append(d)
    print(json.dumps(res if len(res) > 1 else res[0], indent=1))


if __name__ == "__main__":
    main()
