#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into one JSON file for profiles/.

    python tools/ncu_summary.py gpurun_out/r2_prof_*.ncu-rep > profiles/r2_ncu_full_summary.json

Per captured launch: duration, DRAM bytes read / written (the roofline `traffic`), DRAM throughput %, L2 hit rate,
L1 sectors per global-load request (1 request of a warp -> how many 32-byte sectors: 4 for a coalesced 128-byte row,
32 for a fully scattered 8-byte gather), occupancy, registers, shared memory and the largest warp-stall reasons."""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct2",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_ld_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "l1_global_ld_requests",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors_from_l1",
    "lts__t_sectors_op_read.sum": "l2_read_sectors",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
}


def to_num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return s


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        return [{"file": os.path.basename(path), "error": out.stderr[-300:]}]
    rows = list(csv.reader(io.StringIO(out.stdout)))
    header, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = dict(zip(header, r))
        e = {"file": os.path.basename(path), "kernel": d.get("Kernel Name", "")[:120], "id": d.get("ID")}
        for k, name in WANT.items():
            if k in d and d[k] != "":
                v = to_num(d[k])
                u = units[header.index(k)]
                if name == "duration_ns" and u in ("us", "usecond"):
                    v *= 1e3
                if name == "duration_ns" and u in ("ms", "msecond"):
                    v *= 1e6
                if name.endswith("_bytes") and isinstance(v, float):
                    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                e[name] = v
        stalls = {}
        for k in header:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d.get(k):
                name = k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")
                if name != "selected":
                    stalls[name] = to_num(d[k])
        # warps stalled for that reason per issued instruction (ncu "Warp State Statistics"): the larger, the more it limits
        e["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1] if isinstance(kv[1], float) else 0)[:5])
        if "dram_read_bytes" in e and "dram_write_bytes" in e:
            e["dram_bytes"] = e["dram_read_bytes"] + e["dram_write_bytes"]
            if e.get("duration_ns"):
                e["dram_gbs"] = e["dram_bytes"] / e["duration_ns"]
        if e.get("l1_global_ld_requests"):
            e["sectors_per_global_ld_request"] = e["l1_global_ld_sectors"] / e["l1_global_ld_requests"]
        res.append(e)
    return res


if __name__ == "__main__":
    allr = []
    for p in sys.argv[1:]:
        allr += summarise(p)
    print(json.dumps(allr, indent=1))
