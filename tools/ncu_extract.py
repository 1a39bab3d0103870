#!/usr/bin/env python
"""Turns an .ncu-rep capture into the small JSON / CSV summaries kept under profiles/ (run here, no GPU needed):
    python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/name   -> name_raw.csv (selected raw metrics) + name.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [k for k in KEEP if k in hdr]
    kernels = []
    with open(out + "_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[hdr.index(c)] for c in cols])
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            w.writerow([d[c] for c in cols])
            u = dict(zip(hdr, units))

            def num(key):
                try:
                    return float(d[key].replace(",", ""))
                except Exception:
                    return None

            def to_bytes(key):
                v, unit = num(key), u.get(key, "")
                if v is None:
                    return None
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)

            def to_s(key):
                v, unit = num(key), u.get(key, "")
                if v is None:
                    return None
                return v * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9,
                            "second": 1.0}.get(unit, 1e-3)

            rd, wr, t = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum"), to_s("gpu__time_duration.sum")
            kernels.append({"kernel": d["Kernel Name"], "duration_ms": None if t is None else 1e3 * t,
                            "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": None if rd is None else rd + wr,
                            "dram_GBs": None if not t else (rd + wr) / t / 1e9, "registers": num("launch__registers_per_thread"),
                            "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                            "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                            "local_sectors_st": num("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
                            "local_sectors_ld": num("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum")})
    with open(out + ".json", "w") as f:
        json.dump({"source": rep, "kernels": kernels}, f, indent=1)
    for k in kernels:
        print(json.dumps(k))


if __name__ == "__main__":
    main()
