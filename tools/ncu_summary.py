"""Summarise an `ncu --set full` capture for profiles/ and for bench.py's roofline.traffic.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv
    python tools/ncu_summary.py gpurun_out/x_raw.csv <bench kernel name> "<launch description>" [algorithmic_bytes] \
        > profiles/r2_dominant_ncu.json

Picks the longest launch whose kernel name contains the given substring (or the longest launch at all) and prints
one JSON object: DRAM bytes read / written, duration, L2 hit rate, pipe utilisations, bank-conflict share -- the metric
names of /opt/skills/guides/B200_PROFILING.md."""
import csv
import json
import sys


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except Exception:
        return None


def main():
    path, bench_kernel = sys.argv[1], sys.argv[2]
    desc = sys.argv[3] if len(sys.argv) > 3 else ""
    algo = float(sys.argv[4]) if len(sys.argv) > 4 else None
    match = sys.argv[5] if len(sys.argv) > 5 else bench_kernel
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {n: i for i, n in enumerate(hdr)}
    unit = {n: units[i] for n, i in col.items()}

    def scale(name, v):
        u = unit.get(name, "").lower()
        f = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9,
             "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u, 1.0)
        return v * f if v is not None else None

    best = None
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        if match and match not in name:
            continue
        d = scale("gpu__time_duration.sum", num(r[col["gpu__time_duration.sum"]]))
        if d is None:
            continue
        if best is None or d > best[0]:
            best = (d, r)
    if best is None:
        sys.exit("no launch matching %r" % match)
    d, r = best

    def get(name, scaled=True):
        if name not in col:
            return None
        v = num(r[col[name]])
        return scale(name, v) if scaled else v

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out = {
        "bench_kernel": bench_kernel, "kernel_name": r[col["Kernel Name"]][:160], "launch": desc, "source_csv": path,
        "duration_us": d * 1e6, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes": (rd or 0) + (wr or 0), "algorithmic_bytes": algo,
        "traffic_over_algorithmic": ((rd or 0) + (wr or 0)) / algo if algo else None,
        "grid": r[col["Grid Size"]] if "Grid Size" in col else None,
        "block": r[col["Block Size"]] if "Block Size" in col else None,
    }
    for key, name in (("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
                      ("tensor_pipe_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
                      ("tensor_pipe_pct_alt", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
                      ("sm_busy_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                      ("lsu_shared_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                      ("shared_ld_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
                      ("shared_st_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
                      ("achieved_occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                      ("registers_per_thread", "launch__registers_per_thread")):
        v = get(name, scaled=False)
        if v is not None:
            out[key] = v
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
