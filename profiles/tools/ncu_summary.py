"""Per-kernel digest of an `ncu --set full … --page raw --csv` export (one row per profiled launch).

    python profiles/tools/ncu_summary.py profiles/r01_ncu_v4_record_gathers_raw.csv [more.csv …]

Prints what the roofline discussion in DESIGN.md needs: duration, DRAM bytes and %, L1 / L2 throughput %, global-load
requests -> sectors -> L1 data-pipe wavefronts, hit rates, issue utilisation, occupancy, registers and the
largest warp-stall reasons.  Launch durations under ncu are cold-cache and serialised: use them for shares and
ratios, never as bench numbers."""
import csv
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def digest(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name):
        return num(r[col[name]]) if name in col else None

    for r in data:
        name = r[col["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
        t_ns = get(r, "gpu__time_duration.sum")
        unit_t = units[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else "ns"
        t_us = t_ns / 1e3 if unit_t in ("ns", "nsecond") else t_ns * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit_t, 1e-3)
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd_b = rd * scale.get(units[col["dram__bytes_read.sum"]], 1) if rd is not None else None
        wr_b = wr * scale.get(units[col["dram__bytes_write.sum"]], 1) if wr is not None else None
        req = get(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
        sec = get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        hit = get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum")
        wf = get(r, "l1tex__data_pipe_lsu_wavefronts.sum") or get(r, "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg")
        print(f"== {name}   grid {r[col['Grid Size']]} x block {r[col['Block Size']]}")
        print(f"   duration {t_us:9.1f} us   registers {get(r, 'launch__registers_per_thread'):.0f}   "
              f"warps active {get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} %   "
              f"issue active {get(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} %")
        if rd_b is not None:
            print(f"   DRAM read {rd_b / 1e9:.3f} GB + write {wr_b / 1e9:.3f} GB = {(rd_b + wr_b) / 1e9:.3f} GB  "
                  f"({(rd_b + wr_b) / (t_us * 1e-6) / 1e12:.2f} TB/s under ncu)   "
                  f"dram {get(r, 'FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed') or float('nan'):.1f} %")
        print(f"   l1tex throughput {get(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} %   "
              f"L1 hit {get(r, 'l1tex__t_sector_hit_rate.pct'):.1f} %   L2 hit {get(r, 'lts__t_sector_hit_rate.pct'):.1f} %   "
              f"lts throughput {get(r, 'LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed') or float('nan'):.1f} %")
        if req and sec:
            line = f"   global loads: {req / 1e6:.2f} M requests -> {sec / 1e6:.2f} M sectors ({sec / req:.2f} per request"
            if hit is not None:
                line += f", {100 * hit / sec:.1f} % L1 hits"
            line += ")"
            print(line)
        stalls = sorted(((get(r, h) or 0.0, h.split("issue_stalled_")[1].split("_per_issue")[0]) for h in hdr
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")), reverse=True)[:4]
        print("   stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"# {p}")
        digest(p)
