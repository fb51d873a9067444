"""Summarise an .ncu-rep (ncu --set full) into the handful of counters the roofline argument needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv, io, subprocess, sys
KEYS = """gpu__time_duration.sum
launch__grid_size launch__block_size launch__registers_per_thread launch__occupancy_limit_registers launch__waves_per_multiprocessor
sm__warps_active.avg.pct_of_peak_sustained_active sm__maximum_warps_per_active_cycle_pct
dram__bytes_read.sum dram__bytes_write.sum dram__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_bytes.sum lts__throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct lts__t_sectors_op_read.sum lts__t_sectors_op_write.sum lts__t_sectors_op_red.sum lts__t_sectors_op_atom.sum
l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tensor.sum sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.sum
smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio smsp__average_warp_latency_issue_stalled_lg_throttle.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_membar_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    out.append(f"== {d.get('Kernel Name','?')}  grid {d.get('Grid Size','?')} block {d.get('Block Size','?')}")
    for k in KEYS:
        if k in d:
            out.append(f"  {k:95s} {units[hdr.index(k)]:14s} {d[k]}")
    try:
        t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
        def b(k):
            u = units[hdr.index(k)]; v = float(d[k].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        l2 = f", L2 traffic {b('lts__t_bytes.sum')/1e9:.3f} GB ({b('lts__t_bytes.sum')/t/1e9:.0f} GB/s)" if "lts__t_bytes.sum" in hdr else ""
        out.append(f"  -> duration {t*1e3:.3f} ms, DRAM traffic {tr/1e9:.3f} GB ({tr/t/1e9:.0f} GB/s){l2}")
    except Exception as e:
        out.append(f"  (derived failed: {e})")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
