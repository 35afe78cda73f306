"""Summarise `ncu --page raw --csv` exports (one row per captured launch) into a short text table per kernel:
usage: python tools/ncu_csv_summary.py gpurun_out/x_raw.csv [more.csv ...] > profiles/rNN_name.txt
For every distinct kernel name (template arguments kept) the LONGEST captured launch is reported, plus how many launches were
captured and their total duration."""
import csv, sys, collections
WANT = [
 ("gpu__time_duration.sum", "duration"),
 ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
 ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks/SM)"),
 ("launch__occupancy_limit_registers", "occ limit regs (blocks/SM)"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
 ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
 ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
 ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA sub-pipe active %"),
 ("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "DMMA instructions"),
 ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
 ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts %"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle"),
 ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
 ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle"),
 ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle"),
 ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
 ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
 ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "local loads (spills)"),
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    def val(r, name):
        if name not in col: return None
        s = r[col[name]].replace(",", "")
        try: v = float(s)
        except ValueError: return None
        return v * UNIT.get(units[col[name]], 1)
    groups = collections.OrderedDict()
    for r in data:
        groups.setdefault(r[col["Kernel Name"]], []).append(r)
    print(f"==== {path}: {len(data)} captured launches")
    for name, rs in groups.items():
        best = max(rs, key=lambda r: val(r, "gpu__time_duration.sum") or 0)
        tot = sum(val(r, "gpu__time_duration.sum") or 0 for r in rs)
        print(f"--- {name[:150]}\n    launches captured {len(rs)}, total {tot*1e3:.3f} ms; longest launch:")
        for key, label in WANT:
            v = val(best, key)
            if v is None: continue
            if key == "gpu__time_duration.sum": print(f"    {label:34s} {v*1e6:.2f} us")
            elif "bytes" in key or "smem/block" in label: print(f"    {label:34s} {v/1e6:.3f} MB")
            else: print(f"    {label:34s} {v:.2f}")
        dr, dw, dt = val(best, "dram__bytes_read.sum"), val(best, "dram__bytes_write.sum"), val(best, "gpu__time_duration.sum")
        if dr is not None and dt: print(f"    {'dram GB/s (read+write)/duration':34s} {(dr+dw)/dt/1e9:.0f}")
