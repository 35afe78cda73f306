"""Summarise an `ncu --set full` report of the contraction kernel into profiles/ (text) and profiles/traffic.json.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.txt "<command line that was profiled>" """
import csv, json, os, subprocess, sys
rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'gpu__time_duration.sum', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum']
with open(out, "w") as f:
    f.write(cmd + "\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w} [{units[i]}]: " + " | ".join(d[i] for d in data) + "\n")
print(open(out).read())
def val(name, row):
    i = hdr.index(name)
    v = float(row[i].replace(",", ""))
    u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
if "--traffic" in sys.argv:
    best = max(data, key=lambda r: val('gpu__time_duration.sum', r))
    t = val('dram__bytes_read.sum', best) + val('dram__bytes_write.sum', best)
    json.dump(t, open(os.path.join(os.path.dirname(out), "traffic.json"), "w"))
    print("traffic bytes (longest captured launch):", t)
