"""Print the handful of ncu metrics we track from a .ncu-rep (development aid)."""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, unit = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, unit))
    print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "smsp__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
            "smsp__warps_eligible.avg.per_cycle_active",
            # L2 by eviction class: the stream is loaded evict-first, the filter/table evict-last, the prefetch evict-normal
            "lts__t_sectors_lookup_hit.sum", "lts__t_sectors_lookup_miss.sum",
            "lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_miss.sum",
            "lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum",
            "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum",
            "lts__t_sectors_srcunit_ltcfabric_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_ltcfabric_op_read_lookup_miss.sum",
            # atomics (counter updates of the probe): requests and sectors at L2, set accesses at L1
            "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_tex_op_atom.sum",
            "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum",
            "nvlrx__bytes.sum", "nvltx__bytes.sum"]
    for k in keys:
        if k in d: print(f"  {k:82s} {d[k]:>16s} {u[k]}")
    st = [(float(v), k) for k, v in d.items() if "warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and v]
    for v, k in sorted(st, reverse=True)[:8]:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:8.3f} warps/issue")
