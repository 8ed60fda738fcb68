"""Summarise an `ncu --set full` report: per captured launch the metrics the roofline argument needs.

  python profiles/ncu_kernel_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"== {d['Kernel Name'][:110]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for k in KEYS:
        if k in d and d[k] not in ("", "n/a"):
            print(f"   {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
