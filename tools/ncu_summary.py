"""Prints the metrics DESIGN.md / bench.py quote from an .ncu-rep (first captured launch):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import re
import subprocess
import sys

PAT = re.compile(r"^(gpu__time_duration\.sum|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                 r"sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                 r"sm__inst_executed_pipe_fma.*pct_of_peak_sustained_active|"
                 r"dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"lts__t_bytes\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?|"
                 r"l1tex__data_pipe_lsu_wavefronts\.sum\.pct_of_peak_sustained_elapsed|"
                 r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_.*)|"
                 r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                 r"sm__cycles_elapsed\.avg\.per_second|sm__cycles_elapsed\.max|smsp__inst_executed\.sum|"
                 r"smsp__sass_inst_executed_op_tmem_(ldt|stt)\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio)$")

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
print(f"kernel: {name}")
for h, u, v in zip(hdr, units, vals):
    if PAT.match(h):
        print(f"{h:86s} {u:18s} {v}")
