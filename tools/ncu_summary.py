"""Summarise an .ncu-rep (one line per profiled launch): duration, tensor-pipe, issue, DRAM traffic, L2 traffic.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, subprocess, sys, io
KEYS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"), ("smsp__inst_executed.sum", "inst"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        parts = [f"{name} grid={grid}"]
        for k, label in KEYS:
            if k in col:
                v = r[col[k]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                parts.append(f"{label}={v}{units[col[k]] if label in ('dram_rd', 'dram_wr', 'l2_bytes') else ''}")
        print("  " + " ".join(parts))
