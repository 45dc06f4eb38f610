"""Key metrics of every kernel in an .ncu-rep (ncu -i X --page raw --csv) as a markdown table row / JSON."""
import csv, io, json, subprocess, sys
WANT = {
    "gpu__time_duration.sum": "time_us",
    "smsp__inst_executed.sum": "warp_inst",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_bytes.sum": "l2_bytes",
}
def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")].split("(")[0]}
        for i, n in enumerate(h):
            if n in WANT:
                try: v = float(r[i].replace(",", ""))
                except ValueError: continue
                u = units[i]
                if n.startswith("dram__bytes") or n.startswith("lts__t_bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if n == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                d[WANT[n]] = v
        tensor = [ (n, r[i]) for i, n in enumerate(h) if "tensor" in n and "pct" in n]
        d["tensor_metrics"] = {n: v for n, v in tensor[:6]}
        res.append(d)
    return res
if __name__ == "__main__":
    for p in sys.argv[1:]:
        for d in load(p):
            print(json.dumps(d))
