"""Summarise .ncu-rep captures (ncu --set full) as one JSON line per profiled launch: duration, DRAM bytes, tensor / shared-memory
pipe utilisation, registers, cluster size.  Usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/x.jsonl"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed": "smem_pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "launch__registers_per_thread": "registers",
    "launch__cluster_dimension_x": "cluster",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__cycles_elapsed.avg": "cycles",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
}


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    head, units = r[0], r[1]
    for line in r[2:]:
        d = dict(zip(head, line))
        u = dict(zip(head, units))
        row = {"file": path.split("/")[-1], "kernel": d.get("Kernel Name", "")[:110]}
        for k, name in WANT.items():
            hits = [h for h in head if h == k or h.startswith(k)]
            if hits:
                row[name] = d[hits[0]] + (" " + u[hits[0]] if u[hits[0]] and name in ("duration", "dram_read", "dram_write", "smem_dynamic") else "")
        yield row


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for row in rows(p):
            print(json.dumps(row), flush=True)
