#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a small per-kernel CSV summary that can be committed:
duration, tensor-pipe %, DRAM bytes/throughput, L2 throughput, registers, achieved occupancy."""
import csv
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "gpu_dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tc_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_pct",
}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [f"{v} [{units[cols[k]]}]" if k in cols else v for k, v in WANT.items()])
        for r in rows[2:]:
            name = r[cols["Kernel Name"]][:60]
            w.writerow([r[cols["ID"]], name] + [r[cols[k]] if k in cols else "" for k in WANT])
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
