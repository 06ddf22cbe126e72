"""Condenses an ncu report (one kernel, `--set full`) into the few numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py <report.ncu-rep> [<out.json>]"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.max": "sm_cycles",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs_per_thread",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__inst_executed.sum": "warp_instructions",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma_pipe_pct",
    "sm__inst_executed_pipe_tmem.sum": "tmem_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[head.index("Kernel Name")]}
        for i, k in enumerate(head):
            if k in KEYS:
                d[KEYS[k]] = ("%s %s" % (r[i], units[i])).strip()
        out.append(d)
    txt = json.dumps({"report": rep.split("/")[-1], "kernels": out}, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    else:
        print(txt)


if __name__ == "__main__":
    main()
