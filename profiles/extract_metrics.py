"""Key metrics per launch out of an `ncu --set full` report exported with
`ncu -i X.ncu-rep --page raw --csv > raw.csv`.  usage: extract_metrics.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | ms | DRAM read GB | DRAM write GB | traffic GB/s | dram % (ncu) | sm % | occupancy % | regs | grid | inst/launch |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for n, r in enumerate(data):
        def g(k):
            return r[ix[k]] if k in ix else ""
        name = g("Kernel Name").replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
        ms = float(g("gpu__time_duration.sum"))
        if units[ix["gpu__time_duration.sum"]] in ("us", "usecond"):
            ms /= 1e3
        def gb(k):
            v = float(g(k)); u = units[ix[k]]
            return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}[u]
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        print("| %d | %s | %.4f | %.3f | %.3f | %.0f | %.1f | %.1f | %.1f | %s | %s | %.3g |" % (
            n, name, ms, rd, wr, (rd + wr) / (ms * 1e-3) if ms > 0 else 0,
            float(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
            float(g("sm__throughput.avg.pct_of_peak_sustained_elapsed")),
            float(g("sm__warps_active.avg.pct_of_peak_sustained_active")), g("launch__registers_per_thread"),
            g("launch__grid_size"), float(g("smsp__inst_executed.sum"))))


if __name__ == "__main__":
    main(sys.argv[1])
