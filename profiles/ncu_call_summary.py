"""Summarises every launch of an .ncu-rep capture (one encode call = 5 kernels) into a text file + raw CSV.
usage: python profiles/ncu_call_summary.py <rep> <out_prefix> "<description line>" """
import csv, subprocess, sys
rep, prefix, desc = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(prefix + "_raw.csv", "w").write(out)
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1.0, "us": 1e-3, "ns": 1e-6}
lines, tr, tw, tt = [desc], 0.0, 0.0, 0.0
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    lines.append("---- " + d["Kernel Name"].split("(")[0])
    for w in want:
        if w in d:
            lines.append("  %-88s %-10s %s" % (w, u[w], d[w]))
    st = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                 if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")), reverse=True)
    t = sum(s for s, _ in st) or 1
    lines.append("  stall samples: " + ", ".join("%s %.0f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * s / t) for s, k in st[:6]))
    tr += float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u["dram__bytes_read.sum"]]
    tw += float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u["dram__bytes_write.sum"]]
    tt += float(d["gpu__time_duration.sum"].replace(",", "")) * scale[u["gpu__time_duration.sum"]]
lines.append("==== one call: DRAM read %.3f GB + write %.3f GB = %.0f bytes; kernel time under ncu %.3f ms" % (tr / 1e9, tw / 1e9, tr + tw, tt))
open(prefix + ".txt", "w").write("\n".join(lines) + "\n")
print(lines[-1])
