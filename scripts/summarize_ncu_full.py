"""Summarise an `ncu --set full` report (raw page exported with `ncu -i X.ncu-rep --page raw --csv`):
one line per captured launch with duration, DRAM traffic / throughput, tensor-pipe activity, occupancy, registers."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default="-"):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return r[i]
print("%-46s %9s %9s %9s %7s %8s %8s %7s %5s %7s" % ("kernel", "dur_us", "dramRd_MB", "dramWr_MB", "dram%", "tensor%", "hmmaOps%", "warps%", "regs", "smemKB"))
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    name = re.sub(r"\(anonymous namespace\)::|oat::|^void ", "", name)
    name = re.sub(r"\(.*$", "", name)[:46]
    dur = g(r, "gpu__time_duration.sum")
    print("%-46s %9.1f %9.1f %9.1f %7.1f %8.1f %8.1f %7.1f %5d %7.1f" % (
        name, dur / 1e3 if dur > 5e3 else dur,
        g(r, "dram__bytes_read.sum") / (1e6 if g(r, "dram__bytes_read.sum") > 1e4 else 1),
        g(r, "dram__bytes_write.sum") / (1e6 if g(r, "dram__bytes_write.sum") > 1e4 else 1),
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
        g(r, "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", 0.0),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        int(g(r, "launch__registers_per_thread")),
        (g(r, "launch__shared_mem_per_block_dynamic", 0.0) + g(r, "launch__shared_mem_per_block_static", 0.0)) / 1024.0))
