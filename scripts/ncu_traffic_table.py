"""ncu raw-page CSV of scripts/gpu_kernels_once.py -> JSON table {case: {kernel, dur_us, dram_bytes, tensor_pct, ...}}.
The cases are matched to the captured launches in order (helper kernels - combine / finalize - fold into their parent)."""
import csv, json, re, sys
CASES = ["gemm_fwd_qkv", "gemm_fwd_proj", "gemm_fwd_fc1", "gemm_fwd_fc2", "gemm_dgrad_fc2", "gemm_dgrad_fc1",
         "gemm_dgrad_qkv", "gemm_dgrad_proj", "gemm_wgrad_proj", "gemm_wgrad_qkv", "gemm_wgrad_fc1", "gemm_wgrad_fc2",
         "attn_space_fwd", "attn_space_bwd", "attn_time_fwd", "attn_time_bwd", "layernorm_fwd", "layernorm_bwd", "colsum",
         "gemm_dgrad_proj_rowdot", "attn_space_bwd_delta", "attn_time_bwd_delta"]
HELPERS = ("combine", "finalize")
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    v = r[col[name]].replace(",", "") if name in col else ""
    return float(v) if v not in ("", "n/a") else 0.0
unit = {h: rows[1][i] for h, i in col.items()}
def to_bytes(r, name):
    u = unit[name].lower()
    k = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    return num(r, name) * k
def to_us(r, name):
    u = unit[name].lower()
    k = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[u]
    return num(r, name) * k
out, ci = {}, -1
for r in rows[2:]:
    name = re.sub(r"\(.*$", "", re.sub(r"\(anonymous namespace\)::|oat::|^void ", "", r[col["Kernel Name"]]))
    helper = any(h in name for h in HELPERS)
    if not helper:
        ci += 1
    case = CASES[ci]
    d = out.setdefault(case, {"kernels": [], "dur_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    d["kernels"].append(name)
    d["dur_us"] += to_us(r, "gpu__time_duration.sum")
    d["dram_read_bytes"] += to_bytes(r, "dram__bytes_read.sum")
    d["dram_write_bytes"] += to_bytes(r, "dram__bytes_write.sum")
    if not helper:
        d["tensor_pipe_pct"] = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        d["dram_pct"] = num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        d["registers"] = int(num(r, "launch__registers_per_thread"))
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: round((v["dram_read_bytes"] + v["dram_write_bytes"]) / 1e6, 1) for k, v in out.items()}))
