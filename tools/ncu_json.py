"""profiles/issue.json + profiles/traffic.json from a tools/ncu_summary.py text summary (one `ncu --set full` capture of
a C3 step): what bench.py attaches to its `roofline` object.  usage: python tools/ncu_json.py profiles/r2_ncu_full_c3.txt C3"""
import json, re, sys
from pathlib import Path
src, cfg = Path(sys.argv[1]), sys.argv[2]
blocks, cur = {}, None
for line in src.read_text().splitlines():
    if line.startswith("====="):
        name = re.sub(r"^void\s+", "", line[5:].strip())
        name = re.sub(r"[<(].*", "", name).replace("pgs::", "")
        cur = blocks.setdefault(name, {})
        cur["_n"] = cur.get("_n", 0) + 1
        if cur["_n"] > 1:
            cur = {}          # keep the first captured launch of every kernel
        continue
    m = re.match(r"\s+(\S+)\s+([0-9.,eE+-]+)\s*(\S*)", line)
    if m and cur is not None:
        v = float(m.group(2).replace(",", ""))
        unit = m.group(3)
        if unit == "Mbyte": v *= 1e6
        elif unit == "Kbyte": v *= 1e3
        elif unit == "Gbyte": v *= 1e9
        elif unit == "us": v *= 1e-3      # -> ms
        cur[m.group(1)] = v
    if line.strip().startswith("top stalls:") and cur is not None:
        cur["top_stalls"] = line.split(":", 1)[1].strip()
issue = {"_comment": f"per-kernel issue / pipe utilisation from the `ncu --set full` capture of one {cfg} step ({src}): "
                     "the render kernels are bound by instruction issue, not by HBM"}
traffic_p = Path("profiles/traffic.json")
traffic = json.loads(traffic_p.read_text()) if traffic_p.exists() else {}
traffic["_comment"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures ({cfg}: {src})"
for name, b in blocks.items():
    if "gpu__time_duration.sum" not in b or name.startswith("at::"):
        continue
    issue[name] = {
        "issue_active_pct": b.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": b.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warp_instructions": b.get("smsp__inst_executed.sum"),
        "registers": b.get("launch__registers_per_thread"),
        "warps_active_pct": b.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "shared_wavefronts": b.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "time_ms_under_ncu": b.get("gpu__time_duration.sum"),
        "top_stalls": b.get("top_stalls"), "source": str(src)}
    traffic.setdefault(name, {})[cfg] = int(b.get("dram__bytes_read.sum", 0) + b.get("dram__bytes_write.sum", 0))
if cfg == "C3":
    Path("profiles/issue.json").write_text(json.dumps(issue, indent=1))
traffic_p.write_text(json.dumps(traffic, indent=1))
print({k: v.get(cfg) for k, v in traffic.items() if isinstance(v, dict)})
