"""ncu_extract.py — turn an .ncu-rep (ncu --set full) or a launch-list CSV (ncu --metrics gpu__time_duration.sum --csv) into the
markdown extracts committed under profiles/.
  python profiles/ncu_extract.py rep   gpurun_out/prof.ncu-rep  profiles/r1_ncu_x.md  "<title / command>" [samples traffic.json]
  python profiles/ncu_extract.py list  gpurun_out/launches.csv  profiles/r1_launches_x.md "<title>"
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active"]
ENTRY = {"k_grid_backward": "ngp_grid_encode_backward", "k_ffmlp_backward_fused<0, 0>": "ngp_ffmlp_backward",
         "k_ffmlp_backward_fused<0, 1>": "ngp_field_color_backward", "k_ffmlp_backward_dual<0, 0>": "ngp_ffmlp_backward_ex",
         "k_ffmlp_backward_dual<0, 1>": "ngp_field_color_backward_ex", "k_ffmlp_forward<1, 0, 1, 1": "ngp_field_sigma_forward",
         "k_ffmlp_forward<1, 0, 2, 2": "ngp_field_color_forward", "k_march_rays_train": "ngp_march_rays_train",
         "k_composite_train_fwd": "ngp_composite_rays_train_forward_mse", "k_composite_train_bwd": "ngp_composite_rays_train_backward"}
EXTRA = ["lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "smsp__inst_executed_op_global_red.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def rep(path, out, title, samples=None, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for h, i in col.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    lines = [f"# {title}", ""]
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        lines.append(f"## {name[:110]}\n")
        for m in METRICS + EXTRA:
            if m in col and r[col[m]] not in ("", "n/a"):
                lines.append(f"- {m}: {r[col[m]]} {units[col[m]]}")
        st = []
        for h, i in stall_cols:
            try:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        tot = sum(v for v, _ in st) or 1.0
        top = sorted(st, reverse=True)[:4]
        lines.append("- warp stall samples: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in top))
        lines.append("")
        if samples:
            for key, entry in ENTRY.items():
                if key in name and entry not in traffic:
                    b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                    traffic[entry] = {"dram_bytes_per_sample": b / samples, "samples": samples, "source": f"{out} ({path.split('/')[-1]})"}
                    red = "lts__t_sectors_srcunit_tex_op_red.sum"
                    if key == "k_grid_backward" and red in col:
                        traffic[entry]["red_ops_per_sample"] = float(r[col[red]].replace(",", "")) / samples
                        traffic[entry]["l2_red_sector_pct_of_peak"] = float(r[col["lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed"]].replace(",", ""))
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_json and traffic:
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    print("wrote", out, len(data), "kernels", list(traffic))


def launches(path, out, title):
    txt = open(path).read()
    start = txt.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r.get("Metric Unit", "us"), 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v in agg.values())
    lines = [f"# {title}", "", "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  Raw CSV: " + path.split("/")[-1].replace(".csv", "") + ".csv",
             "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {name} | {n} | {v:.1f} | {100 * v / tot:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else None, sys.argv[6] if len(sys.argv) > 6 else None)
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
