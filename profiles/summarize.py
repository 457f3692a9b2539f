"""Turn gpurun_out/ ncu captures into the tracked summaries under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.csv> "<command>"
    python profiles/summarize.py kernel <report.ncu-rep> <out.csv> "<command>"
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sm__cycles_elapsed.max",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    hdr, agg, lines = None, collections.defaultdict(lambda: [0, 0.0, []]), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            name = d["Kernel Name"].replace("<unnamed>::", "").replace(",", ";")[:80]
            if d["Metric Name"] == "gpu__time_duration.sum":
                v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
                agg[name][0] += 1
                agg[name][1] += v
                lines.append(f"{d['ID']},{name},{v:.1f}")
            elif "tensor" in d["Metric Name"]:
                agg[name][2].append(v)
    tot = sum(v[1] for v in agg.values())
    out = [f"# {cmd}", "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
           "kernel,launches,total_us,avg_us,share_pct,tensor_pipe_pct"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        t = sum(v[2]) / len(v[2]) if v[2] else 0.0
        out.append(f"{k},{v[0]},{v[1]:.1f},{v[1] / v[0]:.1f},{100 * v[1] / tot:.2f},{t:.1f}")
    out += ["", "id,kernel,us"] + lines
    open(dst, "w").write("\n".join(out) + "\n")


def kernel(src, dst, cmd):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    out = [f"# {cmd}", "metric,unit," + ",".join(f"launch{i}" for i in range(len(rr) - 2))]
    for i, name in enumerate(h):
        if name in WANT or ("issue_stalled" in name and name.endswith("per_issue_active.ratio") and "not_issued" not in name):
            out.append(",".join([name, rr[1][i]] + [r[i] for r in rr[2:]]))
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
