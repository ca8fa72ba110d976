"""Turns ncu outputs brought back in gpurun_out/ into the text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_X.csv [N]  > profiles/rNN_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep      > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.ratio",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic"]


def launches(path, last=0):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    seq = [(r[kn].split("(")[0], float(r[mv].replace(",", ""))) for r in data
           if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
    if last:
        seq = seq[-last:]
        tot = sum(v for _, v in seq)
        print(f"# {path}: the last {last} launches = one warm step, {tot / 1e6:.3f} ms in total "
              "(ncu per-launch times are cold-cache and serialised: compare SHARES)")
        for k, v in seq:
            print(f"{v / 1e3:10.1f} us {100 * v / tot:5.1f}%  {k}")
        return
    agg = collections.OrderedDict()
    for k, v in seq:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES)")
    print(f"{'total_us':>10} {'share':>6} {'n':>4} {'avg_us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:4d} {v[1] / v[0] / 1e3:9.1f}  {k}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        kernel(sys.argv[2])
