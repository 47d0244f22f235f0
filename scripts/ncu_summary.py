"""Summaries for profiles/: (1) a launch-list CSV -> per-kernel share table; (2) a --set full report -> key metrics.
   python scripts/ncu_summary.py launches <csv>        |   python scripts/ncu_summary.py full <ncu-rep>"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("dfd::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:90]}` | {n} | {t / 1e6:.3f} | {t / tot * 100:.2f} % |")
    print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot / 1e6:.3f} | |")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"\n**{r[kn][:140]}**\n\n| metric | value | unit |\n|---|---:|---|")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"| {h} | {r[i]} | {units[i]} |")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
