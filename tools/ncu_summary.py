"""Turn ncu outputs into the committed text summaries under profiles/.

  python tools/ncu_summary.py launches <launches.csv> [first_id]   -> per-kernel totals / shares of a --metrics gpu__time_duration.sum run
  python tools/ncu_summary.py full <report.ncu-rep>                -> selected metrics per captured launch of a --set full run
"""
import csv
import io
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "smsp__inst_executed.sum"]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("(anonymous namespace)::", "<unnamed>::")
    return name.strip()


def launches(path, first_id=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iid, iname, ival, iunit = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = {}
    n = 0
    for r in rows[1:]:
        if int(r[iid]) < first_id:
            continue
        v = float(r[ival].replace(",", ""))
        v = v / 1e3 if r[iunit] in ("ns", "nsecond") else (v * 1e3 if r[iunit] in ("ms", "msecond") else v)
        k = short(r[iname])
        t = tot.setdefault(k, [0.0, 0])
        t[0] += v; t[1] += 1; n += 1
    total = sum(v[0] for v in tot.values())
    print(f"total_us {total:.1f} kernel_launches {n}")
    for k, (us, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:10.1f} us {100 * us / total:5.1f}%  n={c:4d} avg={us / c:8.1f}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[iname])
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m} = {r[i]} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        full(sys.argv[2])
