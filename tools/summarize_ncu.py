"""Text summary of an ncu report for profiles/: key metrics, stall reasons, opcode mix and (when the
matching cubin is given) SASS size / samples per source region.
usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main(rep):
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, vals = raw[0], raw[-1]
    d = dict(zip(hdr, vals))
    f = lambda k: float(d[k].replace(",", "")) if d.get(k, "").replace(",", "").replace(".", "").replace("-", "").isdigit() else None
    print("report:", rep)
    print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct",
            "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    for k in keys:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], raw[1][hdr.index(k)] if len(raw) > 2 else ""))
    st = [(k.split("stalled_")[1], f(k)) for k in hdr if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and f(k)]
    tot = sum(v for _, v in st)
    print("warp stall samples (all):")
    for k, v in sorted(st, key=lambda x: -x[1])[:9]:
        print("  %-28s %6.1f%%" % (k, 100 * v / tot))
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    h2, data = src[1], src[2:]
    ia, isrc = h2.index("Instructions Executed"), h2.index("Source")
    byop = collections.Counter(); total = 0
    for r in data:
        if not r[ia].isdigit():
            continue
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += int(r[ia]); total += int(r[ia])
    print("SASS instructions in kernel: %d ; executed warp instructions: %d" % (len(data), total))
    print("opcode mix:", ", ".join("%s %.1f%%" % (o, 100 * n / total) for o, n in byop.most_common(12)))


def dram_bytes(rep):
    """dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, in bytes"""
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, vals = raw[0], raw[1], raw[-1]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        tot += float(vals[i].replace(",", "")) * scale
    return tot


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        print(int(dram_bytes(sys.argv[1])))
    else:
        main(sys.argv[1])
