"""Attribute executed instructions / stall samples of one kernel to source lines, from a source-page CSV that was
exported on the GPU box (`ncu -i x.ncu-rep --page source --csv | gzip`) and the line info of the cubin in the local
libcassie2d.so (the same binary travels to the box).
usage: ncu_csv_breakdown.py source.csv.gz <mangled-kernel-substring> [bucket] [--by-file]"""
import collections, csv, gzip, io, os, re, subprocess, sys, tempfile
import numpy as np

src_csv, prefix = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 10
lib = os.environ.get("LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cassierl_b200", "lib", "libcassie2d.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
sass = None
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        m = re.search(r"^\.text\.(\S*%s\S*):" % re.escape(prefix), out, re.M)
        if m:
            sass, name = out, m.group(1)
            break
assert sass, "kernel not found in " + lib
txt = sass.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + name + ":")][0]
end = [i for i, l in enumerate(txt) if l.startswith(".text.") and i > start]
end = end[0] if end else len(txt)
cur, lines, inl = None, [], []
for line in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        lines.append(cur)
op = gzip.open if src_csv.endswith(".gz") else open
rows = list(csv.reader(io.TextIOWrapper(op(src_csv, "rb"))))
hdr, data = rows[1], rows[2:]
def col(name):
    i = hdr.index(name)
    return np.array([int(r[i]) if r[i].isdigit() else 0 for r in data])
ex, sm, thr = col("Instructions Executed"), col("# Samples"), col("Thread Instructions Executed")
st_long, st_noi, st_wait, st_short = col("stall_long_sb"), col("stall_no_inst"), col("stall_wait"), col("stall_short_sb")
isrc = hdr.index("Source")
n = min(len(lines), len(data))
print("kernel %s\nSASS lines: cubin %d, report %d%s" % (name, len(lines), len(data), "" if len(lines) == len(data) else "  (MISMATCH: stale library?)"))
byfile = "--by-file" in sys.argv
agg = collections.defaultdict(lambda: np.zeros(8))
for i in range(n):
    f, l = lines[i]
    k = (f, 0) if byfile else (f, l // bucket * bucket)
    agg[k] += np.array([1, sm[i], ex[i], st_long[i], st_noi[i], st_wait[i], st_short[i], thr[i]])
tot = np.array([1, sm.sum(), ex.sum(), sm.sum(), sm.sum(), sm.sum(), sm.sum(), 1])
print("total executed warp instructions %d, samples %d, avg active threads %.1f" % (ex.sum(), sm.sum(), thr.sum() / max(ex.sum(), 1)))
print("%-34s %6s %8s %7s %8s %8s %7s %8s %6s" % ("source region", "SASS", "samples%", "exec%", "long_sb%", "no_inst%", "wait%", "short_sb%", "thr"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[: int(os.environ.get("TOP", "40"))]:
    print("%-26s %6d %6d %7.1f%% %6.1f%% %7.1f%% %7.1f%% %6.1f%% %7.1f%% %6.1f" % (k[0], k[1], v[0], 100 * v[1] / tot[1], 100 * v[2] / tot[2],
          100 * v[3] / tot[1], 100 * v[4] / tot[1], 100 * v[5] / tot[1], 100 * v[6] / tot[1], v[7] / max(v[2], 1)))
