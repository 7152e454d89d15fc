"""Attribute ncu samples / executed instructions of one kernel to source lines, using the line info of the
cubin inside libcassie2d.so.  usage: ncu_source_breakdown.py report.ncu-rep <mangled-kernel-prefix> [bucket]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
import numpy as np

rep, prefix = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cassierl_b200", "lib", "libcassie2d.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
sass = ""
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if ".text." + prefix in out:
            sass = out
            break
txt = sass.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
end = [i for i, l in enumerate(txt) if l.startswith(".text.") and i > start]
end = end[0] if end else len(txt)
cur, lines = None, []
for line in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
ex = np.array([int(r[ia]) if r[ia].isdigit() else 0 for r in data]); sm = np.array([int(r[isamp]) if r[isamp].isdigit() else 0 for r in data])
def col(name):
    i = hdr.index(name)
    return np.array([int(r[i]) if r[i].isdigit() else 0 for r in data])
st_long, st_noi, st_wait = col("stall_long_sb"), col("stall_no_inst"), col("stall_wait")
n = min(len(lines), len(data))
print("SASS lines: cubin %d, report %d%s" % (len(lines), len(data), "" if len(lines) == len(data) else "  (MISMATCH: stale library?)"))
size = collections.Counter(); s_ = collections.Counter(); e_ = collections.Counter()
l_ = collections.Counter(); n_ = collections.Counter(); w_ = collections.Counter()
for i in range(n):
    f, l = lines[i]; k = (f, l // bucket * bucket); size[k] += 1; s_[k] += sm[i]; e_[k] += ex[i]
    l_[k] += st_long[i]; n_[k] += st_noi[i]; w_[k] += st_wait[i]
# the last three columns are shares of ALL samples of the kernel spent in that stall reason inside the region
print("%-32s %8s %9s %8s %8s %8s %8s" % ("source region", "SASS", "samples%", "exec%", "long_sb%", "no_inst%", "wait%"))
for k, v in s_.most_common(int(os.environ.get("TOP", "30"))):
    print("%-24s %6d %8d %8.1f%% %7.1f%% %7.1f%% %7.1f%% %7.1f%%" % (k[0], k[1], size[k], 100 * v / sm.sum(), 100 * e_[k] / ex.sum(),
          100 * l_[k] / sm.sum(), 100 * n_[k] / sm.sum(), 100 * w_[k] / sm.sum()))
print("never executed SASS: %d of %d" % (int((ex[:n] == 0).sum()), n))
