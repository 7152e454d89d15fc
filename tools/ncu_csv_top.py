"""Top instructions of a kernel by a stall column, with their source line (see ncu_csv_breakdown.py for the inputs).
usage: ncu_csv_top.py source.csv.gz <kernel-substring> <column> [N]"""
import csv, gzip, io, os, re, subprocess, sys, tempfile
src_csv, prefix, colname = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 30
lib = os.environ.get("LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cassierl_b200", "lib", "libcassie2d.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        m = re.search(r"^\.text\.(\S*%s\S*):" % re.escape(prefix), out, re.M)
        if m:
            name = m.group(1); break
txt = out.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + name + ":")][0]
end = [i for i, l in enumerate(txt) if l.startswith(".text.") and i > start]; end = end[0] if end else len(txt)
cur, lines = None, []
for line in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur: lines.append(cur)
rows = list(csv.reader(io.TextIOWrapper((gzip.open if src_csv.endswith(".gz") else open)(src_csv, "rb"))))
hdr, data = rows[1], rows[2:]
ic, ie, isrc, ism = hdr.index(colname), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
order = sorted(range(min(len(data), len(lines))), key=lambda i: -(int(data[i][ic]) if data[i][ic].isdigit() else 0))[:N]
for i in order:
    print("%6s %8s %6s  %-22s %4d  %s" % (data[i][ic], data[i][ie], data[i][ism], lines[i][0], lines[i][1], data[i][isrc].strip()[:70]))
