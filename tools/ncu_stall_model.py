"""Per source region: cycles predicted by the SASS control codes (sum over executed instructions of the
compiler-assigned stall count = what a LONE warp spends if nothing but fixed-latency dependencies holds it up)
next to the cycles ncu's sampler saw there.  Regions where sampled >> predicted wait on variable-latency events
(instruction fetch, local memory); regions where they agree are dependency-chain bound.
usage: ncu_stall_model.py report.ncu-rep <mangled-kernel-prefix> <cubin-name-fragment> [bucket]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, prefix, cub = sys.argv[1], sys.argv[2], sys.argv[3]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 20
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cassierl_b200", "lib", "libcassie2d.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin") and cub in f][0]
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
end = [i for i, l in enumerate(txt) if l.startswith(".text.") and i > start]
end = end[0] if end else len(txt)
kname = txt[start][6:-1]
cur, lines = None, []
for line in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        lines.append(cur)
hexs = subprocess.run(["cuobjdump", "-sass", "-fun", kname, cubin], capture_output=True, text=True).stdout.split("\n")
stall = []
for i, l in enumerate(hexs):
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+.*;\s+/\* 0x[0-9a-f]{16} \*/", l) and i + 1 < len(hexs):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", hexs[i + 1])
        if m2:
            stall.append((int(m2.group(1), 16) >> 41) & 0xf)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
d = dict(zip(raw[0], raw[-1]))
warps = float(d["launch__grid_size"]) * float(d["launch__block_size"]) / 32
cyc = float(d["sm__cycles_elapsed.max"].replace(",", ""))
n = min(len(lines), len(data), len(stall))
print("SASS: lineinfo %d, report %d, control codes %d; %d warps, %.0f cycles elapsed" % (len(lines), len(data), len(stall), warps, cyc))
pred = collections.Counter(); samp = collections.Counter(); ex = collections.Counter()
for i in range(n):
    e = int(data[i][ia]) if data[i][ia].isdigit() else 0
    s = int(data[i][isamp]) if data[i][isamp].isdigit() else 0
    k = (lines[i][0], lines[i][1] // bucket * bucket)
    pred[k] += e * max(stall[i], 1) / warps; samp[k] += s; ex[k] += e / warps
ts = sum(samp.values())
print("%-30s %10s %12s %12s %7s" % ("source region", "instr/warp", "pred cycles", "sampled cyc", "ratio"))
for k, v in sorted(samp.items(), key=lambda kv: -kv[1])[:int(os.environ.get("TOP", "30"))]:
    sc = v / ts * cyc
    print("%-24s %5d %10.0f %12.0f %12.0f %7.2f" % (k[0], k[1], ex[k], pred[k], sc, sc / max(pred[k], 1)))
print("total: instr/warp %.0f, predicted %.0f cycles, elapsed %.0f" % (sum(ex.values()), sum(pred.values()), cyc))
