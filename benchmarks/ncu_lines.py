"""Attribute ncu warp-stall samples (source page, SASS view) to CUDA source lines using nvdisasm line info.
  python benchmarks/ncu_lines.py <report.ncu-rep> <object-or-cubin> <kernel-substring> [top]
Needed because inlined device code in headers does not show in ncu's CUDA view without the source tree."""
import csv, re, subprocess, sys, collections, tempfile, os, glob
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
td = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    obj = glob.glob(os.path.join(td, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = {}, None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = kern in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: i for i, n in enumerate(hdr)}
base = None
agg = collections.Counter(); inst = collections.Counter(); reasons = collections.defaultdict(collections.Counter)
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[0], 16)
    if base is None: base = a
    key = lines.get(a - base, (None, ""))[0]
    s = int(r[ci["# Samples"]] or 0)
    agg[key] += s; tot += s
    inst[key] += int(r[ci["Instructions Executed"]] or 0)
    for n in stall_cols:
        v = int(r[ci[n]] or 0)
        if v: reasons[key][n[6:]] += v
print(f"total samples {tot}, instructions {sum(inst.values())}")
for key, s in agg.most_common(top):
    rs = ", ".join(f"{k}:{v}" for k, v in reasons[key].most_common(3))
    print(f"{100.0*s/tot:5.1f}%  inst {100.0*inst[key]/max(1,sum(inst.values())):5.1f}%  {key}  [{rs}]")
