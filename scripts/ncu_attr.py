"""Attribute an ncu source-page export (SASS level) of one kernel to source functions.

  cuobjdump -xelf all libmaple_b200.so; nvdisasm -g -c *.cubin > all.txt
  ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
  python scripts/ncu_attr.py all.txt sass.csv '<mangled kernel prefix>' [top [csrc dir of the profiled build]]

Prints, per source function (found by line ranges in the files under maple_b200/csrc), the warp instructions executed, the
average active threads, the stall samples and the share of 'no instruction' samples; then the hottest source lines."""
import collections, csv, re, sys

dis, sass, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
csrc = sys.argv[5] if len(sys.argv) > 5 else None  # the sources the profiled library was built from, if not the current ones
# offset -> (file, line)
loc = {}
cur = None
inside = False
for ln in open(dis, errors="replace"):
    if ln.startswith("\t.section\t.text."):
        inside = ln.startswith("\t.section\t.text." + prefix)
        cur = None
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        loc[int(m.group(1), 16)] = cur
# function line ranges
funcs = {}
import glob, os
for f in glob.glob(os.path.join(csrc or os.path.join(os.path.dirname(__file__), "..", "maple_b200", "csrc"), "*")):
    starts = []
    for i, ln in enumerate(open(f, errors="replace"), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__|inline|__forceinline__|__noinline__|\s)+[\w:<>\*&\s]+?\b(\w+)\s*\(", ln)
        if m and not ln.startswith(" ") and m.group(1) not in ("if", "for", "while", "switch", "return"):
            starts.append((i, m.group(1)))
    funcs[os.path.basename(f)] = starts

def func_of(file, line):
    best = "?"
    for s, name in funcs.get(file, []):
        if s <= line:
            best = name
        else:
            break
    return best

rows = list(csv.reader(open(sass)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: collections.Counter())
lines = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    L = loc.get(a - base)
    file, line = L if L else ("?", 0)
    fn = file + ":" + func_of(file, line)
    d = {"inst": int(r[col["Instructions Executed"]]), "thr": int(r[col["Thread Instructions Executed"]]), "samples": int(r[col["# Samples"]])}
    for h in stall_cols:
        d[h] = int(r[col[h]] or 0)
    for k, v in d.items():
        agg[fn][k] += v
        lines[(file, line)][k] += v
        tot[k] += v
print("total warp instructions %.4g, avg threads %.1f, samples %d" % (tot["inst"], tot["thr"] / max(1, tot["inst"]), tot["samples"]))
print("stalls overall: " + "  ".join("%s %.1f%%" % (h[6:], 100.0 * tot[h] / max(1, tot["samples"])) for h in sorted(stall_cols, key=lambda h: -tot[h])[:8]))
print("%-44s %7s %7s %6s %7s | top stalls" % ("function", "inst%", "thr/in", "smp%", "noinst%"))
for fn, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(stall_cols, key=lambda h: -d[h])[:3]
    print("%-44s %6.2f%% %7.1f %5.1f%% %6.1f%% | %s" % (fn[:44], 100.0 * d["inst"] / tot["inst"], d["thr"] / max(1, d["inst"]), 100.0 * d["samples"] / tot["samples"],
                                                     100.0 * d["stall_no_inst"] / max(1, d["samples"]), " ".join("%s %.0f%%" % (h[6:], 100.0 * d[h] / max(1, d["samples"])) for h in st)))
print("hottest lines:")
for (file, line), d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print("  %-22s %5d %-26s inst %5.2f%% thr %4.1f smp %5.2f%%" % (file, line, func_of(file, line)[:26], 100.0 * d["inst"] / tot["inst"], d["thr"] / max(1, d["inst"]), 100.0 * d["samples"] / tot["samples"]))
