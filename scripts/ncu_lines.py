"""Summarise an ncu source page: samples and stall reasons per CUDA source line.
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [file-substring] [top]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path, newline="")))
kernels = []; cur = None; fname = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1]; continue
    if r[0] == "Function Name":
        cur = next((k for k in kernels if k["name"] == r[1][:80] and k["file"] == fname), None)
        if cur is None or fname in cur["seen"]:
            cur = {"name": r[1][:80], "file": fname, "seen": set(), "lines": collections.defaultdict(lambda: collections.Counter()), "src": {}}; kernels.append(cur)
        cur["seen"].add(fname); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur is not None and r[0].isdigit() and fname and want in fname:
        d = dict(zip(hdr, r))
        ln = int(r[0]); cur["src"][ln] = r[1].strip()[:90]
        c = cur["lines"][ln]
        # columns after the 4th are metrics of the line (aggregated over its SASS)
        for k, v in zip(hdr[4:], r[4:]):
            if k.startswith("stall_") and "Not Issued" not in k or k in ("# Samples", "Instructions Executed"):
                try: c[k] += float(v)
                except ValueError: pass
for k in kernels:
    tot = sum(c["# Samples"] for c in k["lines"].values())
    if not tot: continue
    print("==", k["name"], "total samples", int(tot))
    for ln, c in sorted(k["lines"].items(), key=lambda x: -x[1]["# Samples"])[:top]:
        st = sorted(((v, n) for n, v in c.items() if n.startswith("stall_")), reverse=True)[:3]
        print(f"{ln:5d} {int(c['# Samples']):6d} {100*c['# Samples']/tot:5.1f}%  inst {int(c['Instructions Executed']):8d}  " + " ".join(f"{n[6:]}={int(v)}" for v, n in st) + "  | " + k["src"][ln])
