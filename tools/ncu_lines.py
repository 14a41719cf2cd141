"""Per-source-line instruction / stall-sample shares from an ncu report:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_lines.py src.csv '<kernel substr>' [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sel = sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur_file = cur_fn = None; out = []; hdr = None
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]
    elif r[0] == "Function Name": cur_fn = r[1]
    elif r[0] == "Line No": hdr = r
    elif r[0].isdigit() and sel in (cur_fn or ""):
        iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
        out.append((num(r[iE]), num(r[iS]), cur_file, int(r[0]), r[1][:110]))
tot = sum(o[0] for o in out) or 1; tots = sum(o[1] for o in out) or 1
print("total warp-instructions", tot, "samples", tots)
for e, s, f, l, src in sorted(out, reverse=True)[:topn]:
    print(f"{100*e/tot:5.1f}% ex {100*s/tots:5.1f}% smp  {f}:{l}  {src}")
