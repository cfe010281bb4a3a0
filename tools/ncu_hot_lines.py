"""Summarise an .ncu-rep by CUDA source line: warp-stall samples per line for each kernel.

usage: python tools/ncu_hot_lines.py report.ncu-rep [top_n]
(wraps `ncu -i report --page source --csv --print-source cuda,sass`; needs -lineinfo at compile time)
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    secs, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            cur = dict(name=r[1], hdr=None, lines=[])
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None and r and r[0] == "Line No":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and r and r[0].isdigit():
            cur["lines"].append(r)
    seen = set()
    for sec in secs:
        if sec["name"] in seen or not sec["hdr"]:
            continue
        seen.add(sec["name"])
        h = sec["hdr"]
        si = h.index("# Samples")
        ii = h.index("Instructions Executed")
        data = [(int(r[0]), int(r[si]) if r[si].isdigit() else 0, int(r[ii]) if r[ii].isdigit() else 0, r[1]) for r in sec["lines"]]
        tot = sum(d[1] for d in data) or 1
        print(f"==== {sec['name']}  total samples {tot}")
        for ln, smp, ins, src in sorted(data, key=lambda d: -d[1])[:top]:
            print(f"{ln:5d} {100 * smp / tot:5.1f}%  inst {ins:9d}  {src.strip()[:110]}")


if __name__ == "__main__":
    main()
