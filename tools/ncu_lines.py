#!/usr/bin/env python
"""Per-CUDA-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [--top 25] [--kernel-index 0]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Function Name":
            cur["fn"] = r[1]
        elif r and r[0] == "Line No":
            cur["hdr"] = r
        elif cur is not None and r and r[0].isdigit():
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        key = (b["file"], b.get("fn"))
        if key in seen or not b["rows"]:
            continue
        seen.add(key)
        h = b["hdr"]
        si = h.index("# Samples")
        stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        tot = sum(float(r[si]) for r in b["rows"] if r[si].replace(".", "").isdigit()) or 1.0
        print(f"=== {b['file'].split('/')[-1]} :: {b.get('fn', '')[:70]}  ({int(tot)} samples)")
        ok = [r for r in b["rows"] if r[si].replace(".", "").isdigit()]
        body = sorted(ok, key=lambda r: -float(r[si]))[:top]
        for r in body:
            n = float(r[si])
            st = sorted(((float(r[i]) if r[i].replace(".", "").isdigit() else 0.0, h[i][6:]) for i in stall_cols), reverse=True)[:3]
            why = ", ".join(f"{nm} {v / max(n, 1) * 100:.0f}%" for v, nm in st if v > 0)
            print(f"  {n / tot * 100:5.1f}%  L{r[0]:>4s}  {r[1].strip()[:80]:80s} [{why}]")

if __name__ == "__main__":
    main()
