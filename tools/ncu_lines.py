#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep (ncu --set full --import-source on): warp instructions
executed and stall samples per CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_pct]
"""
import csv, io, subprocess, sys, os
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; hdr_list = r; continue
    if hdr is None: continue
    if r[0] != "":   # source line summary row
        def g(h):
            try:
                return int((r[hdr[h]] or "0").split("(")[0])
            except (ValueError, IndexError, KeyError):      # optional columns differ between kernels
                return 0
        lines.append(dict(file=cur_file, line=int(r[0]), src=r[1].strip(), smp=g("# Samples"), inst=g("Instructions Executed"),
                          wait=g("stall_wait"), short=g("stall_short_sb"), long=g("stall_long_sb"), mio=g("stall_mio"),
                          math=g("stall_math"), br=g("stall_branch_resolving"), noinst=g("stall_no_inst"), bar=g("stall_barrier"),
                          shx=g("L1 Wavefronts Shared Excessive"), sh=g("L1 Wavefronts Shared")))
tots = sum(l["smp"] for l in lines); toti = sum(l["inst"] for l in lines)
print("total samples %d, warp instructions %d" % (tots, toti))
for l in lines:
    if l["smp"] >= min_pct / 100 * tots or l["inst"] >= min_pct / 100 * toti:
        print("%-28s %4d inst %5.1f%% smp %5.1f%% (wait %4.1f sh %4.1f lg %4.1f mio %4.1f math %4.1f br %4.1f) shwf %9d x%9d | %s" % (
            l["file"][:28], l["line"], 100 * l["inst"] / toti, 100 * l["smp"] / tots, 100 * l["wait"] / tots, 100 * l["short"] / tots,
            100 * l["long"] / tots, 100 * l["mio"] / tots, 100 * l["math"] / tots, 100 * l["br"] / tots, l["sh"], l["shx"], l["src"][:90]))
