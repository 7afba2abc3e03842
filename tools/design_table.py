"""Markdown table of profiles/r02_matrix.jsonl for DESIGN.md section 4."""
import json
import sys

rows = {}
for l in open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_matrix.jsonl"):
    d = json.loads(l)
    c = d["config"]
    rows[(c["robot"], c["op"], d["dtype"])] = d
names = {"iiwa14": "iiwa14", "hyq": "HyQ", "atlas": "Atlas"}
print("| robot | op | FP64 | FP32 |\n|---|---|---|---|")
for r in ("iiwa14", "hyq", "atlas"):
    for op in ("rnea_grad", "minv", "rnea", "crba"):
        cells = []
        for dt in ("f64", "f32"):
            d = rows.get((r, op, dt))
            cells.append("%.2e (%.2f \\| %.2f)" % (d["value"], d["roofline"]["frac"] or 0, d["roofline_hbm"]["frac"]) if d else "-")
        print("| %s | %s | %s | %s |" % (names[r], op if op != "rnea" else "rnea (c)", cells[0], cells[1]))
    cells = []
    for dt in ("f64", "f32"):
        a, b = rows.get((r, "fd", dt)), rows.get((r, "fd_grad", dt))
        cells.append("%.2e / %.2e" % (a["value"], b["value"]) if a and b else "-")
    print("| %s | fd / fd_grad | %s | %s |" % (names[r], cells[0], cells[1]))
