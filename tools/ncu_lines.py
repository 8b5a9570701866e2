#!/usr/bin/env python3
"""Per-source-line view of one kernel of an ncu report: the SASS rows of
`ncu -i REP --page source --csv` joined with the line info of `nvdisasm -g` of the same cubin.

    python tools/ncu_lines.py REP.ncu-rep SASS_FILE FUNCTION_SUBSTRING [--top N] [--regions FILE]

SASS_FILE: output of `nvdisasm -g -c <cubin>` (cubins: `cuobjdump -xelf all libiss_cuda.so`).
--regions FILE: lines "label: a-b, c-d, file.h" (line ranges of the main source file, or a file
name for inlined code); prints the table per region instead of per line."""
import argparse
import collections
import csv
import io
import re
import subprocess


def sass_lines(path, func):
    """offset -> (file, line) of the first function whose name contains `func`"""
    out = {}
    active = False
    cur = ("?", 0)
    for ln in open(path):
        if ln.startswith("//----") and ".text." in ln:
            active = func in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("sass")
    ap.add_argument("func")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--regions")
    ap.add_argument("--main", default="sampler.cu")
    ap.add_argument("--kernel", default=None, help="substring of the demangled kernel name in the report")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    # the csv holds one block per kernel: "Kernel Name",... then a header row
    blocks = txt.split('"Kernel Name",')
    want = a.kernel or a.func
    blk = next(b for b in blocks[1:] if want in b.split("\n")[0])
    rows = list(csv.reader(io.StringIO(blk)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    lines = sass_lines(a.sass, a.func)
    agg = collections.defaultdict(lambda: collections.Counter())
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[col["Address"]], 16)
        if base is None:
            base = addr
        key, _ = lines.get(addr - base, (("?", 0), ""))
        c = agg[key]
        c["samples"] += int(r[col["# Samples"]] or 0)
        c["inst"] += int(r[col["Instructions Executed"]] or 0)
        c["tinst"] += int(r[col["Thread Instructions Executed"]] or 0)
        for s in reasons:
            c[s] += int(r[col[s]] or 0)
    for c in agg.values():
        tot.update(c)
    print("total: %d stall samples, %.3e warp instructions, %.1f of 32 lanes active per instruction"
          % (tot["samples"], tot["inst"], tot["tinst"]/max(1, tot["inst"])))

    def show(label, c):
        top = sorted(((c[s], s[6:]) for s in reasons), reverse=True)[:3]
        print("%-58s %6.1f%% %6.1f%% %6.1f  %s" % (
            label, 100.0*c["samples"]/max(1, tot["samples"]), 100.0*c["inst"]/max(1, tot["inst"]),
            c["tinst"]/max(1, c["inst"]),
            ", ".join("%s %d%%" % (n, round(100.0*v/max(1, c["samples"]))) for v, n in top)))

    print("%-58s %7s %7s %6s  %s" % ("where", "samples", "w.instr", "lanes", "top stall reasons"))
    if a.regions:
        regs = []
        for ln in open(a.regions):
            if ":" not in ln:
                continue
            label, spec = ln.rsplit(":", 1)
            items = []
            for it in spec.split(","):
                it = it.strip()
                m = re.match(r"(\d+)-(\d+)$", it)
                if m:
                    items.append((a.main, int(m.group(1)), int(m.group(2))))
                elif it:
                    items.append((it, 0, 10**9))
            regs.append((label.strip(), items))
        racc = collections.OrderedDict((lab, collections.Counter()) for lab, _ in regs)
        racc["other"] = collections.Counter()
        for (f, l), c in agg.items():
            for lab, items in regs:
                if any(f == ff and lo <= l <= hi for ff, lo, hi in items):
                    racc[lab].update(c)
                    break
            else:
                racc["other"].update(c)
        for lab, c in sorted(racc.items(), key=lambda kv: -kv[1]["samples"]):
            show(lab, c)
    else:
        for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:a.top]:
            show("%s:%d" % key, c)


if __name__ == "__main__":
    main()
