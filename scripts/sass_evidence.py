#!/usr/bin/env python
"""Static evidence of the shipped library, written under profiles/ (no GPU needed):
  profiles/<tag>_ptxas_table.txt   registers / spills / stack / shared memory of EVERY kernel (nvcc -Xptxas -v rebuild)
  profiles/<tag>_sass_counts.txt   per-kernel counts of the Blackwell tensor / TMA / TMEM / barrier mnemonics
                                   (cuobjdump -sass of video_similarity_search_b200/libslic_b200.so)
usage: python scripts/sass_evidence.py [tag]      (default tag r2)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_similarity_search_b200 import build as b  # noqa: E402

MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "DMMA", "DFMA", "FHADD",
             "FMNMX3", "REDG", "RED.", "ATOMG", "MEMBAR", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    short = []
    for o in out:      # drop the parameter list, keep the template arguments
        m = re.match(r"^(void )?(slic::\w+(<.*?>)?)\((?!bool|int)", o)
        short.append(m.group(2) if m else o)
    return short


def ptxas_table(path):
    rows = []
    for src in b.SOURCES:
        r = subprocess.run([b._nvcc()] + b.NVCC_FLAGS + ["-c", os.path.join(b.CSRC, src), "-o", "/dev/null"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        cur = None
        for ln in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '([^']+)'", ln)
            if m:
                cur = {"src": src, "name": m.group(1)}
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m:
                cur["stack"], cur["sst"], cur["sld"] = map(int, m.groups())
            m = re.search(r"Used (\d+) registers", ln)
            if m:
                cur["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", ln)
                cur["smem"] = int(s.group(1)) if s else 0
    names = demangle([r["name"] for r in rows])
    with open(path, "w") as f:
        f.write("# nvcc %s, flags: %s\n" % (subprocess.run([b._nvcc(), "--version"], capture_output=True, text=True)
                                           .stdout.strip().splitlines()[-1], " ".join(b.NVCC_FLAGS)))
        f.write("# %-18s %5s %6s %7s %7s %8s  kernel\n" % ("source", "regs", "stack", "spillst", "spillld", "smem(st)"))
        for r, nm in zip(rows, names):
            f.write("%-20s %5d %6d %7d %7d %8d  %s\n" % (r["src"], r.get("regs", -1), r.get("stack", 0), r.get("sst", 0),
                                                        r.get("sld", 0), r.get("smem", 0), nm))
        spilled = [nm for r, nm in zip(rows, names) if r.get("sst", 0) or r.get("sld", 0)]
        f.write("# kernels: %d, with spills: %d %s\n" % (len(rows), len(spilled), spilled))
    return rows


def sass_counts(path):
    sass = subprocess.run(["cuobjdump", "-sass", b.LIB_PATH], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        cur["_total"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                cur[mn] += 1
                if mn in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM"):
                    cur[op] += 1
    names = demangle(list(per.keys()))
    total = collections.Counter()
    with open(path, "w") as f:
        f.write("# cuobjdump -sass %s : instruction counts per kernel (static)\n" % os.path.relpath(b.LIB_PATH, ROOT))
        for (k, c), nm in zip(per.items(), names):
            total.update(c)
            tc = {m: c[m] for m in c if m != "_total" and c[m]}
            if any(m.startswith(("UTC", "UTMA", "LDTM")) for m in tc):
                f.write("%s\n    instructions %d; %s\n" % (nm, c["_total"], ", ".join("%s %d" % kv for kv in sorted(tc.items()))))
        f.write("# whole library: %d kernels, %d instructions\n" % (len(per), total["_total"]))
        for mn in MNEMONICS:
            f.write("#   %-8s %d\n" % (mn, total[mn]))
        f.write("# (HMMA = legacy mma.sync tensor instructions: must be 0; UTCHMMA = tcgen05.mma kind::f16; UTMALDG = TMA tensor\n"
                "#  load; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; SYNCS = mbarrier; STL/LDL = local-memory spills)\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    b.build()
    prof = os.path.join(ROOT, "profiles")
    ptxas_table(os.path.join(prof, "%s_ptxas_table.txt" % tag))
    sass_counts(os.path.join(prof, "%s_sass_counts.txt" % tag))
    print("written", tag)


if __name__ == "__main__":
    main()
