"""Dynamic instruction counts and stall samples per SOURCE LINE of one kernel: joins the SASS page of an ncu report
(per-instruction executed counts, in address order) with the line table of the library the report was taken from
(nvdisasm -gi on its cubin; instructions attributed to the outermost frame inside the anchor file).

    python tools/ncu_lines.py <report.ncu-rep> <library.so> <kernel name substring> [events per launch] [anchor file]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(library, kernel, anchor):
    work = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(library)], cwd=work, check=True, capture_output=True)
    cubin = [os.path.join(work, f) for f in os.listdir(work) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
    sections = re.split(r"\n\s*\.section\s+\.text\.", text)
    body = next(p for p in sections[1:] if kernel in p.split(",")[0])
    out, pending, current = [], [], None
    for line in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            if pending:
                frames = [f for f in pending if anchor in f[0]]
                current = frames[-1] if frames else (os.path.basename(pending[-1][0]), pending[-1][1])
                pending = []
            out.append((m.group(1), current))
    return out


def main():
    report, library, kernel = sys.argv[1:4]
    events = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    anchor = sys.argv[5] if len(sys.argv) > 5 else "ecmc_spec.cuh"
    text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header = rows[1]
    col = {name: header.index(name) for name in ("Source", "Instructions Executed", "# Samples", "stall_long_sb", "stall_no_inst",
                                                 "stall_wait", "stall_short_sb", "stall_math", "stall_branch_resolving")}
    sass = [r for r in rows[2:] if len(r) > col["stall_wait"]]
    lines = sass_lines(library, kernel, anchor)
    if len(lines) != len(sass):
        print(f"warning: {len(sass)} instructions in the report, {len(lines)} in the library", file=sys.stderr)
    per_line = collections.defaultdict(lambda: collections.Counter())
    for row, (opcode, where) in zip(sass, lines):
        reported = row[col["Source"]].split()
        reported = [t for t in reported if not t.startswith("@")][0].rstrip(";")
        if reported.split(".")[0] != opcode.split(".")[0]:
            print(f"warning: opcode mismatch {reported} vs {opcode}", file=sys.stderr)
            break
        where = where if isinstance(where, tuple) else ("?", 0)
        key = where[1] if anchor in where[0] else f"{where[0]}:{where[1]}"
        c = per_line[key]
        c["executed"] += int(row[col["Instructions Executed"]])
        c["samples"] += int(row[col["# Samples"]])
        for name in ("stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math", "stall_branch_resolving"):
            c[name] += int(row[col[name]])
    total = sum(c["executed"] for c in per_line.values())
    samples = sum(c["samples"] for c in per_line.values())
    print(f"{total / events:.1f} warp instructions per event, {samples} samples")
    print(" line   inst/event  samples%   long_sb no_inst wait short_sb math branch")
    for key in sorted(per_line, key=lambda k: (isinstance(k, str), k)):
        c = per_line[key]
        if c["executed"] / events < 0.5 and c["samples"] < 0.003 * samples:
            continue
        print(f"{str(key):>12} {c['executed'] / events:8.1f} {100.0 * c['samples'] / samples:8.1f}   "
              + " ".join(f"{100.0 * c[n] / samples:6.1f}" for n in ("stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb",
                                                                     "stall_math", "stall_branch_resolving")))


if __name__ == "__main__":
    main()
