"""Static SASS instruction count per source line of one kernel (nvdisasm -gi on the library's cubin; no GPU needed).
With inlining info the count is attributed to the OUTERMOST frame in a chosen file, so a line of the kernel body
carries the instructions of everything it calls.

    python tools/sass_lines.py <library.so> <kernel name substring> [file substring, default ecmc_spec.cuh]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    library, kernel = sys.argv[1], sys.argv[2]
    anchor = sys.argv[3] if len(sys.argv) > 3 else "ecmc_spec.cuh"
    work = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(library)], cwd=work, check=True, capture_output=True)
    cubin = [os.path.join(work, f) for f in os.listdir(work) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
    sections = re.split(r"\n\s*\.section\s+\.text\.", text)
    body = next(p for p in sections[1:] if kernel in p.split(",")[0])
    counts, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    current = None
    pending = []  # file/line annotations of the next instruction: first = innermost, following "inlined at" = outer frames
    for line in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            if pending:
                frames = [f for f in pending if anchor in f[0]]
                current = frames[-1] if frames else pending[-1]
                pending = []
            if current:
                counts[current] += 1
                ops[current][m.group(1).split(".")[0]] += 1
    total = sum(counts.values())
    print(f"{total} instructions")
    for (path, number), count in sorted(counts.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if anchor in path:
            top = ", ".join(f"{k} {v}" for k, v in ops[(path, number)].most_common(4))
            print(f"{number:5d} {count:5d}  {top}")
    other = sum(c for (path, _), c in counts.items() if anchor not in path)
    print(f"other files: {other}")


if __name__ == "__main__":
    main()
