"""Static SASS opcode histogram of one kernel of libecmc_b200.so (no GPU needed): a cheap first look at what an edit did
to the instruction mix before GPU time is spent on it (the dynamic count per event comes from ncu, profiles/README.md).

    python tools/sass_histogram.py [library.so] [substring of the mangled kernel name]

Default kernel: the bench's event_kernel<LJ, none, LJ, single occupant, no records> (C2)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = (("fp64", r"^D(ADD|MUL|FMA|SETP|MNMX)|^MUFU\.(RCP64H|RSQ64H)"),
           ("control flow", r"^(BRA|BSSY|BSYNC|BREAK|CALL|RET|EXIT|WARPSYNC|BMOV|JMP|NANOSLEEP)"),
           ("moves / selects", r"^(MOV|IMAD\.MOV|SEL|FSEL|PRMT|SHFL|UMOV|R2UR|S2R|S2UR|CS2R|P2R|R2P)"),
           ("integer", r"^(IMAD|IADD3|LOP3|SHF|LEA|ISETP|POPC|FLO|VOTE|REDUX|PLOP3|I2F|F2I|I2I|IABS|UIADD3|ULOP3|USHF|UISETP|ULEA|UIMAD|VIADD|VIMNMX)"),
           ("memory", r"^(LD|ST|ATOM|RED|LDG|STG|LDS|STS|LDL|STL|LDC|ULDC|MEMBAR|CCTL)"))


def kernels(library):
    out = subprocess.run(["cuobjdump", "-sass", library], capture_output=True, text=True, check=True).stdout
    current, table = None, collections.OrderedDict()
    for line in out.splitlines():
        match = re.search(r"Function : (\S+)", line)
        if match:
            current = table.setdefault(match.group(1), [])
            continue
        match = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if match and current is not None:
            current.append(match.group(1))
    return table


def main():
    library = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jellyfysh_b200", "libecmc_b200.so")
    pattern = sys.argv[2] if len(sys.argv) > 2 else "event_kernelILi2ELi0ELi2ELb1ELb0E"
    for name, opcodes in kernels(library).items():
        if pattern not in name:
            continue
        print(f"{name}: {len(opcodes)} instructions (static)")
        by_class = collections.Counter()
        for opcode in opcodes:
            for label, regex in CLASSES:
                if re.match(regex, opcode):
                    by_class[label] += 1
                    break
            else:
                by_class["other"] += 1
        print("  " + ", ".join(f"{label} {count}" for label, count in by_class.most_common()))
        top = collections.Counter(op.split(".")[0] if not op.startswith("IMAD.MOV") else "IMAD.MOV" for op in opcodes)
        print("  " + ", ".join(f"{op} {count}" for op, count in top.most_common(16)))


if __name__ == "__main__":
    main()
