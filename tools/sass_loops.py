#!/usr/bin/env python3
"""Static instruction mix of the loops of one kernel, from `cuobjdump -sass`.

Usage: tools/sass_loops.py <object or .so> <kernel-name-substring> [--min N]

For every backward branch (a loop) the body is the address range [target, branch]; the script prints the number
of instructions per issue pipe (ALU / FMA / LSU / other, classification from tools/alu_peak.cu's measurements on
B200: VIADD.16x2, IMAD* on the FMA pipe; VIADDMNMX, VIMNMX*, LOP3, PRMT, SHF, LEA, IADD3, ISETP, SEL on the ALU pipe)
and the most frequent opcodes.  With no GPU in the development container this is the first check of a kernel change:
the turbo decoder is bound by ALU-pipe issue, so ALU instructions per trellis step is the number to watch."""
import collections
import re
import subprocess
import sys

FMA = ("VIADD.16x2", "IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "FFMA2", "FADD2", "FMUL2")
ALU = ("VIADDMNMX", "VIMNMX", "LOP3", "PRMT", "SHF", "LEA", "IADD3", "ISETP", "SEL", "VIADD", "IABS", "FMNMX", "FSETP", "MOV",
       "POPC", "FLO", "BREV", "SGXT", "BMSK", "I2I", "F2I", "I2F", "VABSDIFF", "IMNMX", "PLOP3", "CS2R", "FSEL", "VOTE", "P2R", "R2P")
LSU = ("LDS", "STS", "LDG", "STG", "LD.", "ST.", "LDL", "STL", "ATOMS", "ATOMG", "ATOM", "RED", "LDSM", "LDGSTS", "CCTL", "LDC", "SHFL",
       "MEMBAR", "ERRBAR", "LDGDEPBAR", "DEPBAR")


def pipe(op):
    if op.startswith("VIADD.16x2") or op.startswith("IMAD"):
        return "fma"
    for p in LSU:
        if op.startswith(p):
            return "lsu"
    for p in FMA:
        if op.startswith(p):
            return "fma"
    for p in ALU:
        if op.startswith(p):
            return "alu"
    return "other"


def main():
    path, name = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 16
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    ins, on = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            on = name in line
            if on:
                print("==", line.split(":")[1].strip())
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            text = m.group(2).strip()
            pred = ""
            mm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
            if mm:
                pred, text = mm.group(1), mm.group(2)
            ins.append((int(m.group(1), 16), text.split()[0], text, pred))
    addr_idx = {a: i for i, (a, _, _, _) in enumerate(ins)}
    print("instructions:", len(ins))
    loops = []
    for i, (a, op, text, _) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) in addr_idx and int(m.group(1), 16) <= a:
                loops.append((addr_idx[int(m.group(1), 16)], i))
    for lo, hi in sorted(loops):
        n = hi - lo + 1
        if n < min_len:
            continue
        # skip bodies that merely enclose a longer inner loop listing: still printed, flagged
        inner = [(l, h) for (l, h) in loops if l >= lo and h <= hi and (l, h) != (lo, hi) and h - l + 1 >= min_len]
        c = collections.Counter(pipe(op) for _, op, _, _ in ins[lo:hi + 1])
        ops = collections.Counter(op for _, op, _, _ in ins[lo:hi + 1])
        print(f"loop {ins[lo][0]:#07x}..{ins[hi][0]:#07x}  n={n:5d}  alu={c['alu']:5d} fma={c['fma']:5d} lsu={c['lsu']:4d} other={c['other']:4d}"
              f"  encloses {len(inner)} loop(s)")
        print("     ", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
