"""Annotate cuobjdump -sass output of one kernel with the scheduling control bits (sm_100a, 128-bit encoding):
stall count, yield, write / read scoreboard set by the instruction, and the mask of scoreboards it waits for.
Lets a scoreboard wait that covers an unrelated in-flight load be found without a GPU.

usage: python scripts/sass_ctrl.py <object or .so> <kernel-name-substring> [first [last]]
"""
import re, subprocess, sys

def kernel_lines(obj, needle):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
    start = None
    for i, l in enumerate(out):
        if "Function :" in l:
            if start is not None:
                return out[start:i]
            if needle in l:
                start = i
    return out[start:] if start is not None else []

def decode(lines):
    ins = []
    pat = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
    cur = None
    for l in lines:
        m = pat.match(l)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]
            continue
        m = pat2.match(l)
        if m and cur is not None:
            hi = int(m.group(1), 16)
            ctrl = hi >> 41
            ins.append(dict(addr=cur[0], text=cur[1], stall=ctrl & 15, yield_=(ctrl >> 4) & 1, wr=(ctrl >> 5) & 7,
                            rd=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 63))
            cur = None
    return ins

if __name__ == "__main__":
    ins = decode(kernel_lines(sys.argv[1], sys.argv[2]))
    a = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    b = int(sys.argv[4]) if len(sys.argv) > 4 else len(ins)
    for i, d in enumerate(ins[a:b], a):
        wr = "-" if d["wr"] == 7 else str(d["wr"])
        rd = "-" if d["rd"] == 7 else str(d["rd"])
        wt = "".join(str(k) for k in range(6) if d["wait"] >> k & 1) or "-"
        print(f'{i:5d} s{d["stall"]:<2d} w{wr} r{rd} wait[{wt:<6s}] {d["text"]}')
