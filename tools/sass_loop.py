"""Static look at one kernel's SASS: opcode histogram of the whole function and of its hottest loop
(the backward branch spanning the most instructions).  Usage: python tools/sass_loop.py <obj> <substring of the mangled name>"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = next((f for f in funcs[1:] if pat in f.split("\n", 1)[0]), None)
    if body is None:
        sys.exit(f"no function matching {pat}")
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(body.split("\n", 1)[0].strip(), "-", len(ins), "instructions")
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
    loops.sort(reverse=True)
    for span, lo, hi in loops[:3]:
        ops = collections.Counter()
        for addr, text in ins:
            if lo <= addr <= hi:
                t = re.sub(r"^@!?U?P\d+\s+", "", text)
                ops[t.split()[0].split(".")[0]] += 1
        n = sum(ops.values())
        print(f"loop 0x{lo:x}..0x{hi:x}: {n} instructions:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
