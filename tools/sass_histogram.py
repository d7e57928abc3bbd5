#!/usr/bin/env python
"""Static SASS evidence for the hot kernels (runs here: cuobjdump needs no GPU).  For every kernel whose mangled name matches
the regex: instruction count, opcode histogram of the whole function and of its largest loop body (backward branch span).
usage: python tools/sass_histogram.py <object.o> <regex> [more regexes ...]"""
import collections
import re
import subprocess
import sys


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2)))
    if name:
        yield name, body


def opcode(txt):
    txt = re.sub(r"^@!?U?P\d+\s+", "", txt)
    return txt.split()[0].split(".")[0]


def hist(ins):
    c = collections.Counter(opcode(t) for _, t in ins)
    return " ".join(f"{k}:{v}" for k, v in c.most_common())


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:140]


obj = sys.argv[1]
for name, ins in functions(obj):
    if not any(re.search(rx, name) for rx in sys.argv[2:]):
        continue
    loops = []
    for addr, txt in ins:
        if "BRA" in txt:
            t = re.search(r"0x([0-9a-f]+)", txt)
            if t and int(t.group(1), 16) < addr:
                loops.append((int(t.group(1), 16), addr))
    print(f"== {demangle(name)}")
    print(f"   {len(ins)} instructions; FP64 (DADD/DMUL/DFMA/DSETP/MUFU.64): {sum(1 for _, t in ins if opcode(t) in ('DADD', 'DMUL', 'DFMA', 'DSETP'))}")
    print(f"   all: {hist(ins)}")
    if loops:
        a, b = max(loops, key=lambda x: x[1] - x[0])
        body = [(ad, t) for ad, t in ins if a <= ad <= b]
        fp = sum(1 for _, t in body if opcode(t) in ("DADD", "DMUL", "DFMA"))
        print(f"   largest loop body 0x{a:x}-0x{b:x}: {len(body)} instructions, {fp} DADD/DMUL/DFMA")
        print(f"   loop: {hist(body)}")
