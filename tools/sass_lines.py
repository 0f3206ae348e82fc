#!/usr/bin/env python
"""dev helper: static SASS instruction count per source line (and per opcode class) of one kernel.
Usage: python tools/sass_lines.py <kernel-name-substring> [lib.so]"""
import collections, os, re, subprocess, sys, tempfile
kern = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else "tiktoken_b200/csrc/libb200bpe.so"
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
ALU = ("LOP3", "SHF", "IADD3", "VIADD", "ISETP", "PRMT", "LEA", "SEL", "POPC", "FLO", "BREV", "IMNMX", "VIMNMX", "MOV", "PLOP3", "IABS", "LOP")
FMA = ("IMAD", "FFMA", "FMUL", "HFMA2")
cur = None; line = None; by_line = collections.Counter(); by_pipe = collections.Counter(); by_line_alu = collections.Counter()
for l in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m: cur = m.group(1); continue
    if cur is None or kern not in cur: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        op = m.group(1).split(".")[0]
        pipe = "fma" if op in FMA else "alu" if op in ALU else "other"
        by_line[line] += 1; by_pipe[pipe] += 1
        if pipe == "alu": by_line_alu[line] += 1
print(dict(by_pipe))
for (ln, c) in by_line.most_common(int(os.environ.get("TOP", "40"))):
    print(f"{c:5d} (alu {by_line_alu[ln]:4d})  {ln[0]}:{ln[1]}")
