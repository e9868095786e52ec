"""SASS / ptxas evidence for a kernel of the built objects (runs on the CPU box):
   python tools/sass_report.py <object.o> <ptxas.log> <mangled-name-substring> <out.txt> [loop-opcode]
Writes the ptxas -v resource lines of every matching entry point, the opcode histogram of the first match and an excerpt of
its densest stretch of <loop-opcode> instructions (default DMMA)."""
import collections
import re
import subprocess
import sys

obj, log, pat, out = sys.argv[1:5]
loop_op = sys.argv[5] if len(sys.argv) > 5 else "DMMA"
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
funcs, cur = collections.OrderedDict(), None
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
        funcs[cur].append(ln)
names = [f for f in funcs if pat in f]
lines = ["# %s: entry points matching %r" % (obj, pat), ""]
plog = open(log).read().splitlines()
for i, ln in enumerate(plog):
    if "Compiling entry function" in ln and pat in ln:
        lines.append(ln.strip())
        for j in range(i + 1, min(i + 5, len(plog))):
            if "Compiling entry function" in plog[j]:
                break
            lines.append("    " + plog[j].strip())
for nm in names[:1]:
    ins = funcs[nm]
    ops = collections.Counter()
    for ln in ins:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    lines += ["", "## opcode histogram of %s (%d SASS instructions)" % (nm, len(ins))]
    for op, c in ops.most_common(40):
        lines.append("  %-12s %6d" % (op, c))
    marks = [i for i, ln in enumerate(ins) if loop_op in ln]
    if marks:
        best, bi = 0, marks[0]
        for i in marks:   # densest 80-instruction window
            c = sum(1 for j in marks if i <= j < i + 80)
            if c > best:
                best, bi = c, i
        lines += ["", "## densest %s stretch (80 instructions from #%d, %d %s)" % (loop_op, bi, best, loop_op)]
        lines += [ln.rstrip()[:120] for ln in ins[bi:bi + 80]]
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(lines), "lines")
