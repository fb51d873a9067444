"""Print registers / spills per kernel from build/*.ptxas.log (nvcc -Xptxas -v)."""
import glob, os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for log in sorted(glob.glob(os.path.join(here, "build", "*.ptxas.log"))):
    txt = open(log).read()
    blocks = re.split(r"Compiling entry function '", txt)[1:]
    for b in blocks:
        name = b.split("'")[0]
        regs = re.search(r"Used (\d+) registers", b)
        spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", b)
        smem = re.search(r"(\d+) bytes smem", b)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"cymf::|\((?:cymf::)?\w+Args<\w+>(?:, int)?\)|void ", "", d)[:90]
        print(f"{os.path.basename(log)[:-10]:6s} {d:92s} regs={regs.group(1):>3s} stack={spill.group(1):>4s} spill={spill.group(2):>4s} smem={smem.group(1) if smem else 0}")
