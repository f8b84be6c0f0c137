"""Registers / stack / spill bytes per kernel from the ptxas -v output of the last build (csrc/build.log)."""
import os, re, sys
log = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mcp_raytracer_b200", "csrc", "build.log")
t = open(log).read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers")
import subprocess
for m in pat.finditer(t):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(rt::")[0]
    if len(sys.argv) > 2 and sys.argv[2] not in name: continue
    print(f"{name[:70]:70s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill st/ld {m.group(3):>4s}/{m.group(4):>4s}")
