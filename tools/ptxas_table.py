"""Registers / stack / spills of every pf_step_tile instantiation from a `nvcc -Xptxas -v` log."""
import re, subprocess, sys
t = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else "pf_step_tile"
for m in re.finditer(r"Compiling entry function '(\S+)'[^\n]*\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores[^\n]*\n[^\n]*Used (\d+) registers", t):
    name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
    if pat in name:
        print(name.replace('void pz::', '').replace('(pz::TileArgs)', '')[:100], 'stack', m.group(2), 'spill', m.group(3), 'regs', m.group(4))
