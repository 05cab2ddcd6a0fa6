"""Prints SASS size / instruction mix of every kernel in libmincurv_b200.so (code must stay I-cache friendly)."""
import re, subprocess, sys, os
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "global_racetrajectory_optimization_b200", "libmincurv_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0][:48]
    addrs = re.findall(r'/\*([0-9a-f]{4,6})\*/\s+[A-Z@]', f)
    if addrs:
        print(f"{name:50s} {int(addrs[-1],16)/1024:7.1f} KB  DMMA {f.count('DMMA'):4d} DFMA {f.count('DFMA'):5d} LDL {f.count('LDL'):4d} STL {f.count('STL'):4d} MUFU {f.count('MUFU'):3d}")
