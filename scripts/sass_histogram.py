"""SASS opcode histogram per kernel of libpwc_b200.so (cuobjdump -sass): the checkable form of the
"Blackwell-native" claim (UTMALDG = TMA tensor loads, UTMAPF = TMA prefetch, SYNCS = mbarrier, UCGABAR = cluster
barrier, PREEXIT = griddepcontrol.launch_dependents, STG.E.ENL2.256 = 256-bit stores, REDG = vector reductions,
LDGSTS = cp.async; no UTC*MMA / LDTM / STTM: the path uses no tensor cores).
usage: python scripts/sass_histogram.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pwc_net_pytorch_b200", "lib", "libpwc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
INTERESTING = ("UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UBLKCP", "SYNCS", "UCGABAR", "PREEXIT", "ACQBULK", "LDGSTS",
               "ARRIVES", "REDG", "RED.", "ATOMS", "ATOMG", "STG.E.ENL2.256", "LDS.128", "LDS.64", "LDS ", "STS", "LDG", "STG",
               "FFMA2", "FFMA", "FMUL", "IMAD", "SHFL", "BAR.", "UTC", "LDTM", "STTM", "HMMA", "LDL", "STL", "MUFU", "CCTL")
kern = None
hist = collections.OrderedDict()
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)[:110]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(2)
        hist[kern]["total"] += 1
        for key in INTERESTING:
            if op.startswith(key.strip()) if not key.endswith(" ") else op == key.strip():
                hist[kern][key.strip()] += 1
                break
print(f"# SASS opcode histogram per kernel: cuobjdump -sass {os.path.relpath(lib, ROOT)} (static instruction counts)")
tot = collections.Counter()
for k, h in hist.items():
    tot.update(h)
    print(f"\n== {k}")
    print("   " + "  ".join(f"{op}={n}" for op, n in sorted(h.items(), key=lambda kv: -kv[1])))
print("\n== whole library")
print("   " + "  ".join(f"{op}={n}" for op, n in sorted(tot.items(), key=lambda kv: -kv[1])))
for key in ("UTC", "LDTM", "STTM", "HMMA"):
    print(f"   {key}* instructions: {tot.get(key, 0)}")
