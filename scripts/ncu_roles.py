"""Per-role view of a forward-kernel ncu capture (source page): where each warp role's samples go.
usage: python scripts/ncu_roles.py file.ncu-rep [kernel-regex]
Roles are told apart by the source lines of warpcorr_fwd_tma.cuh (function names in the CUDA view)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "warpcorr_fwd_tma"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src = open("pwc_net_pytorch_b200/csrc/warpcorr_fwd_tma.cuh").read().split("\n")
# line ranges of the role functions in the current source
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"void (fwd_role_\w+|global_tap_values)\(", l) or re.match(r"(warpcorr_fwd_tma_kernel)\(", l)
    if m:
        marks.append((i, m.group(1)))
marks.append((len(src) + 1, "end"))


def role_of(line):
    name = "pre"
    for (a, n) in marks:
        if line >= a:
            name = n
    return {"fwd_role_tma": "T", "fwd_role_taps": "P", "fwd_role_bilinear": "B", "warpcorr_fwd_tma_kernel": "C",
            "global_tap_values": "B(cold)"}.get(name, name)


cur = None
hdr = None
agg = collections.defaultdict(collections.Counter)
lines = collections.defaultdict(collections.Counter)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[2] != "-":
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    if n == 0:
        continue
    role = role_of(int(r[0])) if cur == "warpcorr_fwd_tma.cuh" else "inl:" + cur
    agg[role]["n"] += n
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            agg[role][h[6:]] += int(r[ix[h]])
    lines[role][(int(r[0]), r[1].strip()[:80])] += n
tot = sum(v["n"] for v in agg.values())
print("total samples", tot)
for role, v in sorted(agg.items(), key=lambda kv: -kv[1]["n"]):
    top = ", ".join(f"{s}:{c}" for s, c in v.most_common(8) if s != "n")
    print(f"{role:28s} {v['n']:6d} {100 * v['n'] / tot:5.1f}%  {top}")
    for (ln, txt), c in lines[role].most_common(6):
        print(f"      {c:5d}  L{ln}: {txt}")
