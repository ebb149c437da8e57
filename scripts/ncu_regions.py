"""Aggregates ncu warp-state samples of one kernel by code region (between sync/memory markers).
usage: python scripts/ncu_regions.py file.ncu-rep [min_samples] [kernel_index]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = rows[starts[kidx]:starts[kidx + 1]]
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 5]
ix = {h: i for i, h in enumerate(hdr)}
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 60
tot = sum(int(r[ix['# Samples']]) for r in data)
base = int(data[0][ix['Address']], 16)
print('kernel', rows[0][1][:90]); print('total samples', tot, 'instructions', len(data))
prev = 0
for i, r in enumerate(data):
    s = r[ix['Source']]
    if any(k in s for k in ('SYNCS', 'BAR.SYNC', 'UTMALDG', 'EXIT', 'ATOMS', 'STG', 'LDG', 'SHFL')) or i == len(data) - 1:
        seg = data[prev:i + 1]
        c = sum(int(x[ix['# Samples']]) for x in seg)
        if c >= mins:
            st = {}
            for x in seg:
                for h in hdr:
                    if h.startswith('stall_') and 'Not Issued' not in h and x[ix[h]] not in ('', '0'):
                        st[h] = st.get(h, 0) + int(x[ix[h]])
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"{int(seg[0][ix['Address']],16)-base:6x}-{int(r[ix['Address']],16)-base:6x} n={len(seg):4d} samples={c:5d} ({100*c/tot:4.1f}%) "
                  f"ends@ {s.strip()[:52]:52s} exec={r[ix['Instructions Executed']]:>8s} {top}")
        prev = i + 1
