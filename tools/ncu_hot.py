"""Dev tool: the hottest SASS instructions (by warp-stall samples) of one kernel of an `ncu --set full --import-source
on` report.   usage: python tools/ncu_hot.py rep.ncu-rep kernel-regex [launch-skip] [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], []
for r in rows[h + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break                                              # the next kernel's block
    if len(r) == len(hdr):
        data.append(r)
ix = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(rows[0][1] if rows and len(rows[0]) > 1 else "", "| samples", tot, "| SASS instructions", len(data))
agg = {}
for r in data:
    for k in hdr:
        if k.startswith("stall_") and "(Not" not in k and r[ix[k]] not in ("", "0"):
            agg[k] = agg.get(k, 0) + int(r[ix[k]])
print("stall mix:", ", ".join(f"{k[6:]} {v * 100 / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for pos, r in sorted(enumerate(data), key=lambda pr: -int(pr[1][ix["# Samples"]]))[:top_n]:
    st = {k[6:]: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "(Not" not in k and r[ix[k]] not in ("", "0")}
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[ix['# Samples']]) * 100 / max(tot, 1):5.1f}%  #{pos:4d} {r[ix['Source']].strip()[:64]:64s} exec={r[ix['Instructions Executed']]:>8s} "
          f"smem_excess={r[ix['L1 Wavefronts Shared Excessive']]:>7s} {best}")
if len(sys.argv) > 5:                                      # samples per block of `bin` consecutive instructions
    b = int(sys.argv[5])
    for lo in range(0, len(data), b):
        blk = data[lo:lo + b]
        sm = sum(int(r[ix["# Samples"]]) for r in blk)
        ex = max(int(r[ix["Instructions Executed"]]) for r in blk)
        first = next((r[ix["Source"]].strip() for r in blk if any(t in r[ix["Source"]] for t in ("UTC", "LDGSTS", "BAR", "LDTM", "SYNCS", "SHFL", "STG", "ATOM"))), "")
        print(f"  #{lo:4d}-{lo + len(blk) - 1:4d}  {sm * 100 / max(tot, 1):5.1f}%  max exec {ex:9d}  {first[:50]}")
