"""Summarise an `ncu --page source --csv --print-source sass` dump: opcode histogram by executed instructions,
stall samples by reason, per kernel.  Usage: python tools/ncu_src_summary.py file.csv [kernel-index]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kernels, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
sel = int(sys.argv[2]) if len(sys.argv) > 2 else None
for ki, k in enumerate(kernels):
    if sel is not None and ki != sel:
        continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    ops = collections.Counter(); samples = collections.Counter(); stalls = collections.Counter()
    tot = 0
    for r in k["rows"]:
        src = r[h["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.rstrip(";")
        n = int(r[h["Instructions Executed"]] or 0)
        ops[op] += n; tot += n
        samples[op] += int(r[h["# Samples"]] or 0)
        for name in h:
            if name.startswith("stall_") and "(Not Issued)" not in name:
                stalls[name] += int(r[h[name]] or 0)
    print(f"== kernel {ki}: {k['name']}  warp-instructions={tot}")
    for op, n in ops.most_common(28):
        print(f"   {op:28s} {n:12d} {100.0*n/tot:6.2f}%   samples {samples[op]}")
    st = sum(stalls.values())
    print("   stalls:", ", ".join(f"{n[6:]}={100.0*v/st:.1f}%" for n, v in stalls.most_common(8)))
