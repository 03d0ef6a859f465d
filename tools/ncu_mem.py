#!/usr/bin/env python
"""Per-instruction shared/global memory table of an ncu --set full --import-source capture (first launch):
python tools/ncu_mem.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page is a sequence of blocks: ["Kernel Name", name], header, data...
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["data"].append(r)
for b in blocks:
    if sub not in b["name"]:
        continue
    idx = {h: i for i, h in enumerate(b["hdr"])}
    tot_w = tot_i = tot_x = 0
    print(b["name"][:90])
    print("| instruction | executed M | smem wavefronts M | ideal M | excessive M | L2 sectors M | samples |")
    for r in b["data"]:
        try:
            w = int(r[idx["L1 Wavefronts Shared"]] or 0); wi = int(r[idx["L1 Wavefronts Shared Ideal"]] or 0)
            wx = int(r[idx["L1 Wavefronts Shared Excessive"]] or 0); l2 = int(r[idx["L2 Theoretical Sectors Global"]] or 0)
            ex = int(r[idx["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        tot_w += w; tot_i += wi; tot_x += wx
        if w > 200000 or l2 > 200000:
            print(f"| `{r[idx['Source']].strip()[:60]}` | {ex/1e6:.2f} | {w/1e6:.2f} | {wi/1e6:.2f} | {wx/1e6:.2f} | {l2/1e6:.2f} | {r[idx['# Samples']]} |")
    print(f"total smem wavefronts {tot_w/1e6:.1f} M, ideal {tot_i/1e6:.1f} M, excessive {tot_x/1e6:.1f} M\n")
    # stall totals
    st = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    tot = {k: 0 for k in st}
    for r in b["data"]:
        for k in st:
            try: tot[k] += int(r[idx[k]] or 0)
            except (ValueError, IndexError): pass
    s = sum(tot.values())
    print("stall samples: " + ", ".join(f"{k[6:]} {100*v/s:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.005*s))
    break
