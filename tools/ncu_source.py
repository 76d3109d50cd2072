"""Per-instruction view of one .ncu-rep (development aid): executed count per tile and stall samples."""
import csv, subprocess, sys
rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 2**31 / 512
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]; rows = rows[2:]
isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iex]) for r in rows); ts = sum(int(r[ismp]) for r in rows)
print(f"inst/tile {tot / tiles:.2f}  samples {ts}")
for i, r in enumerate(rows):
    top = sorted(((int(r[j] or 0), hdr[j][6:]) for j in stalls), reverse=True)[:2]
    tops = " ".join(f"{n}:{c}" for c, n in top if c)
    print(f"{i:4d} {int(r[iex]) / tiles:7.3f} {100.0 * int(r[ismp]) / ts:6.2f}%  {r[isrc].strip():70s} {tops}")
