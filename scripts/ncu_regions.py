"""SASS regions of equal execution count from an .ncu-rep (source page): where the warp instructions go.
   python scripts/ncu_regions.py rep.ncu-rep [launch] [min_pct]"""
import csv, io, subprocess, sys
def main(rep, launch=0, minpct=0.8):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr)]
    # the export repeats the function once per source view: keep the first copy
    seen = set(); d2 = []
    for r in data:
        if not r[0].startswith("0x"): continue
        if r[0] in seen: break
        seen.add(r[0]); d2.append(r)
    data = d2
    ia, isrc, ins, ii, it = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
    def F(x):
        try: return float(x)
        except ValueError: return 0.0
    data = [r for r in data if r[0].startswith("0x")]
    ti = sum(F(r[ii]) for r in data); ts = sum(F(r[ins]) for r in data)
    print(f"{len(data)} SASS lines, {ti:.4g} warp instructions, {ts:.0f} samples")
    reg = []
    for k, r in enumerate(data):
        c = F(r[ii])
        if reg and abs(c - reg[-1]["c"]) <= 0.03 * max(c, reg[-1]["c"], 1):
            g = reg[-1]; g["end"] = k; g["n"] += 1; g["sum"] += c; g["smp"] += F(r[ins]); g["thr"] += F(r[it]) * c
        else:
            reg.append({"start": k, "end": k, "c": c, "n": 1, "sum": c, "smp": F(r[ins]), "thr": F(r[it]) * c})
    base = int(data[0][ia], 16)
    for g in reg:
        if 100 * g["sum"] / ti >= minpct:
            print(f"sass[{g['start']:5d}..{g['end']:5d}] +{int(data[g['start']][ia],16)-base:#06x} n={g['n']:4d} exec={g['c']:9.4g} "
                  f"instr%={100*g['sum']/ti:5.1f} samples%={100*g['smp']/ts:5.1f} cyc/instr={(g['smp']/ts)/(g['sum']/ti):4.2f} thr={g['thr']/max(g['sum'],1):4.1f}  {data[g['start']][isrc].strip()[:44]}")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, float(sys.argv[3]) if len(sys.argv) > 3 else 0.8)
