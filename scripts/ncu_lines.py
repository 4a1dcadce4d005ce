"""Per-source-line share of stall samples and executed instructions from an .ncu-rep captured with
--import-source on (run here, CPU only):  python scripts/ncu_lines.py rep.ncu-rep [launch] [file-substring]"""
import csv, io, subprocess, sys, collections

def main(rep, launch=0, fsub="bsw_duo.cuh", thresh=0.4):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur = None; hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1]; hdr = None; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or len(r) < len(hdr): continue
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); i_t = hdr.index("Thread Instructions Executed")
        key = (cur, r[0])
        a = agg.setdefault(key, [r[1], 0.0, 0.0, 0.0, collections.Counter()])
        try:
            a[1] += float(r[i_s] or 0); a[2] += float(r[i_i] or 0); a[3] += float(r[i_t] or 0)
        except ValueError:
            continue
        for name in ("stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_no_inst", "stall_math",
                     "stall_dispatch", "stall_not_selected", "stall_selected"):
            if name in hdr:
                try: a[4][name] += float(r[hdr.index(name)] or 0)
                except ValueError: pass
    ts = sum(a[1] for a in agg.values()); ti = sum(a[2] for a in agg.values())
    print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
    perfile = collections.Counter(); perfile_i = collections.Counter()
    for (f, ln), a in agg.items():
        perfile[f] += a[1]; perfile_i[f] += a[2]
    for f in perfile: print(f"  {f}: samples {100*perfile[f]/ts:.1f}%  instr {100*perfile_i[f]/ti:.1f}%")
    print("line samples% instr%  thr/inst  top stalls | source")
    for (f, ln), a in agg.items():
        if fsub not in f: continue
        if 100 * a[1] / ts >= thresh or 100 * a[2] / ti >= thresh:
            top = ",".join(f"{k[6:]}:{100*v/ts:.1f}" for k, v in a[4].most_common(3) if v)
            print(f"{ln:>4s} {100*a[1]/ts:6.2f} {100*a[2]/ti:6.2f} {a[3]/max(a[2],1):6.1f}  {top:40s} | {a[0].strip()[:100]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, sys.argv[3] if len(sys.argv) > 3 else "bsw_duo.cuh")
