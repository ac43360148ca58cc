"""Per-phase (between block barriers) instruction mix and stall shares of one kernel from an .ncu-rep's source page."""
import csv, io, subprocess, sys
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print("##", rows[0][1][:110])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    f = lambda r, k: float(r[ix[k]] or 0) if r[ix[k]].replace('.', '', 1).isdigit() else 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(f(r, "# Samples") for r in data); tot_i = sum(f(r, "Instructions Executed") for r in data)
    print("SASS instructions", len(data), " warp-instructions executed %.3g" % tot_i, " samples", int(tot_s))
    phase, agg = 0, {}
    for r in data:
        src = r[ix["Source"]]
        a = agg.setdefault(phase, {"n": 0, "samples": 0, "inst": 0, "ops": {}, **{s: 0 for s in stalls}})
        a["n"] += 1; a["samples"] += f(r, "# Samples"); a["inst"] += f(r, "Instructions Executed")
        tok = src.split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        a["ops"][op] = a["ops"].get(op, 0) + f(r, "Instructions Executed")
        for s in stalls: a[s] += f(r, s)
        if "BAR.SYNC" in src: phase += 1
    for p, a in agg.items():
        print("phase %d: %4d SASS, %5.1f%% of samples, %5.1f%% of executed instructions" % (p, a["n"], 100 * a["samples"] / tot_s, 100 * a["inst"] / tot_i))
        print("   ops:   ", ", ".join("%s %.1f%%" % (k, 100 * v / tot_i) for k, v in sorted(a["ops"].items(), key=lambda x: -x[1])[:12]))
        print("   stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s) for k, v in sorted(((s, a[s]) for s in stalls), key=lambda x: -x[1])[:6]))
for p in sys.argv[1:]:
    main(p)
