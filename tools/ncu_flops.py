"""Executed FP32 flops of one kernel from an .ncu-rep's source page (no GPU needed): thread-level predicated-on
instruction counts of the FP32 opcodes, weighted FFMA2 = 4, FADD2/FMUL2 = 2, FFMA = 2, FADD/FMUL = 1 flops."""
import csv, io, json, subprocess, sys
W = {"FFMA2": 4, "FADD2": 2, "FMUL2": 2, "FFMA": 2, "FADD": 1, "FMUL": 1}
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, warp = {}, {}
    for r in data:
        tok = r[ix["Source"]].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + float(r[ix["Predicated-On Thread Instructions Executed"]] or 0)
        warp[op] = warp.get(op, 0) + float(r[ix["Instructions Executed"]] or 0)
    flops = sum(ops.get(k, 0) * w for k, w in W.items())
    print(json.dumps({"kernel": rows[0][1][:100], "fp32_flops": flops, "warp_instructions": sum(warp.values()),
                      "packed_warp_instructions": sum(warp.get(k, 0) for k in ("FFMA2", "FADD2", "FMUL2")),
                      "thread_instr": {k: ops.get(k, 0) for k in W}}))
for p in sys.argv[1:]:
    main(p)
