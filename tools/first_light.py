"""Scratch GPU bring-up script: engine vs oracle on the first run of a fixture + a rough timing."""
import sys, time, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import gpsacq_loader, oracle
ga = gpsacq_loader.load()

def check(name, fc, fs, path, nblocks=32):
    data = open(path, "rb").read()[: 5120 * nblocks]
    acq = ga.Acquisition(fc, fs)
    print(name, acq.info)
    o = oracle.Oracle(fc, fs)
    # replica time: bit exact
    bad = 0
    for sv in range(32):
        a = acq.replica_time(sv); b = oracle.replica_time(fs, sv)
        bad += int((a.view(np.uint32) != b.view(np.uint32)).sum())
    print(" replica_time mismatching floats:", bad)
    e = max(np.abs(acq.replica_spectrum(sv) - o.code_spectrum(sv)).max() / np.abs(o.code_spectrum(sv)).max() for sv in (0, 7, 31))
    print(" replica_spectrum rel err:", e)
    t = time.time(); pk = acq.search_blocks(data); print(" search_blocks wall", time.time() - t, acq.stage_times())
    xs = acq.block_spectrum(0); xo = o.sample(data[:5120])
    print(" block_spectrum rel err:", np.abs(xs - xo).max() / np.abs(xo).max())
    po = o.search_blocks(data)
    for b in (0, 4, 7, 31):
        cs = acq.cell_stats(b); mp, mi, tp = o.cells(data[b * 5120:(b + 1) * 5120], b % 32)
        print("  blk", b, "cells: max rel", np.abs(cs["max_pwr"] / mp - 1).max(), "tot rel", np.abs(cs["tot_pwr"] / tp - 1).max(),
              "argmax mismatches", int((cs["max_idx"] != mi).sum()))
    print(" lo_shift equal:", np.array_equal(pk["lo_shift"], po["lo_shift"]), " ca_shift equal:", np.array_equal(pk["ca_shift"], po["ca_shift"]),
          " snr rel:", np.abs(pk["snr"] / po["snr"] - 1).max(), " flags equal:", np.array_equal(pk["flags"], po["flags"]))
    print(ga.format_run(0, pk[:32]))
    # rough throughput: repeat a 512-block batch
    reps = (512 * 5120) // len(data) + 1
    big = (data * reps)[: 512 * 5120]
    acq.search_blocks(big)
    t = time.time(); acq.search_blocks(big); dt = time.time() - t
    st = acq.stage_times()
    ncell = 512 * acq.n_doppler
    print(" 512-block batch: wall %.2f ms, stages %s, cells/s (cell kernel) %.3e, e2e %.3e" % (dt * 1e3, st, ncell / (st["cells_ms"] * 1e-3), ncell / dt))
    acq.close()

check("nottingham", 4.092e6, 5.456e6, ROOT / "tests/golden/nottingham_fs5456_if4092_runs0-3.bin")
check("gps_sig", 2.046e6, 8.184e6, ROOT / "tests/golden/gps_sig_fs8184_if2046_runs0-1.bin")
