"""GRID-mode throughput on one GPU for the BASELINE.json configs[1..4] shapes (secondary numbers;
bench.py's headline is the REF-mode workload).  Prints one JSON line per config."""
import importlib, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gpsacq_loader
ga = gpsacq_loader.load()
sg = importlib.import_module("gnss_gps_sdr_b200.siggen")
import torch

CONFIGS = {
    "C1": dict(fs=5.456e6, fc=4.092e6, max_fo=5000.0, step=500.0, K=1, n_acq=64),
    "C2": dict(fs=8.184e6, fc=2.046e6, max_fo=5000.0, step=500.0, K=1, n_acq=64),
    "C3": dict(fs=2.8e6, fc=0.62e6, max_fo=100000.0, step=250.0, K=10, n_acq=4),
    "C4": dict(fs=8.184e6, fc=2.046e6, max_fo=100000.0, step=100.0, K=10, n_acq=1),
    "X10": dict(fs=10e6, fc=2.6e6, max_fo=5000.0, step=500.0, K=1, n_acq=64),     # the receiver's own rate (c/gps.h:23-24): exact-length W = 10000
}
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
for name in (sys.argv[1:] or list(CONFIGS)):
    c = CONFIGS[name]
    W = int(round(c["fs"] / 1000))
    sats = sg.default_constellation(c["fs"], seed=1575420000, max_doppler=0.9 * c["max_fo"])
    bits = sg.synth_capture(W * c["K"] * c["n_acq"], c["fs"], c["fc"], sats, seed=3)
    acq = ga.Acquisition(c["fc"], c["fs"], c["max_fo"], mode=1, doppler_step=c["step"], noncoh_blocks=c["K"], max_blocks=c["n_acq"])
    n_acq = min(c["n_acq"], acq.info["max_acq"])
    dev = torch.device("cuda", 0)
    d_bits = torch.from_numpy(bits[: n_acq * acq.acq_bytes]).to(dev)
    d_out = torch.zeros(n_acq * 32 * 32, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); acq.set_stream(stream.cuda_stream)
    for _ in range(3):
        acq.acquire_device(d_bits.data_ptr(), n_acq, d_out.data_ptr())
    torch.cuda.synchronize()
    reps = 10 if name in ("C1", "C2") else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        acq.acquire_device(d_bits.data_ptr(), n_acq, d_out.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = acq.stage_times()
    pk = np.frombuffer(d_out.cpu().numpy().tobytes(), ga.PEAK_DTYPE)
    D = acq.n_doppler
    corr = n_acq * 32 * D * c["K"]
    bpc = acq.info["bytes_per_corr"]
    print(json.dumps({"config": name, **{k: c[k] for k in ("fs", "max_fo", "step", "K")}, "n_doppler": D, "window": W,
                      "acquisitions_per_batch": n_acq, "ms_per_batch": ms, "correlations_per_s": corr / ms * 1e3,
                      "cells_per_s": corr / c["K"] / ms * 1e3, "acquisitions_per_s": n_acq / ms * 1e3,
                      "stage_ms": {k: round(v, 3) for k, v in st.items()},
                      "bytes_per_corr": bpc, "contract_gbs": corr * bpc / ms / 1e6, "frac_of_hbm_peak": corr * bpc / ms / 1e6 / peak,
                      "detected": int((pk["snr"] >= 25).sum()), "peaks_sha256": __import__("hashlib").sha256(pk.tobytes()).hexdigest()[:16], "fft_len_embedded": acq.info["fft_len"],
                      "cell_threads": acq.info["cell_threads"], "cell_ctas": acq.info["cell_ctas"]}))
    acq.close()
