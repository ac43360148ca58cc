"""gnss-gps-sdr_b200 -- B200-native GPS L1 C/A acquisition engine (host-side Python mirror).

The product is ``csrc/libgpsacq.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/gpsacq.h``) plus the C++ host tree in ``c/`` that keeps the reference's
``gps_test`` / ``SearchInit`` / ``SearchTask`` / ``CACODE`` surface.  This package is a
thin ctypes binding over the same C ABI, used by the tests and by bench.py.

The directory name contains a hyphen (it is the name the build contract asks for), so
import it through ``load()`` in the repo-root ``gpsacq_loader.py`` or add the directory's
parent to ``sys.path`` and use ``importlib``; inside the package everything is relative.
"""
from .acq import (Acquisition, AcquisitionGroup, SearchService, handoff, HANDOFF_DTYPE, EVENT_DTYPE, synth_capture_gpu, bits_to_iq8, bits_to_iq8_device, sig_gen_literal, GpsAcqError, PEAK_DTYPE, CELL_DTYPE, lib_path, load_library,
                  search_task_text, format_run, NUM_SATS, FFT_LEN, SNR_THRESHOLD)

__all__ = ["Acquisition", "AcquisitionGroup", "SearchService", "handoff", "HANDOFF_DTYPE", "EVENT_DTYPE", "synth_capture_gpu", "bits_to_iq8", "bits_to_iq8_device", "sig_gen_literal", "GpsAcqError", "PEAK_DTYPE", "CELL_DTYPE", "lib_path", "load_library",
           "search_task_text", "format_run", "NUM_SATS", "FFT_LEN", "SNR_THRESHOLD"]
