"""Runs LAST (file name): the whole-file acceptance test must not silently fall back to the strided runs."""
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
FULL_CAPTURE = ROOT / "oracle" / "_ref" / "data" / "gps.samples.1bit.I.fs5456.if4092.bin"
RUN_BYTES = 32 * 5120


def test_whole_capture_is_staged_on_this_box():
    """The 55 MB capture is git-ignored but travels with gpurun snapshots (oracle/_ref/data, staged by
    `make -C oracle` where /root/reference exists).  Fails -- does not skip -- when it is missing, so that a
    round whose whole-file parity (tests/test_gpu_parity.py::test_whole_capture_*) only saw the strided runs is visible."""
    assert FULL_CAPTURE.exists() and 340 * RUN_BYTES <= FULL_CAPTURE.stat().st_size < 341 * RUN_BYTES, \
        "oracle/_ref/data/gps.samples.1bit.I.fs5456.if4092.bin did not travel: only the strided runs were compared"
