#!/bin/bash
# round 2, session zzz (the last 40 seconds of GPU budget): tests/host/host_globals.cpp -- a caller that changes FC / max_fo
# between SearchInit() and SearchTask() -- against the golden stdout, without pytest / torch start-up
mkdir -p gpurun_out
P=gnss-gps-sdr_b200
g++ -O1 -std=c++17 -I$P/c tests/host/host_globals.cpp $P/c/search_offline.cpp -o /tmp/hg -L$P/csrc -lgpsacq -Wl,-rpath,$PWD/$P/csrc -lpthread || exit 1
for w in fc max_fo; do /tmp/hg tests/golden/nottingham_fs5456_if4092_runs0-3.bin $w > gpurun_out/host_globals_$w.txt 2> gpurun_out/host_globals_$w.err; echo "$w rc=$? t=$SECONDS"; done
cd tests && python - <<'PY'
from conftest import CAPTURES, compare_runs, parse_stdout, strip_banner
c = CAPTURES["nottingham"]
ref_runs, _ = parse_stdout(strip_banner(c["stdout"].read_text()))
for w in ("fc", "max_fo"):
    got, tail = parse_stdout(open(f"../gpurun_out/host_globals_{w}.txt").read())
    assert tail == ["run out of file!"] and len(got) == c["runs"], (w, tail, len(got))
    compare_runs(got, ref_runs[: c["runs"]])
    print(w, "matches the golden stdout:", len(got), "runs")
PY
echo "t=$SECONDS"
