"""CPU-only hygiene: every `file:line` citation of the reference in the headers, the sources and the design documents points
inside that file (checked where /root/reference exists: the development container; skipped on the GPU box)."""
import re
from pathlib import Path

import pytest

from conftest import ROOT

REF = Path("/root/reference")
PAT = re.compile(r"((?:c/)?[A-Za-z0-9_.]+\.(?:cpp|h|m|md)):(\d+)(?:-(\d+))?")


def _sources():
    out = [ROOT / "include" / "gpsacq.h", ROOT / "INTEGRATION.md", ROOT / "DESIGN.md", ROOT / "README.md"]
    for top, exts in (("gnss-gps-sdr_b200", (".h", ".cuh", ".cu", ".cpp", ".py")), ("oracle", (".c", ".py", ".cpp", ".h"))):
        out += [p for p in (ROOT / top).rglob("*") if p.suffix in exts and "_ref" not in p.parts]
    return out


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only in the development container")
def test_reference_citations_point_inside_the_cited_files():
    n_lines, checked, bad = {}, 0, []
    for f in _sources():
        for m in PAT.finditer(f.read_text(errors="ignore")):
            p = next((c for c in (REF / m.group(1), REF / "c" / m.group(1)) if c.is_file()), None)
            if p is None:
                continue                                   # not a reference file (this repository's own files are cited too)
            if p not in n_lines:
                n_lines[p] = p.read_text(errors="ignore").count("\n") + 1
            lo, hi = int(m.group(2)), int(m.group(3) or m.group(2))
            checked += 1
            if not (1 <= lo <= hi <= n_lines[p]):
                bad.append(f"{f.relative_to(ROOT)}: {m.group(0)} (file has {n_lines[p]} lines)")
    assert checked > 100 and not bad, bad
