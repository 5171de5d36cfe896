"""Vendor the reference's own test-suite, byte for byte, into tests/ref_tests/ (SURVEY section 4).

    python scripts/vendor_ref_tests.py            # run in the build container, where /root/reference exists

The eight files under /root/reference/tests (23 tests) are the reference's acceptance tests for the very API the
mirror package `dl_scl_polar` re-implements on the GPU.  They are TEST INFRASTRUCTURE: nothing under polar_code_b200/
imports them, they are copied unmodified (tests/ref_tests/MANIFEST.json records the SHA-256 of every source file so a
reviewer can check that), and they run against the mirror because the repo root provides a top-level `dl_scl_polar`
alias.  /root/reference does not exist on the GPU box and the mirror has no CPU fallback, which is why the files have
to travel with the repo instead of being collected in place; tests/conftest.py marks them `gpu` at collection time.
"""
import hashlib
import json
import shutil
import sys
from pathlib import Path

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tests")
DST = Path(__file__).resolve().parents[1] / "tests" / "ref_tests"

if __name__ == "__main__":
    DST.mkdir(parents=True, exist_ok=True)
    manifest = {}
    for f in sorted(SRC.glob("test_*.py")):
        shutil.copyfile(f, DST / f.name)
        manifest[f.name] = hashlib.sha256(f.read_bytes()).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps({"source": "heimrih/polar_code tests/ (unmodified)", "sha256": manifest}, indent=1) + "\n")
    print(f"vendored {len(manifest)} files into {DST}")
