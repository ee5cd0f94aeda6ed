"""A short run of the reader fuzz harness (tools/fuzz_readers.py) inside the CPU suite: mutated DCD / XTC / TRR / HDF5 files and
mutated text inputs of a job must be read or refused with host.HostError -- no other exception, crash or hang.  The long runs
against the AddressSanitizer build are recorded in profiles/r02_asan_host.txt (tools/asan_host.sh, tools/fuzz_readers.sh)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [21, 22])
def test_mutated_inputs_are_read_or_refused(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_readers.py"), "60", str(seed)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "findings: 0" in r.stdout
