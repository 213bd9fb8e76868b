"""CPU: FusedStep / legacy API / inference host orchestration against a fake backend that type-checks every C-ABI
call (argument count and kinds per codae._C.SIGNATURES).  Runs in a subprocess because it monkeypatches torch."""
import os
import subprocess
import sys


def test_host_paths_issue_well_typed_kernel_sequences():
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "_host_dryrun.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DRY RUN OK" in r.stdout
    assert "fp32 launches 27" in r.stdout and "bf16 launches 27" in r.stdout   # 1 + 8 + 1 + 8 + 7 + counter + fused clip/Adam
