"""CPU: FusedStep's multi-stream schedules are free of data races.  tests/_schedule_check.py replays every schedule (default,
cooperative, fp32 engine, data parallel) against fake streams / events that record the
ordering the real ones impose, and checks that every pair of kernel calls touching overlapping bytes (at least one write) is
ordered by stream order, wait_event or wait_stream -- across three consecutive steps, evaluate() and flush().  The weight-tile
prefetch that programmatic dependent launches issue ahead of griddepcontrol.wait is modelled as a separate, earlier read.
Runs in a subprocess because it monkeypatches torch.cuda."""
import os
import subprocess
import sys


def test_every_conflicting_pair_of_launches_is_ordered():
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "_schedule_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SCHEDULES OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    assert "self-test: missing wait_event        2 calls on 2 streams, 1 unordered" in r.stdout      # the checker can see a race
    for tag in ("default (norm-free update)", "cooperative clip+Adam", "data parallel, peer kernel", "data parallel, overlapped all-reduce"):
        assert tag in r.stdout
