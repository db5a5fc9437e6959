"""TEST INFRASTRUCTURE: runs bench.py's own main() against the interpreted library (tests/emu/) on a box without a GPU, with
torch.cuda's device calls stubbed, so that edits to bench.py are exercised end to end (JSON contract, program marshalling,
e2e leg) before the next device run.  The numbers it prints are meaningless; tests/test_emu_kernels.py only checks the
line's keys and the count."""
import runpy
import sys
import time

import torch


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.Event = _Event
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{**k, "device": "cpu"}) if k.get("device") == "cuda" else _tensor(*a, **k)

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    runpy.run_path("bench.py", run_name="__main__")
