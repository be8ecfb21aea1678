"""Multi-GPU, multi-process parity of the ghost-shell modes (needs >= 2 GPUs; skipped on a single-GPU box).
The actual checks live in tests/mp_ghost_modes.py, launched here with torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(280)
def test_ghost_modes_agree_across_processes():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2                      # the configuration this script was validated with (2 x B200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "mp_ghost_modes.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "MP_GHOST_MODES_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
