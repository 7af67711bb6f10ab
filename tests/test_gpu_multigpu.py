"""Multi-GPU parity: the row-sharded persistent network kernel — in-kernel all-gather of the
stage vectors as LL packets over NVLink — against the oracle on the FULL state, on every rank, at the fp64 tolerance
1e-9.  Covers dense dot coupling (C3 shape, N = 1024, Euler + RK4), the global `vsum` coupling (kuramoto_scalar: every rank
reduces its complete replica), separable pair coupling (two phases per evaluation), gamma cascades + delay ring buffers
(C4 shape) and gather / scatter edges.  Spawns `torch.distributed.run` with one process per GPU; skipped on boxes with fewer
than two GPUs (run it with `gpurun --gpus 2`; the log of this round's run is profiles/r02_multigpu_tests.txt)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_network_matches_oracle_on_every_rank(gpu, world):
    from pyrates_b200 import runtime as rt
    if rt.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {rt.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=850)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MULTIGPU_RESULT ")]
    assert p.returncode == 0 and lines, p.stdout[-3000:] + p.stderr[-3000:]
    res = json.loads(lines[-1][len("MULTIGPU_RESULT "):])
    assert res["ok"] and res["world"] == world
    for per_rank in res["errors_by_rank"]:
        assert all(e <= 1e-9 for e in per_rank.values()), per_rank
