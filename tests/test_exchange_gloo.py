"""N > 1 host logic on CPU: one process per domain over gloo (world_size 2 and 4).  The schedule
and neighbour tables the CUDA library uses for its NCCL exchange are executed with gloo
point-to-point operations and compared with the all-ranks model of comms.c."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dims", ["2,1,1", "2,2,1"])
def test_exchange_schedule_over_gloo(built, dims):
    world = eval(dims.replace(",", "*"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "_gloo_exchange_worker.py"), dims]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(", ok") == world
