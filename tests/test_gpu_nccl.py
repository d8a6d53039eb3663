"""The multi-GPU path proper: one process per GPU under torchrun, receivers sharded by ring-0 range, the sharded fast-round
tally over NCCL (csrc/fast_paxos.cu: count-weighted-sum all-reduce, ambiguous-bucket refinement).  Needs >= 2 GPUs: run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu -q`; on a 1-GPU box it is skipped.  See tests/nccl_worker.py
for what every rank checks against the oracle (>= 256 receivers per rank on every batch of a C4-shaped stream, a dissenting
proposal, a 12-bit bucket collision that forces the refinement path)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_stream_and_tally_against_oracle(world):
    have = _gpus()
    if have < world:
        pytest.skip("needs %d GPUs (this box has %d)" % (world, have))
    from oracle import oracle_py
    oracle_py.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py"), "20000", "256"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:] + "\n" + r.stderr[-6000:])
    assert "nccl worker ok" in r.stdout
