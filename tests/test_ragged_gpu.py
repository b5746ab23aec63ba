"""GPU: the multi-rank ragged-strip merge (peer stores through CUDA-IPC mapped memory + stream-ordered flags) must equal
the single-GPU merge bit for bit, with every pixel written by exactly one rank. Runs `world` PROCESSES; on a 1-GPU box they
share cuda:0 (IPC between processes works on one device too), on a multi-GPU box each takes its own GPU (NVLink peer stores)."""
import glob
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, tmp_path, same_gpu):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(REPO, "tests", "ragged_worker.py"), str(r), str(world), str(port),
                               str(tmp_path)] + (["same_gpu"] if same_gpu else []),
                              cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    checked = 0
    for ref_path in sorted(glob.glob(os.path.join(str(tmp_path), "case*_ref.pt"))):
        ci = os.path.basename(ref_path).split("_")[0]
        ref = torch.load(ref_path)
        assert not torch.isnan(ref).any()
        parts = [torch.load(os.path.join(str(tmp_path), f"{ci}_rank{r}.pt")) for r in range(world)]
        written = torch.stack([~torch.isnan(p) for p in parts]).sum(0)
        assert int(written.min()) == 1 and int(written.max()) == 1, f"{ci}: every pixel is written by exactly one rank"
        merged = torch.zeros_like(ref)
        for p in parts:
            merged = torch.where(torch.isnan(p), merged, p)
        assert torch.equal(merged, ref), ci
        checked += 1
    assert checked >= 3


@pytest.mark.parametrize("world", [2, 3])
def test_ragged_strips_equal_the_single_gpu_merge(world, tmp_path):
    same_gpu = torch.cuda.device_count() < world
    _run(world, tmp_path, same_gpu)
