"""GPU: the database-sharded kNN with the fused peer-memory exchange (orbm_exchange_* / orbm_knn2_sharded_device, the path
bench.py times at N > 1) against the CPU oracle, with world sizes 2 and 3 as separate processes sharing one GPU -- CUDA IPC
and the flag protocol do not care whether the peers' buffers live on another device, so the N > 1 data path is covered by
`pytest -m gpu` on a one-GPU box (the NVLink loads themselves are exercised by bench.py --gpus N, which carries the same
oracle check in its JSON line: match.parity_checked)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,ndb,nq", [(2, 30000, 3000), (3, 20001, 1111)])
def test_fused_exchange_matches_oracle(orbx, world, ndb, nq):
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_sharded_exchange_worker.py"), str(ndb), str(nq)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
