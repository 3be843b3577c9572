"""CPU, world_size 2, gloo: the N>1 path of bench.py — env replicas sharded by global index with no data-path
collective.  Each rank steps its shard (1-lane host emulation of the device code); the union must equal an
unsharded run, and the timing reduction must be a max over ranks."""
import os
import socket
import subprocess
import sys

import numpy as np

from tests._gloo_worker import run_shard

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_shards_equal_unsharded_run(tmp_path):
    port = _free_port()
    out = str(tmp_path / "rank%d.npz")
    procs = []
    for rank in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py"), out], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    r0, r1 = np.load(out % 0), np.load(out % 1)
    assert list(r0["seeds"]) == [1000, 1001, 1002] and list(r1["seeds"]) == [1003, 1004, 1005]
    assert float(r0["tmax"]) == 2.0 and float(r1["tmax"]) == 2.0  # max over ranks, on every rank
    loc, rew = run_shard(list(range(1000, 1006)), 12)
    assert np.array_equal(np.concatenate([r0["loc"], r1["loc"]]), loc)
    assert np.array_equal(np.concatenate([r0["rew"], r1["rew"]]), rew)
