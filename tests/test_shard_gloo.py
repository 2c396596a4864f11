"""CPU, world_size 2 over gloo: the N>1 host logic (index sharding + the statistics all-reduce).
The per-rank statistics are produced with numpy here (no GPU); on the GPU box the same vector
comes from kfpos_batch_error_stats."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from roskfpos_b200.shard import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for n in (1, 7, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from roskfpos_b200.shard import shard_bounds, reduce_stats
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N = 10007
    rng = np.random.default_rng(123)            # same data on every rank, each takes its slice
    err = rng.normal(0, 0.2, size=(3, N))
    lo, hi = shard_bounds(N, rank, world)
    e = err[:, lo:hi]
    local = [float((e ** 2).sum()), float((e[:2] ** 2).sum()), float(hi - lo), 0.0]
    s, rmse, rmse_xy = reduce_stats(local)
    if rank == 0:
        print(json.dumps({"s": list(s), "rmse": rmse, "rmse_xy": rmse_xy, "world": world}))
    dist.destroy_process_group()
""")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_stats_allreduce_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    got = json.loads(line)
    rng = np.random.default_rng(123)
    err = rng.normal(0, 0.2, size=(3, 10007))
    assert got["world"] == 2 and got["s"][2] == 10007
    assert abs(got["s"][0] - (err ** 2).sum()) <= 1e-9 * (err ** 2).sum()
    assert abs(got["rmse"] - np.sqrt((err ** 2).sum() / 10007)) < 1e-12
    assert abs(got["rmse_xy"] - np.sqrt((err[:2] ** 2).sum() / 10007)) < 1e-12
