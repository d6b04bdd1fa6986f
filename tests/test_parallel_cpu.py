"""Host-side multi-GPU logic on CPU: the shard split and the profile reduction with gloo, world size 2."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_shard_split_is_the_reference_split(lib):
    """xrays.cpp:423-432: batch = N/G, the first N%G shards get one more ray, contiguous."""
    from graph_framework_b200.rays import shard_sizes, shard_offsets
    assert shard_sizes(10, 4) == [3, 3, 2, 2]
    assert shard_sizes(100000, 8) == [12500]*8
    assert shard_sizes(3, 8) == [1, 1, 1, 0, 0, 0, 0, 0]
    for total, shards in ((1000003, 8), (17, 5), (1, 1)):
        sizes = shard_sizes(total, shards)
        assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
        assert sorted(sizes, reverse=True) == sizes
        offs = shard_offsets(total, shards)
        assert offs[0] == 0 and offs[-1] == total


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from graph_framework_b200 import parallel
from oracle import port
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1001
rng = np.random.default_rng(0)
state = {"x": rng.uniform(-1, 1, n), "y": rng.uniform(-1, 1, n), "z": rng.uniform(-1, 1, n), "w": rng.uniform(0, 1, n)}
mine = parallel.shard_state(state)
off, size = parallel.my_shard(n)
assert size == len(mine["x"]) and np.array_equal(mine["x"], state["x"][off:off + size])
lo, hi, bins = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), (4, 5, 6)
local = port.deposit(mine["x"], mine["y"], mine["z"], mine["w"], lo, hi, bins)
hist = torch.from_numpy(local.copy())
parallel.allreduce_profile(hist, total_rays=n)
full = port.deposit(state["x"], state["y"], state["z"], state["w"], lo, hi, bins)/n
assert np.allclose(hist.numpy(), full, rtol=1e-13, atol=1e-300), np.abs(hist.numpy() - full).max()
# the same with the blocks of a trace overlapped: every rank deposits 7 blocks of its shard, the reductions run asynchronously
reducer = parallel.OverlappedProfileReducer(bins, total_rays=n)
expected = np.zeros(bins)
for block in range(7):
    w = state["w"]*(block + 1)
    buf = reducer.begin_block()
    buf += torch.from_numpy(port.deposit(mine["x"], mine["y"], mine["z"], w[off:off + size], lo, hi, bins))
    reducer.end_block()
    expected += port.deposit(state["x"], state["y"], state["z"], w, lo, hi, bins)
got = reducer.finish().numpy()
assert np.allclose(got, expected/n, rtol=1e-13, atol=1e-300), np.abs(got - expected/n).max()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_profile_allreduce_two_ranks_gloo(lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port_no), str(script)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
