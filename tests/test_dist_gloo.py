"""world_size-2 run of the multi-GPU plumbing on CPU (gloo): contiguous sharding of the chain list and the single
all-gather the thermodynamic-integration tail needs.  No collective exists on the hot path itself."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pyhillfit_b200 import dist as pd
    from pyhillfit_b200 import ti
    pd.init_process_group("gloo")
    assert pd.world() == (world, rank, rank)
    temps = ti.temperature_ladder()
    n_pairs, R = 5, 3
    ids, tt = ti.build_chain_list(n_pairs, temps, R)
    weights = np.array([2, 4, 4, 5, 4], dtype=float)[ids]
    bounds = pd.shard_bounds(weights, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # stand-in for the per-chain means a rank's sampler would hold: a known function of the global chain index
    local = torch.from_numpy(np.sin(np.arange(lo, hi)) * 10.0 - tt[lo:hi])
    full = pd.all_gather_varlen(local, bounds).numpy()
    want = np.sin(np.arange(len(ids))) * 10.0 - tt
    assert np.array_equal(full, want)
    means = full.reshape(n_pairs, len(temps), R).mean(axis=2)
    log_py = ti.log_py_from_means(temps, means)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.concatenate([[lo, hi], log_py]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_and_gather(tmp_path):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 5 * 41 * 3          # shards tile the chain list
    assert np.array_equal(r0[2:], r1[2:])                                   # every rank ends with the same integrals
    assert abs((r0[1] - r0[0]) - (r1[1] - r1[0])) < 41 * 3 * 2              # balanced by weight, not by count
