"""world_size-2 gloo run of the sharding logic on CPU: per-rank env slices reproduce the single-process key
tree, and the reporting reductions (max time, summed work, global statistics) are correct."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, total_envs: int, out_dir: str) -> None:
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from ksim_b200 import distributed as D
    from ksim_b200.envinit import make_env_init
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
    from ksim_b200.program import build_program

    r, _, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    spec = make_spec()
    prog = build_program(spec)
    rz = {k: f(spec.model) for k, f in RANDOMIZERS.items()}
    lo, hi = D.shard_range(total_envs, rank, world)
    init = make_env_init(prog, rz, hi - lo, seed=11, env_offset=lo, total_envs=total_envs)
    np.save(os.path.join(out_dir, f"keys_{rank}.npy"), init.init_keys)
    np.save(os.path.join(out_dir, f"rand_{rank}.npy"), init.rand)
    t = D.max_over_ranks(10.0 + rank)
    work = D.sum_over_ranks(float(hi - lo))
    x = torch.arange(lo, hi, dtype=torch.float32)
    mean, std = D.global_mean_std(x)
    # gradient averaging (the PPO step's one collective): mixed shapes / dtypes, small bucket to force several flushes
    grads = [torch.full((5, 3), float(rank + 1)), torch.full((7,), 10.0 * (rank + 1)), torch.full((2, 2), float(rank), dtype=torch.float64)]
    D.allreduce_mean_(grads, bucket_bytes=64)
    # the PPO update's collective as the trainer issues it: one in-place all-reduce of the flat gradient buffer; a rank whose
    # backward left some parameter without a gradient still reduces the same number of elements (zeros in the flat buffer)
    from ksim_b200 import optim, policy

    torch.manual_seed(0)
    mlp = policy.MLP(6, 3, 8, 2, "elu")
    fp = optim.FlatParams(mlp)
    fp.zero_grad()
    xin = torch.full((4, 6), float(rank + 1))
    (mlp.layers[0].reference(xin).sum() if rank == 1 else mlp.reference(xin).sum()).backward()  # rank 1: only layer 0 gets a gradient
    local = fp.grad.clone()
    scale = fp.allreduce_grads()
    np.save(os.path.join(out_dir, f"flat_local_{rank}.npy"), local.numpy())
    np.save(os.path.join(out_dir, f"flat_reduced_{rank}.npy"), (fp.grad * scale).numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), np.array([t, work, mean, std]))
        np.save(os.path.join(out_dir, "grads.npy"), np.array([float(g.reshape(-1)[0]) for g in grads]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path) -> None:  # noqa: ANN001
    from ksim_b200.distributed import shard_range
    from ksim_b200.envinit import make_env_init
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
    from ksim_b200.program import build_program

    total, world = 13, 2  # ragged on purpose
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    spec = make_spec()
    prog = build_program(spec)
    full = make_env_init(prog, {k: f(spec.model) for k, f in RANDOMIZERS.items()}, total, seed=11)
    keys = np.concatenate([np.load(tmp_path / f"keys_{r}.npy") for r in range(world)])
    rand = np.concatenate([np.load(tmp_path / f"rand_{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(keys, full.init_keys)
    np.testing.assert_array_equal(rand, full.rand)
    t, work, mean, std = np.load(tmp_path / "reduced.npy")
    assert t == 11.0 and work == total
    assert mean == pytest.approx(np.arange(total).mean()) and std == pytest.approx(np.arange(total).std())
    assert [shard_range(total, r, world) for r in range(world)] == [(0, 7), (7, 13)]
    np.testing.assert_allclose(np.load(tmp_path / "grads.npy"), [1.5, 15.0, 0.5])
    local = [np.load(tmp_path / f"flat_local_{r}.npy") for r in range(world)]
    for r in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"flat_reduced_{r}.npy"), (local[0] + local[1]) / 2, rtol=1e-6, atol=1e-7)
    assert np.abs(local[1]).sum() > 0 and (local[1] == 0).sum() > (local[0] == 0).sum()


def test_allreduce_mean_single_process_is_identity() -> None:
    from ksim_b200.distributed import allreduce_mean_

    g = [torch.arange(6.0).reshape(2, 3)]
    allreduce_mean_(g)
    assert torch.equal(g[0], torch.arange(6.0).reshape(2, 3))


def test_shard_range_edge_cases() -> None:
    from ksim_b200.distributed import shard_range

    assert shard_range(0, 0, 4) == (0, 0)
    assert [shard_range(3, r, 8) for r in range(8)] == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    covered = [i for r in range(8) for i in range(*shard_range(32768, r, 8))]
    assert covered == list(range(32768))
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)
