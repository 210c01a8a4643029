"""world_size-2 gloo test of the multi-rank host logic (no GPU): agent->rank blocks, the all-gather of theta/psi
rows through dqgp_b200.exchange_rows, and a replicated consensus that is identical on every rank and equal to the
single-process reference trajectory (oracle local updates driven by synthetic gradients)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _trajectory(A, P, iters, seed=3):
    """Single-process reference: main.py:2513-2555 with gradients replaced by a deterministic function of (agent, z)."""
    from oracle import agent_step, torus
    rs = np.random.RandomState(seed)
    theta, psi = np.round(rs.rand(A, P), 4), np.round(rs.rand(A, P), 4)
    zs = []
    for _ in range(iters):
        z = np.round(torus.update_z(theta, psi, 100.0), 4)
        zs.append(z)
        for a in range(A):
            grad4 = np.round(50.0 * np.sin(3.0 * z + a), 4)
            theta[a], psi[a] = agent_step.local_update(torus.wrap(z), grad4, psi[a], 100.0, 100.0)
    return np.array(zs), theta, psi


def _worker(rank, world, port, A, P, iters, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dqgp_b200 as d
    from oracle import agent_step, torus
    rs = np.random.RandomState(3)
    # theta_a, psi_a side by side in one (A, 2, P) buffer, as AdmmEngine keeps them: ONE all-gather per iteration
    rows = torch.from_numpy(np.stack([np.round(rs.rand(A, P), 4), np.round(rs.rand(A, P), 4)], axis=1).copy())
    theta, psi = rows[:, 0], rows[:, 1]
    block = d.agent_block(rank, world, A)
    _, _, admm = d.create_riemannian_framework(P, rho=100.0)
    zs = []
    for _ in range(iters):
        z = np.round(admm.update_z(theta.numpy(), psi.numpy()), 4)          # replicated consensus, agent order
        zs.append(z)
        loc = torch.empty((len(block), 2, P), dtype=torch.float64)
        for i, a in enumerate(block):
            grad4 = np.round(50.0 * np.sin(3.0 * z + a), 4)
            t, p = agent_step.local_update(torus.wrap(z), grad4, psi[a].numpy(), 100.0, 100.0)
            loc[i, 0], loc[i, 1] = torch.from_numpy(t), torch.from_numpy(p)
        d.exchange_rows(rows, loc, None, world)
    out[rank] = (np.array(zs), theta.numpy().copy(), psi.numpy().copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("A,P", [(4, 12), (8, 48)])
def test_two_rank_consensus_matches_single_process(A, P):
    iters, world = 3, 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, A, P, iters, out), nprocs=world, join=True)
    zs_ref, th_ref, ps_ref = _trajectory(A, P, iters)
    for r in range(world):
        zs, th, ps = out[r]
        assert np.array_equal(zs, zs_ref), f"rank {r}: consensus trajectory differs"
        assert np.array_equal(th, th_ref) and np.array_equal(ps, ps_ref)


def test_agent_block_partition():
    import dqgp_b200 as d
    assert list(d.agent_block(0, 1, 4)) == [0, 1, 2, 3]
    assert [list(d.agent_block(r, 4, 8)) for r in range(4)] == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert list(d.agent_block(7, 8, 16)) == [14, 15]
    with pytest.raises(ValueError):
        d.agent_block(0, 3, 8)


def _fake_predict(Xtr, Ytr, Xva, *args, **kw):
    """Stand-in for the GPU prediction (this test is about fold ownership and the gather): ridge fit on the host."""
    w = np.linalg.solve(Xtr.T @ Xtr + 1e-3 * np.eye(Xtr.shape[1]), Xtr.T @ Ytr)
    mean = Xva @ w
    return mean, np.full(len(Xva), 0.05 + 0.001 * len(Xtr) / 100.0), None, None, None


def _cv_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dqgp_b200 import predict
    rng = np.random.default_rng(5)
    X = rng.standard_normal((203, 3)); Y = X @ np.array([0.5, -1.0, 2.0]) + 0.1 * rng.standard_normal(203)
    out[rank] = predict.k_fold_cross_validation_consensus(X, Y, None, 3, 1, 0.1, k_folds=5, rank=rank, world_size=world,
                                                          _predict=_fake_predict)
    dist.destroy_process_group()


def test_cv_folds_sharded_over_two_ranks_equal_single_process():
    from dqgp_b200 import predict
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cv_worker, args=(2, port, out), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((203, 3)); Y = X @ np.array([0.5, -1.0, 2.0]) + 0.1 * rng.standard_normal(203)
    ref = predict.k_fold_cross_validation_consensus(X, Y, None, 3, 1, 0.1, k_folds=5, _predict=_fake_predict)
    assert ref["valid_folds"] == 5
    for r in range(2):
        assert out[r] == ref
