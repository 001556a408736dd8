"""Host-side multi-GPU logic on CPU: probe sharding + the single all-reduce of the accumulators,
exercised with world_size 2 over gloo.  The operator is a CPU stand-in (a dense symmetric matrix) — the sharding
layer treats the matvec as opaque, exactly as the estimators treat `Xfun`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lip_b200  # noqa: F401
from lip_b200 import _dist


def test_probe_slice_partitions_every_row_once():
    for B in (0, 1, 2, 5, 64, 257):
        for ws in (1, 2, 3, 8):
            rows = []
            sizes = []
            for r in range(ws):
                s = _dist.probe_slice(B, r, ws)
                rows += list(range(B))[s]
                sizes.append(s.stop - s.start)
            assert rows == list(range(B))
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        _dist.probe_slice(4, 2, 2)


def test_single_process_is_a_plain_mean():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(12, 12, generator=g)
    A = A @ A.T
    E = torch.randint(0, 2, (10, 12), generator=g).float() * 2 - 1
    got = _dist.hutchinson_sharded(lambda v: A @ v, E)
    ref = torch.stack([e @ (A @ e) for e in E]).mean()
    assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    r, w, _ = _dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, ws) and _dist.world() == (rank, ws)
    g = torch.Generator().manual_seed(1)               # identical inputs on every rank (probes are inputs)
    A = torch.randn(40, 40, generator=g)
    A = A @ A.T + 40 * torch.eye(40)
    E = torch.randint(0, 2, (B, 40), generator=g).float() * 2 - 1
    calls = []

    def Xfun(V):                                        # batched closure: sees only this rank's rows
        calls.append(V.shape[0])
        return V @ A
    Xfun._lip_batched = True
    tr = _dist.hutchinson_sharded(Xfun, E)

    def integrand(matvec, V):                           # an SLQ-like per-probe scalar: v . log-ish quadratic form
        return (V * matvec(V)).sum(1).log()
    integrand._lip_batched = True
    lq = _dist.slq_sharded(integrand, Xfun, E)
    out[rank] = (float(tr), float(lq), sum(calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 1])
def test_world_size_2_gloo_matches_single_process(B):
    ws = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(ws, port, B, out), nprocs=ws, join=True)
        res = dict(out)
    g = torch.Generator().manual_seed(1)
    A = torch.randn(40, 40, generator=g)
    A = A @ A.T + 40 * torch.eye(40)
    E = torch.randint(0, 2, (B, 40), generator=g).float() * 2 - 1
    quad = (E * (E @ A)).sum(1)
    ref_tr, ref_lq = float(quad.mean()), float(quad.log().mean())
    for r in range(ws):
        tr, lq, _ = res[r]
        assert abs(tr - ref_tr) <= 1e-5 * abs(ref_tr)
        assert abs(lq - ref_lq) <= 1e-5 * abs(ref_lq)
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]       # every rank holds the same global estimate
    # each estimator call pushes a rank's own rows once: 2 calls x rows owned
    owned = [(_dist.probe_slice(B, r, ws).stop - _dist.probe_slice(B, r, ws).start) for r in range(ws)]
    assert [res[r][2] for r in range(ws)] == [2 * o for o in owned]


def _worker_points(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    _dist.init_from_env(backend="gloo")
    g = torch.Generator().manual_seed(3)
    J = torch.randn(9, 4, 30, generator=g)                 # per-point Jacobians [M, K, D] (identical on every rank)
    V = torch.randn(5, 30, generator=g)
    sl = _dist.point_slice(J.shape[0])
    Jl = J[sl]

    def local(v):                                          # this rank's partial GGN-vector product: sum_i J_i^T J_i v
        return torch.einsum("mkd,mke,be->bd", Jl, Jl, v)
    local._lip_batched = True
    res = _dist.point_sharded(local)(V)
    out[rank] = res.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_operator_sums_partial_products():
    """GGN(Z) v = sum over ranks of GGN(Z_rank) v, one all-reduce of the [B, D] block."""
    ws, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_points, args=(ws, port, out), nprocs=ws, join=True)
        res = dict(out)
    g = torch.Generator().manual_seed(3)
    J = torch.randn(9, 4, 30, generator=g)
    V = torch.randn(5, 30, generator=g)
    ref = torch.einsum("mkd,mke,be->bd", J, J, V).numpy()
    for r in range(ws):
        np.testing.assert_allclose(res[r], ref, rtol=1e-5, atol=1e-5)


def _worker_zgrad(rank, ws, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    _dist.init_from_env(backend="gloo")
    g = torch.Generator().manual_seed(5)
    T = torch.randn(6, 3, 20, 20, generator=g)             # a bilinear "Z-gradient" tensor: dZ[m, i] = sum_b u_b^T T[m, i] v_b
    U = torch.randn(B, 20, generator=g)
    V = torch.randn(B, 20, generator=g)
    seen = []

    def zgrad(u, v):                                       # stands in for closure.zgrad: sums over the probes it is given
        seen.append(u.shape[0])
        return torch.einsum("bd,mide,be->mi", u, T, v)
    res = _dist.zgrad_sharded(zgrad, U, V)
    out[rank] = (res.numpy().copy(), sum(seen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 1])
def test_zgrad_sharded_sums_probe_slices(B):
    """the Z-gradient of a probe-averaged quantity shards over probes with one all-reduce of [M, in]
    (B = 1: one rank owns no probe and still joins the reduction)."""
    ws, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_zgrad, args=(ws, port, B, out), nprocs=ws, join=True)
        res = dict(out)
    g = torch.Generator().manual_seed(5)
    T = torch.randn(6, 3, 20, 20, generator=g)
    U = torch.randn(B, 20, generator=g)
    V = torch.randn(B, 20, generator=g)
    ref = torch.einsum("bd,mide,be->mi", U, T, V).numpy()
    for r in range(ws):
        np.testing.assert_allclose(res[r][0], ref, rtol=1e-5, atol=1e-5)
    assert sum(res[r][1] for r in range(ws)) == B          # every probe pair is pushed exactly once


# ---------------------------------------------------------------------------------------------- hybrid probe x basis sharding
def test_group_layout_prefers_probe_groups():
    assert _dist.group_layout(1, 4) == (1, 1)
    assert _dist.group_layout(2, 4) == (2, 1)
    assert _dist.group_layout(4, 4) == (4, 1)
    assert _dist.group_layout(8, 4) == (4, 2)          # BASELINE's SLQ case: 4 probes on 8 GPUs -> pairs share a probe's bases
    assert _dist.group_layout(8, 1) == (1, 8)
    assert _dist.group_layout(8, 3) == (2, 4)
    assert _dist.group_layout(8, 16) == (8, 1)
    assert _dist.group_layout(8, 4, probes_per_group=2) == (2, 4)
    for ws in range(1, 17):
        for B in range(1, 9):
            P, S = _dist.group_layout(ws, B)
            assert P * S == ws and 1 <= P <= B
    with pytest.raises(ValueError):
        _dist.group_layout(0, 4)


def _hybrid_worker(rank, ws, port, B, force_S, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    _dist.init_from_env(backend="gloo")
    from lip_b200 import matfree
    g = torch.Generator().manual_seed(3)
    probes = torch.randn(B, 17, generator=g)
    seen = {}

    def fake_quadrature(matvec, mine, k, *, form, fn, clip_min, comm, model=None):     # CPU stand-in for the native recurrence
        seen["rows"], seen["comm"] = mine.clone(), comm
        return (mine ** 2).sum(1) * k

    def fake_comms(S, slot=0):
        return None if S == 1 else _dist.NativeComm(None, S, rank % S)

    matfree.slq_quadrature, _dist.native_comms = fake_quadrature, fake_comms
    if force_S:
        _dist.group_layout = lambda w, b: (w // force_S, force_S)
    got = float(_dist.slq_logdet_hybrid(None, probes, 5))
    ref = float(((probes ** 2).sum(1) * 5).mean())
    P, S = _dist.group_layout(ws, B)
    ok = abs(got - ref) <= 1e-5 * abs(ref)
    ok = ok and torch.equal(seen["rows"], probes[_dist.probe_slice(B, rank // S, P)])
    ok = ok and ((seen["comm"] is None) == (S == 1)) and (S == 1 or seen["comm"].rank == rank % S)
    out[rank] = 1.0 if ok else 0.0
    dist.destroy_process_group()


@pytest.mark.parametrize("B,force_S", [(4, 0), (1, 0), (3, 2)])
def test_hybrid_slq_layout_world_size_2_gloo(B, force_S):
    """world_size 2: (B=4 -> two probe groups, no basis sharding), (B=1 -> one group of two ranks sharing the probe's
    bases), (B=3 forced into one 2-rank group): every probe is evaluated by exactly one group, group members see the same rows, only the
    group leader contributes to the all-reduced mean."""
    ws = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_hybrid_worker, args=(ws, _free_port(), B, force_S, out), nprocs=ws, join=True)
    assert [out[r] for r in range(ws)] == [1.0, 1.0]
