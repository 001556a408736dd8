"""Probe sharding across the GPUs of one box (the reference is single-device, src/data.py:90-93).

Every estimator on the hot path is a mean / sum over independent probe columns (stochtrace.py:19,34; matfree's
`estimator` mean; sample.py:155 independent samples).  One process per GPU: weights, points and the activation
cache are replicated, rank r owns a contiguous slice of the probe rows and runs its own Krylov recurrences; the
only communication is ONE all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of the partial
accumulators — a few floats per estimator call.  Nothing here touches the data path.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) of the initialised default process group, (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world_size, local_rank).  A single-process run (WORLD_SIZE unset or 1) needs no group."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, ws, local


def probe_slice(num_probes: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> slice:
    """Contiguous, balanced slice of the probe rows owned by `rank` (sizes differ by at most one; the first
    num_probes % world_size ranks get the extra row; a rank may own zero rows when num_probes < world_size)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    if num_probes < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad shard request: num_probes={num_probes} rank={rank} world_size={world_size}")
    base, extra = divmod(num_probes, world_size)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def sharded_mean(per_probe_fn: Callable[[torch.Tensor], torch.Tensor], probes: torch.Tensor) -> torch.Tensor:
    """mean_b per_probe_fn(probes)[b] over ALL probe rows, each rank evaluating only its slice.
    `probes` is the full [B, n] probe matrix (identical on every rank: probes are inputs);
    `per_probe_fn` maps a [b_loc, n] block to b_loc per-probe values.  Returns the global mean on every rank."""
    B = probes.shape[0]
    mine = probes[probe_slice(B)]
    acc = torch.zeros(1, dtype=torch.float64, device=probes.device)
    if mine.shape[0] > 0:
        acc += per_probe_fn(mine).double().sum()
    allreduce_sum_(acc)
    return (acc / B).to(torch.float32)[0]


def hutchinson_sharded(Xfun: Callable, probes: torch.Tensor) -> torch.Tensor:
    """stochastic_trace_estimator_mvp (stochtrace.py:22-34) with the probe rows sharded over ranks."""
    def quad(E):
        Y = Xfun(E) if getattr(Xfun, "_lip_batched", False) else torch.stack([Xfun(e) for e in E])
        return (E * Y.reshape(E.shape)).sum(1)
    return sharded_mean(quad, probes)


def slq_sharded(integrand: Callable, matvec, probes: torch.Tensor, *parameters) -> torch.Tensor:
    """matfree.stochtrace.estimator(integrand, sampler)(matvec, key) with the sampler's rows sharded over ranks:
    the mean over all probes of integrand(matvec, probe) (train_inducing.py:156-163)."""
    def quad(E):
        if getattr(integrand, "_lip_batched", False):
            return integrand(matvec, E, *parameters).reshape(-1)
        return torch.stack([integrand(matvec, e, *parameters) for e in E]).reshape(-1)
    return sharded_mean(quad, probes)


def zgrad_sharded(zgrad_fn: Callable, cotangents: torch.Tensor, vectors: torch.Tensor) -> torch.Tensor:
    """Probe-sharded gradient with respect to Z (SURVEY §8 rows e / f1): sum_b d/dZ <cotangents[b], operator(vectors[b])> with each
    rank pushing only its slice of the probe pairs through `zgrad_fn` (a closure's .zgrad: lip_zgrad summed over its probes) and ONE
    all-reduce of the [M, in] result — the Z-gradient is tiny next to the [B, D] probe blocks, so nothing else crosses NVLink.
    Both arguments are the full [B, ...] blocks, identical on every rank; returns the global sum on every rank."""
    B = cotangents.shape[0]
    if vectors.shape[0] != B:
        raise ValueError(f"zgrad_sharded: {B} cotangents against {vectors.shape[0]} vectors")
    sl = probe_slice(B)
    out = None
    if sl.stop > sl.start:
        out = zgrad_fn(cotangents[sl], vectors[sl]).contiguous()
    # a rank that owns no probe still takes part in the all-reduce; it learns the shape from a size exchange
    shape = torch.zeros(8, dtype=torch.int64, device=cotangents.device)
    if out is not None:
        shape[0] = out.dim()
        shape[1:1 + out.dim()] = torch.tensor(out.shape, dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(shape, op=dist.ReduceOp.MAX)
    if out is None:
        nd = int(shape[0])
        if nd == 0:
            raise ValueError("zgrad_sharded: no rank owns a probe")
        out = torch.zeros(tuple(int(v) for v in shape[1:1 + nd]), dtype=torch.float32, device=cotangents.device)
    return allreduce_sum_(out)


# ---------------------------------------------------------------------------------------------- Krylov-basis (D) sharding
def group_layout(world_size: int, num_probes: int, probes_per_group: int = 1) -> Tuple[int, int]:
    """(P, S): P probe groups of S ranks each, P * S = world_size.  Inside a group the S ranks share every Krylov vector of the
    group's probes column-wise (lip_slq_quadrature_sharded); across groups nothing is communicated.  Probes are the free axis, so P is
    the largest divisor of world_size that leaves every group at least `probes_per_group` probes: 4 probes on 1 / 2 / 4 / 8 GPUs ->
    (1, 1), (2, 1), (4, 1), (4, 2).  (probes_per_group = 2 with pipelines = 2 in slq_logdet_hybrid runs two recurrences per GPU on two
    streams; measured on 2 x B200 it does not pay — 1.17 s against 0.92 s for the k = 409 logdet: the tcgen05 GEMM CTAs of one
    pipeline cannot share an SM with the basis kernels of the other (shared memory), and two host threads contend for the launch path —
    profiles/r02_dist_slq_2gpu.txt.)"""
    if world_size < 1 or num_probes < 1:
        raise ValueError(f"bad layout request: world_size={world_size} num_probes={num_probes}")
    want = max(1, num_probes // max(1, probes_per_group))
    P = max(p for p in range(1, world_size + 1) if world_size % p == 0 and p <= want)
    return P, world_size // P


class NativeComm:
    """A library-owned NCCL communicator (lip_comm) over a contiguous block of ranks of the default process group."""

    def __init__(self, handle, world_size, rank):
        self.handle, self.world, self.rank = handle, world_size, rank


_NATIVE_COMMS = {}


def native_comms(shard_size: int, slot: int = 0) -> Optional[NativeComm]:
    """Collective over the default group: partitions the ranks into consecutive blocks of `shard_size` and returns this rank's
    lip_comm (None when shard_size == 1).  The leaders' NCCL unique ids travel in one all-gather of 128 bytes per rank.
    `slot` distinguishes independent communicators over the same ranks (one per concurrent stream: operations on ONE NCCL
    communicator must be enqueued in the same order on every rank, which two host threads cannot promise each other)."""
    rank, ws = world()
    if shard_size <= 1 or ws == 1:
        return None
    if ws % shard_size:
        raise ValueError(f"native_comms: shard_size={shard_size} does not divide world_size={ws}")
    if (shard_size, slot) in _NATIVE_COMMS:
        return _NATIVE_COMMS[(shard_size, slot)]
    import ctypes
    from . import _cabi as cabi
    L = cabi.lib()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    leader = rank // shard_size * shard_size
    buf = (ctypes.c_uint8 * 128)()
    if rank == leader:
        cabi.check(L.lip_comm_unique_id(buf), "lip_comm_unique_id")
    mine = torch.tensor(list(buf), dtype=torch.uint8).to(dev)
    allids = torch.empty(ws * 128, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allids, mine)
    idb = (ctypes.c_uint8 * 128)(*allids[leader * 128:(leader + 1) * 128].cpu().tolist())
    h = ctypes.c_void_p()
    cabi.check(L.lip_comm_create(idb, shard_size, rank - leader, ctypes.byref(h)), "lip_comm_create")
    comm = NativeComm(h, shard_size, rank - leader)
    _NATIVE_COMMS[(shard_size, slot)] = comm
    return comm


_PIPELINE_MODELS = {}


def _second_model(bm):
    """BoundModel.clone() of `bm`, cached for as long as `bm` lives (weak on the original)."""
    import weakref
    hit = _PIPELINE_MODELS.get(id(bm))
    if hit is not None and hit[0]() is bm:
        return hit[1]
    clone = bm.clone()
    _PIPELINE_MODELS[id(bm)] = (weakref.ref(bm), clone)
    if len(_PIPELINE_MODELS) > 4:
        _PIPELINE_MODELS.pop(next(iter(_PIPELINE_MODELS)))
    return clone


def slq_logdet_hybrid(matvec, probes: torch.Tensor, num_matvecs: int, *, form="gkl", fn="log", clip_min=None,
                      pipelines: int = 1) -> torch.Tensor:
    """mean_b |v_b|^2 e1^T f(T_b) e1 over ALL probe rows (matfree.stochtrace.estimator of an SLQ integrand, train_inducing.py:156-163)
    on every GPU of the job.  Layout (group_layout): probes over groups of ranks; inside a group the Krylov bases are cut column-wise.
    pipelines = 2 (experimental, off by default: see group_layout) runs a rank's probes as two concurrent recurrences — two host
    threads, two CUDA streams, two model handles (BoundModel.clone), two communicators.  `probes` is the full [B, n] matrix, identical on every rank.  One all-reduce of a float64
    accumulator ends the call."""
    import threading
    from . import matfree
    rank, ws = world()
    B = probes.shape[0]
    P, S = group_layout(ws, B)
    g = rank // S
    mine = probes[probe_slice(B, g, P)]
    nmine = int(mine.shape[0])
    npipe = max(1, min(int(pipelines), nmine)) if getattr(matvec, "_lip_model", None) is not None else 1
    comms = [native_comms(S, slot) for slot in range(npipe)]             # collective: every rank creates the same set, in order
    acc = torch.zeros(1, dtype=torch.float64, device=probes.device)
    if nmine > 0:
        if npipe == 1:
            q = matfree.slq_quadrature(matvec, mine, num_matvecs, form=form, fn=fn, clip_min=clip_min, comm=comms[0])
        else:
            dev = probes.device
            main = torch.cuda.current_stream(dev)
            parts = [mine[probe_slice(nmine, i, npipe)] for i in range(npipe)]
            models = [None] + [_second_model(matvec._lip_model) for _ in range(npipe - 1)]
            if npipe > 2:
                models = [None] + [matvec._lip_model.clone() for _ in range(npipe - 1)]
            streams = [torch.cuda.Stream(device=dev) for _ in range(npipe)]
            out, errs = [None] * npipe, [None] * npipe

            def run(i):
                try:
                    torch.cuda.set_device(dev)                              # the CUDA current device is per host thread
                    with torch.cuda.stream(streams[i]):
                        streams[i].wait_stream(main)
                        out[i] = matfree.slq_quadrature(matvec, parts[i], num_matvecs, form=form, fn=fn, clip_min=clip_min,
                                                        comm=comms[i], model=models[i])
                except BaseException as e:                                   # noqa: BLE001
                    errs[i] = e

            threads = [threading.Thread(target=run, args=(i,)) for i in range(npipe)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            for e in errs:
                if e is not None:
                    raise e
            for s_ in streams:
                main.wait_stream(s_)
            q = torch.cat(out)
        if rank % S == 0:                       # every rank of a group holds the same values: the leader contributes them
            acc += q.double().sum()
    allreduce_sum_(acc)
    return (acc / B).to(torch.float32)[0]


# ---------------------------------------------------------------------------------------------- point (M) sharding
def point_slice(num_points: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> slice:
    """Contiguous, balanced slice of the inducing points owned by `rank` (sharding over M for large
    point sets).  Same partition rule as probe_slice."""
    return probe_slice(num_points, rank, world_size)


def point_sharded(local_fn: Callable[[torch.Tensor], torch.Tensor]) -> Callable[[torch.Tensor], torch.Tensor]:
    """GGN(Z) v = sum_g GGN(Z_g) v: wraps a rank-local operator (bound to this rank's points, already scaled with the
    GLOBAL recalibration N / M) into the global one with ONE all-reduce of the [B, D] result per application.
    Every rank must call the wrapper with the same `v`."""

    def global_fn(v):
        out = local_fn(v)
        return allreduce_sum_(out.contiguous())

    for attr in ("_lip_batched", "_lip_model", "_lip_kind"):
        if hasattr(local_fn, attr):
            setattr(global_fn, attr, getattr(local_fn, attr))
    global_fn._lip_native = False      # needs the all-reduce: the Krylov routines must call it back, not run the local model alone
    return global_fn
