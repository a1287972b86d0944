"""Multi-GPU partitioning: one process per GPU, `torch.distributed` for the plumbing.

The likelihood path shards in two ways, neither with a collective on the data path:
  * batch prediction — patients are block-sharded, the cluster models are replicated (rebuilt on every rank
    from (x, y, theta)); one all_gather of the P x (K+1) probability matrix at the end;
  * alpha / seed sweeps — independent Gibbs chains are dealt round-robin to the ranks, each with its own
    engine, stream and RandomState; one gather of the per-chain results (z, Nk, log-likelihood traces).
A single chain does not shard (strictly sequential patient order): "replicas only".
Works with backend "nccl" (GPU tensors) and "gloo" (CPU tensors; used by the CPU tests).
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous balanced block of range(n) owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank {} outside world of size {}".format(rank, world))
    return (n * rank) // world, (n * (rank + 1)) // world


def chains_for_rank(n_chains, world, rank):
    """Round-robin deal of chain ids 0..n_chains-1."""
    return list(range(rank, n_chains, world))


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _device_for_backend():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_rows(local, n_total):
    """Concatenate row blocks (shard_bounds order) of a 2-D float64 array across ranks; every rank gets
    the full n_total x C matrix.  Single-process: returns `local`."""
    import torch
    dist = _dist()
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    cols = local.shape[1]
    cap = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    dev = _device_for_backend()
    buf = torch.zeros((cap, cols), dtype=torch.float64, device=dev)
    buf[:local.shape[0]] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        parts.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(parts, axis=0)


def predict_sharded(score_fn, X_new, Y_new, a_draws):
    """Batch prediction with patients sharded across ranks.  `score_fn(X, Y, a)` is the per-rank scorer
    (`MoGP_constrained.predict_batch` of a replicated model); returns the full probability matrix."""
    dist = _dist()
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    P = X_new.shape[0]
    lo, hi = shard_bounds(P, world, rank)
    local = score_fn(X_new[lo:hi], Y_new[lo:hi], np.asarray(a_draws)[lo:hi]) if hi > lo else None
    # a rank whose shard is empty (fewer patients than ranks) contributes a 0-row block: every rank takes the same path
    # through the collectives, the column count comes from the ranks that scored something
    cols = local.shape[1] if local is not None else 0
    if dist is not None and world > 1:
        import torch
        t = torch.tensor([cols], dtype=torch.int64, device=_device_for_backend())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cols = int(t.item())
    if local is None:
        local = np.zeros((0, cols))
    return all_gather_rows(local, P)


def gather_chain_results(local_results, n_chains):
    """local_results: {chain_id: dict(z=int array, ll=float array, ...)} of this rank's chains.
    Returns the list of all chains' results, ordered by chain id, on every rank."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return [local_results[c] for c in range(n_chains)]
    world = dist.get_world_size()
    gathered = [None] * world
    dist.all_gather_object(gathered, local_results)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[c] for c in range(n_chains)]
