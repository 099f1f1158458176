"""Multi-GPU plumbing: independent proof instances shard one per rank with no data-path
collective; the only exchange is the final gather of the (small) transcripts on rank 0.  Backend-agnostic:
NCCL over NVLink on the GPU box, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def instance_seed(base_seed, rank):
    """Deterministic RNG seed of the proof instance owned by `rank` (BASELINE config 5: seeds s, s+1, ...)."""
    return (base_seed + rank) & 0xFFFFFFFFFFFFFFFF


def shard_instances(n_instances, rank, world):
    """Round-robin assignment of independent instances (proofs / oracles / codewords) to ranks."""
    return list(range(rank, n_instances, world))


def gather_transcripts(local: bytes, device=None):
    """All ranks contribute a byte string; every rank returns the list of all of them (rank order).
    Variable lengths are handled by a length all-gather followed by a padded payload all-gather."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    lens = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(lens, n)
    mx = int(max(int(t.item()) for t in lens))
    buf = torch.zeros(max(mx, 1), dtype=torch.uint8, device=dev)
    if local:
        buf[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    return [bytes(b[:int(l.item())].cpu().numpy().tobytes()) for b, l in zip(bufs, lens)]


def max_over_ranks(value: float, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
