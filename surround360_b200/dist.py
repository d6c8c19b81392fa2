"""Multi-GPU plumbing: frames are independent units (prev_frame_data_dir NONE), so rank r of W
renders frames r, r+W, r+2W, ... and the only collective is one broadcast of the serialised
rig at start-up (torch.distributed; NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def frames_for_rank(n_frames, rank, world):
    """frame f -> rank f mod world (SURVEY 8e / BASELINE config 4)."""
    return list(range(rank, n_frames, world))


def chain_ranges(n_frames, world, chains_per_rank):
    """Video mode (prev_frame_data_dir != NONE): frame i needs frame i-1, so the sequence is cut into world x chains_per_rank
    contiguous ranges (SURVEY 8e option A; scripts/batch_process_video.py:137-158 renders one such range per invocation). Returns
    ranges[rank][chain] = (first, last_exclusive). The first frame of every range starts without history, like the first frame of a
    video: the output deviates from a single-chain render at those world x chains_per_rank - 1 range heads (and converges again
    within a few frames, the regularisation being a per-frame blend towards the previous flow)."""
    total = world * chains_per_rank
    base, extra = divmod(n_frames, total)
    out, start = [], 0
    for r in range(world):
        row = []
        for c in range(chains_per_rank):
            n = base + (1 if r * chains_per_rank + c < extra else 0)
            row.append((start, start + n))
            start += n
        out.append(row)
    return out


def broadcast_bytes(buf, src=0, device=None):
    """Broadcast a uint8 numpy buffer from `src` to every rank; returns the received array."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.ascontiguousarray(buf, np.uint8)
    dev = device if device is not None else torch.device("cpu")
    rank = dist.get_rank()
    n = torch.tensor([len(buf) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(buf, np.uint8)).to(dev)
    else:
        t = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return t.cpu().numpy()


def gather_frame_counts(n_local, device=None):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [n_local]
    dev = device if device is not None else torch.device("cpu")
    t = torch.tensor([n_local], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
