"""Multi-GPU layout of the hot path: config 5 = N independent 4K streams, one per GPU / rank, no
data-path collective (the path shards by stream).  The only cross-rank exchange is the measurement itself:
max-over-ranks of the device time, from which the whole-job frames/s follows.  Used by bench.py (NCCL) and by the
world_size-2 gloo test (tests/test_cpu_multi_rank.py)."""
import torch
import torch.distributed as dist


def stream_seed(rank):
    """Seed of the synthetic command-record stream a rank decodes (independent content per stream)."""
    return 1000 + 100 * rank


def reference_seed(rank):
    return 7 + rank


def max_over_ranks(ms_local, device="cpu"):
    """Device time of the slowest rank (all ranks get the same value)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local)
    t = torch.tensor([float(ms_local)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_frames_per_second(frames_per_rank, ms_local, device="cpu"):
    """Whole-job throughput: every rank processed frames_per_rank pictures; the job took as long as the slowest rank."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    ms = max_over_ranks(ms_local, device)
    return frames_per_rank * world / (ms / 1000.0), ms
