"""Multi-GPU plumbing for the one place the path shards: samples are independent because
all randomness is keyed by (pixel, frame_index, bounce) (sample_secondary_ray.hlsl:52), so GPU `rank`
of `world` renders a contiguous block of samples of every pixel into its own FP32 sum buffer
(scene + BVH replicated), and ONE reduce of the W x H x 4 sum buffers to rank 0 per batch of frames is
the only exchange. torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def first_frame(steps: int, rank: int, world: int, first: int = 0) -> int:
    """Rank `rank` renders the contiguous block of `steps` frames starting here (one bpt_render call,
    so the library can keep several samples in flight per wave)."""
    return first + rank * steps


def split_frames(total_frames: int, rank: int, world: int, first: int = 0) -> tuple[int, int]:
    """Strong scaling: ONE job of `total_frames` samples split over the ranks in contiguous blocks — (first frame, count) of `rank`; the
    first `total_frames % world` ranks take one frame more. The blocks tile [first, first + total_frames) exactly, so the reduced sum
    buffer is the single-GPU buffer of the same job (up to FP32 association across ranks)."""
    base, extra = divmod(total_frames, world)
    return first + rank * base + min(rank, extra), base + (1 if rank < extra else 0)


def wave_schedule(frames: int, max_wave: int) -> list[int]:
    """Sample waves of a FINITE frame-at-a-time job (PathTracingPass with prefetch): the frames of a wave become available together when
    its last bounce ends, and their images then leave over PCIe while the next wave is traced — only the LAST wave's read-back is exposed.
    So the waves shrink towards the end: each is 4/5 of what is left (at most `max_wave`, the library's wave size at this resolution).
    Every frame belongs to exactly one wave; a one-frame job is one wave of one."""
    out, left = [], int(frames)
    while left > 0:
        out.append(min(int(max_wave), max(1, -(-4 * left // 5))))
        left -= out[-1]
    return out


def frame_index(step: int, steps: int, rank: int, world: int, first: int = 0) -> int:
    """frame_index of the `step`-th frame this rank renders."""
    return first_frame(steps, rank, world, first) + step


def frames_of_rank(steps: int, rank: int, world: int, first: int = 0) -> list[int]:
    return [frame_index(k, steps, rank, world, first) for k in range(steps)]


def reduce_sums(sum_buffer: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """In-place sum of every rank's FP32 accumulation buffer into rank `dst` (no-op for world 1)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sum_buffer, dst=dst, op=dist.ReduceOp.SUM)
    return sum_buffer


def resolve(sum_buffer: torch.Tensor, total_samples: int) -> torch.Tensor:
    """image = sum / N with alpha 1 (pt_accumulate.hlsl:3-11 restated as sum-then-divide so that the
    result does not depend on how samples were distributed over GPUs)."""
    out = sum_buffer * (1.0 / float(total_samples))
    out[..., 3] = 1.0
    return out


def comm_init(ctx, device: torch.device | None = None) -> None:
    """Gives `ctx` (a capi.Context of the CUDA library) its own NCCL communicator over the ranks of the current
    torch.distributed group: rank 0 draws the unique id (bpt_comm_unique_id), torch.distributed only carries its 128 bytes.
    Afterwards ctx.reduce(root) is the one exchange step of SURVEY §8e, issued by the library on its own stream."""
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(ctx.L.comm_unique_id()), dtype=torch.uint8).clone()
    if device is not None and dist.get_backend() == "nccl":
        uid = uid.to(device)
    dist.broadcast(uid, src=0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)


# ---- DDGI update (last sentence): shard by probe index, all-gather the per-ray results ---------------------------
def probe_range(num_probes: int, rank: int, world: int) -> tuple[int, int]:
    """(first_probe, count) of rank `rank`: contiguous blocks, the first `num_probes % world` ranks get one probe more."""
    base, extra = divmod(num_probes, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def allgather_probe_rays(local_rays: torch.Tensor, num_probes: int, rays_per_probe: int) -> torch.Tensor:
    """The one exchange step of a sharded DDGI update: every rank contributes the (count * rays_per_probe, 4) float32 results of its
    probe_range and receives the full (num_probes * rays_per_probe, 4) array in probe order — the input of bpt_blend_probes, which
    every rank then runs for all probes (blending is ~1 % of the update). Uneven ranges are padded to the largest block for the
    collective (NCCL / gloo all_gather needs equal sizes) and trimmed afterwards. No-op for world 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rays
    world = dist.get_world_size()
    counts = [probe_range(num_probes, r, world)[1] for r in range(world)]
    rows = max(counts) * rays_per_probe
    padded = torch.zeros((rows, 4), dtype=local_rays.dtype, device=local_rays.device)
    padded[: local_rays.shape[0]] = local_rays
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[: c * rays_per_probe] for p, c in zip(parts, counts)], dim=0)
