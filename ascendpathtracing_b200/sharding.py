"""Stripe partition of a frame across ranks and assembly of the gathered stripes.

The reference splits the flat, x-major ray array into 8 contiguous slices = vertical stripes
(src/render.cpp:24-27, scripts/gen_data.py:32); rank r of G renders columns [x0, x1) and the only exchange of
the whole path is one gather of the resolved 8-bit stripes at the end."""


def stripe(width, rank, world):
    """Columns [x0, x1) of rank `rank`: contiguous, balanced to within one column, covering [0, width)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(width, world)
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


def path_slice(width, height, samples, rank, world):
    """[first, first+count) of the N-long path arrays owned by `rank` (contiguous because paths are x-major)."""
    x0, x1 = stripe(width, rank, world)
    per_col = height * 4 * samples
    return x0 * per_col, (x1 - x0) * per_col


def interleaved_stripes(width, rank, world, parts):
    """`parts` column ranges per rank, dealt round-robin across the frame: piece j of the width/(world*parts) partition goes
    to rank j % world.  One wide stripe per rank leaves the ranks unevenly loaded when path length varies across the image
    (materials, Russian roulette, deep mirror paths: 80 % strong-scaling efficiency on 8 GPUs for the 10 k-sphere scene);
    dealing narrower stripes evens that out with no communication.  Returns [(x0, x1), ...] in ascending x (possibly empty
    ranges when the frame has fewer columns than pieces)."""
    if parts < 1:
        raise ValueError("parts must be >= 1")
    return [stripe(width, j, world * parts) for j in range(rank, world * parts, world)]


def gather_interleaved(local, width, parts, group=None):
    """All-gather of per-rank buffers [H, sum of the rank's piece widths, 3] (pieces side by side, ascending x) into the
    full [H, W, 3] frame: the counterpart of interleaved_stripes()."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    h = local.shape[0]
    widest = max(sum(b - a for a, b in interleaved_stripes(width, r, world, parts)) for r in range(world))
    pad = torch.zeros((h, widest, 3), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    frame = torch.empty((h, width, 3), dtype=local.dtype, device=local.device)
    for r in range(world):
        off = 0
        for x0, x1 in interleaved_stripes(width, r, world, parts):
            frame[:, x0:x1] = out[r][:, off:off + x1 - x0]
            off += x1 - x0
    return frame


def strided_columns(width, rank, world):
    """(x0, step, n_columns) of the finest interleave: rank r renders image columns r, r + world, r + 2 world, ... in ONE
    launch (PtParams.column_step = world), so every rank sees the same mix of cheap and expensive columns and pays the
    persistent kernel's tail once."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return rank, world, len(range(rank, width, world))


def gather_strided(local, width, group=None):
    """All-gather of per-rank dense images [H, n_columns(rank), 3] rendered with strided_columns() into the [H, W, 3] frame."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    h = local.shape[0]
    widest = -(-width // world)
    pad = torch.zeros((h, widest, 3), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    frame = torch.empty((h, width, 3), dtype=local.dtype, device=local.device)
    for r in range(world):
        n = len(range(r, width, world))
        frame[:, r::world] = out[r][:, :n]
    return frame


def gather_stripes(local, width, group=None):
    """All-gather of per-rank stripes [H, x1-x0, 3] uint8 (torch tensors, any backend) into the full [H, W, 3] frame.
    Stripes may differ by one column, so each is padded to the widest before the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    h = local.shape[0]
    widest = -(-width // world)
    pad = torch.zeros((h, widest, 3), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    frame = torch.empty((h, width, 3), dtype=local.dtype, device=local.device)
    for r in range(world):
        x0, x1 = stripe(width, r, world)
        frame[:, x0:x1] = out[r][:, :x1 - x0]
    return frame
