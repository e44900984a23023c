"""Image-level sharding of a batch over GPUs: images are independent units, image k of n goes to
rank k*world//n (contiguous blocks), and there is no data-path collective.  torch.distributed is used only for the
plumbing around it (barriers, max-over-ranks timing, gathering small per-image summaries)."""
from __future__ import annotations


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous block of item indices owned by `rank`: item k -> rank k*world//n_items."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    lo = -(-rank * n_items // world)          # ceil(rank*n/world)
    hi = -(-(rank + 1) * n_items // world)
    return range(lo, hi)


def owner_of(item: int, n_items: int, world: int) -> int:
    return item * world // n_items


def run_sharded(images, process, world: int, rank: int, dist=None):
    """Runs `process(image) -> summary tensor` on this rank's shard and gathers every image's summary on all ranks
    (object gather of small summaries only; pixel data never crosses ranks).  Returns {image index: summary}."""
    mine = {k: process(images[k]) for k in shard_range(len(images), world, rank)}
    if dist is None or world == 1:
        return mine
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    out = {}
    for part in gathered:
        out.update(part)
    return out
