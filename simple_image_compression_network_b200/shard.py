"""Batch sharding of the image stream over GPUs: every image (numReps index) is an independent
application of the layer with read-only weights (the SWG buffers reset per image, slidingwindow.h:1320,1351),
so rank r of W takes a contiguous range of images; weights are replicated; there is no data-path collective."""
from __future__ import annotations


def shard_range(n_images: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of the global batch owned by `rank` (strong-scaling split, remainder to low ranks) -- the rule
    fcb_shard_range / fcb_pool_run apply inside the library (csrc/fcb_pool.cu; tests/test_multirank.py keeps the two in step)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_images, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def weak_range(images_per_gpu: int, rank: int) -> tuple[int, int]:
    """Weak scaling (the benchmark): a fixed number of images per GPU; global image ids of this rank."""
    return rank * images_per_gpu, (rank + 1) * images_per_gpu
