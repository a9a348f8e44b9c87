"""Contiguous pose-range sharding: rank/device g of G gets [g*n/G, (g+1)*n/G).
Identical to the split mbik_solve_batch_multi uses inside the library; no collective is involved."""


def shard_range(n_poses: int, rank: int, world: int):
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_poses * rank // world, n_poses * (rank + 1) // world
