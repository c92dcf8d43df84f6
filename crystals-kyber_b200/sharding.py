"""Sharding of independent keys / ciphertexts across the GPUs of one box.

Every item is independent, so the batch index range is cut into contiguous shards, one per rank; there is
no collective on the data path.  Seeds are derived from the GLOBAL item index, so the concatenated outputs
do not depend on the number of ranks.
"""


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous shard [begin, end) of rank `rank`: [rank*n/W, (rank+1)*n/W)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (n_items * rank) // world_size, (n_items * (rank + 1)) // world_size
