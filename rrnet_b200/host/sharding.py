"""Image-sharded evaluation across the GPUs of one box.

The reference shards validation images with a DistributedSampler and "gathers" detections through
result files on a shared disk (operators/rrnet_operator.py:277-278, scripts/RRNet/eval.py:13-18).
Here every rank runs the post-backbone path on a contiguous range of images and the padded detections
[imgs,K,6] + per-image counts are exchanged with ONE all-gather each - or one in total through
`all_gather_result_blobs` - (NCCL over NVLink on GPUs; the
same code runs on gloo/CPU tensors in the tests).  No kernel of the path is followed by a collective,
so there is nothing to fuse a collective into."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced ranges: the first n_items % world ranks get one extra item."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(rows, counts, n_images, K):
    """rows [sum(counts),6] image-major + counts (list/1-D tensor) -> padded [n_images,K,6] (zeros) ."""
    out = rows.new_zeros(n_images, K, rows.size(1))
    base = 0
    for i, c in enumerate([int(c) for c in counts]):
        out[i, :c] = rows[base:base + c]
        base += c
    return out


def all_gather_detections(padded, counts, group=None):
    """padded [n_local,K,6] and counts [n_local] int32 (same n_local on every rank; pad the last shard)
    -> (all_padded [world*n_local,K,6], all_counts [world*n_local]) on every rank, rank-major order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return padded, counts
    all_p = padded.new_empty((world * padded.size(0),) + tuple(padded.shape[1:]))
    all_c = counts.new_empty(world * counts.size(0))
    dist.all_gather_into_tensor(all_p, padded.contiguous(), group=group)
    dist.all_gather_into_tensor(all_c, counts.contiguous(), group=group)
    return all_p, all_c


def all_gather_result_blobs(blob, n_rows, n_images, group=None):
    """ops.EvalPath.result_blob of every rank with ONE all-gather (it holds the final rows [n_rows,6] followed by the
    int32 per-image counts [n_images + 1], last = total rows, as raw bits in the same fp32 buffer)
    -> (rows [world,n_rows,6], counts [world,n_images+1] int32), rank-major."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = blob.reshape(1, -1)
    else:
        out = blob.new_empty(world, blob.numel())
        dist.all_gather_into_tensor(out.view(-1), blob.contiguous(), group=group)
    rows = out[:, : n_rows * 6].reshape(world, n_rows, 6)
    counts = out[:, n_rows * 6: n_rows * 6 + n_images + 1].contiguous().view(torch.int32)
    return rows, counts


def unpack_detections(all_padded, all_counts, n_images):
    """-> list of [count_i,6] tensors for the first n_images images (drops shard padding)."""
    return [all_padded[i, : int(all_counts[i])] for i in range(n_images)]
