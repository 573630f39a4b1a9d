"""Multi-GPU side of the path: the batch is partitioned per image, ranks
never exchange anything on the data path, and ONE all-gather of fixed-size padded
detection records happens at the end (eval / final detections)."""
import torch
import torch.distributed as dist


def image_partition(num_images, rank, world_size):
    """Rank r owns images r, r+W, r+2W, ...."""
    return list(range(rank, num_images, world_size))


def pack_detections(boxes, scores, labels, max_per_img=100):
    """[k,4] boxes, [k] scores, [k] labels -> fixed record fp32 [max_per_img, 6] + int32 count."""
    k = min(int(scores.numel()), max_per_img)
    rec = torch.zeros((max_per_img, 6), dtype=torch.float32, device=scores.device)
    if k:
        rec[:k, :4] = boxes[:k]
        rec[:k, 4] = scores[:k]
        rec[:k, 5] = labels[:k].to(torch.float32)
    return rec, torch.tensor([k], dtype=torch.int32, device=scores.device)


def gather_detections(records, counts, group=None):
    """records [n_local, M, 6], counts [n_local] on every rank (same n_local) -> lists over
    ALL images in global image order (image i lives on rank i % W at local slot i // W)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return records, counts
    W = dist.get_world_size(group)
    recs = [torch.empty_like(records) for _ in range(W)]
    cnts = [torch.empty_like(counts) for _ in range(W)]
    dist.all_gather(recs, records.contiguous(), group=group)
    dist.all_gather(cnts, counts.contiguous(), group=group)
    # interleave: global image g = slot * W + rank
    rec = torch.stack(recs, dim=1).reshape((-1,) + tuple(records.shape[1:]))
    cnt = torch.stack(cnts, dim=1).reshape(-1)
    return rec, cnt
