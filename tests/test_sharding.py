"""Host logic of the image-sharded multi-GPU eval (rrnet_b200.host.sharding) on CPU: world_size-2
gloo processes exchange padded detections with the same calls the NCCL path uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rrnet_b200.host import sharding


def test_shard_range_is_contiguous_and_balanced():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_images, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every image i has (i*7 % (K+1)) detections whose rows encode (image, row)
        lo, hi = sharding.shard_range(n_images, rank, world)
        per = -(-n_images // world)                                  # equal shard capacity (padded)
        counts, rows = [], []
        for i in range(lo, hi):
            c = (i * 7) % (K + 1)
            counts.append(c)
            r = torch.zeros(c, 6)
            r[:, 0] = i
            r[:, 1] = torch.arange(c)
            rows.append(r)
        rows = torch.cat(rows) if rows else torch.zeros(0, 6)
        padded = sharding.pack_detections(rows, counts, per, K)
        cnt = torch.zeros(per, dtype=torch.int32)
        cnt[: len(counts)] = torch.tensor(counts, dtype=torch.int32)
        all_p, all_c = sharding.all_gather_detections(padded, cnt)
        assert tuple(all_p.shape) == (world * per, K, 6)
        # rank-major order with per-shard padding: image i of shard r sits at r*per + (i - lo_r)
        ok = True
        for r in range(world):
            rlo, rhi = sharding.shard_range(n_images, r, world)
            for i in range(rlo, rhi):
                slot = r * per + (i - rlo)
                c = (i * 7) % (K + 1)
                ok &= int(all_c[slot]) == c
                ok &= bool((all_p[slot, :c, 0] == i).all()) and bool((all_p[slot, :c, 1] == torch.arange(c)).all())
                ok &= bool((all_p[slot, c:] == 0).all())
        # the one-buffer form (ops.EvalPath.result_blob): rows then the int32 counts as raw bits
        n_rows = per * K
        blob = torch.zeros(n_rows * 6 + per + 1)
        blob[: n_rows * 6] = padded.reshape(-1)
        cview = blob[n_rows * 6:].view(torch.int32)
        cview[:per] = cnt
        cview[per] = int(cnt.sum())
        rows_all, counts_all = sharding.all_gather_result_blobs(blob, n_rows, per)
        ok &= tuple(rows_all.shape) == (world, n_rows, 6) and tuple(counts_all.shape) == (world, per + 1)
        ok &= bool(torch.equal(rows_all.reshape(world * per, K, 6), all_p))
        ok &= bool(torch.equal(counts_all[:, :per].reshape(-1), all_c))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_all_gather_detections_gloo_world2():
    world, n_images, K = 2, 5, 9
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_images, K, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_single_process_passthrough():
    p = torch.zeros(3, 4, 6)
    c = torch.tensor([1, 0, 4], dtype=torch.int32)
    ap, ac = sharding.all_gather_detections(p, c)
    assert ap is p and ac is c
    got = sharding.unpack_detections(ap, ac, 3)
    assert [t.shape[0] for t in got] == [1, 0, 4]
