"""Two ranks, two GPUs, NCCL: the batch-sharded hot path equals the single-GPU result (BASELINE config 5).

Needs >= 2 GPUs (``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu``); skipped otherwise.  Each rank
post-processes and trains on its contiguous shard; the detections come back through ``sharding.PeerGather`` (NVLink peer
memory) AND ``sharding.gather_packed`` (NCCL all_gather), the per-image losses through ``sharding.reduce_image_losses``
(one all_reduce; the reference gathers its loss the same way, train.py:185-186, loss.py:210-213).  Rank 0 also runs the
whole batch alone: gathered detections must be bit-identical, the reduced batch-mean losses equal to 1e-5 (fp32
summation order differs between one mean over B and a sum of shard sums).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

BATCH, NCLS, MAX_GT = 6, 20, 12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import pytorch_object_detection_b200 as P
    from pytorch_object_detection_b200 import ops, sharding, workloads as W

    x = W.head_outputs(BATCH, NCLS, W.VOC_LEVELS, seed=901)
    gt, labels = W.gt_boxes(BATCH, MAX_GT, W.VOC_HW, NCLS, seed=902)
    lo, hi = sharding.shard_bounds(BATCH, world, rank)
    nb, k = hi - lo, 1000
    mine = [[t[lo:hi].to(dev).requires_grad_(True) for t in part] for part in x]
    head = P.FCOSHead(0.05, 0.6, k, W.STRIDES)
    step = P.FCOSTargetLoss(W.STRIDES, W.FCOS_RANGES, "giou")

    # ---- detections: shard -> packed buffer -> both gathers ---------------------------------------------------
    pk = ops.packed_detections(nb, k, dev)
    with torch.no_grad():
        s, c, b, n = head.detect([[t.detach() for t in part] for part in mine], clip_hw=W.VOC_HW, out_packed=pk)
    assert sharding.packed_of(s).data_ptr() == pk.data_ptr()
    via_nccl = sharding.gather_packed(pk)
    peer = sharding.PeerGather(pk.numel(), 2, dev)
    via_peer_all = peer.gather(0, pk).clone()
    via_peer_root = peer.gather(1, pk, root=0).clone()
    torch.cuda.synchronize()
    assert torch.equal(via_peer_all, via_nccl), "peer-memory all-gather differs from the NCCL all_gather"
    if rank == 0:
        assert torch.equal(via_peer_root, via_nccl), "peer-memory gather to rank 0 differs from the NCCL all_gather"
    with pytest.raises(ValueError):
        sharding.PeerGather(pk.numel(), 1, dev)                               # one slot cannot be read race-free

    # ---- losses: shard step -> per-image losses -> one all_reduce --------------------------------------------
    losses = step([mine, gt[lo:hi].to(dev), labels[lo:hi].to(dev)])
    losses[3].backward()
    per = step.per_image
    red = sharding.reduce_image_losses([per["cls"], per["cnt"], per["reg"]], BATCH)

    if rank == 0:
        full = [[t.to(dev).requires_grad_(True) for t in part] for part in x]
        with torch.no_grad():
            fs, fc, fb, fn = head.detect([[t.detach() for t in part] for part in full], clip_hw=W.VOC_HW)
        for r in range(world):
            r_lo, r_hi = sharding.shard_bounds(BATCH, world, r)
            gs, gc, gb, _, gn = ops.detection_views(via_nccl[r], r_hi - r_lo, k)
            assert torch.equal(gn, fn[r_lo:r_hi])
            for i in range(r_hi - r_lo):
                m = int(gn[i])
                assert torch.equal(gs[i, :m], fs[r_lo + i, :m]) and torch.equal(gc[i, :m], fc[r_lo + i, :m])
                assert torch.equal(gb[i, :m], fb[r_lo + i, :m])
        whole = P.FCOSTargetLoss(W.STRIDES, W.FCOS_RANGES, "giou")
        want = whole([full, gt.to(dev), labels.to(dev)])
        want[3].backward()
        for got, w in zip(red, want[:3]):
            assert abs(float(got) - float(w)) <= 1e-5 * abs(float(w)), (float(got), float(w))
        # gradients of the shard are those of the whole batch scaled by B / shard size (each rank's mean is over its
        # own images; data-parallel training averages the ranks' gradients)
        scale = nb / BATCH
        for part_s, part_f in zip(mine, full):
            for a, f in zip(part_s, part_f):
                torch.testing.assert_close(a.grad * scale, f.grad[lo:hi], rtol=1e-5, atol=1e-9)
        with open(os.path.join(out_dir, "ok"), "w") as fh:
            fh.write(f"detections of {BATCH} images bit-identical; losses {[float(v) for v in red]}\n")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_shards_equal_single_gpu(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))
