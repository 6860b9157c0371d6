"""N > 1 bookkeeping on CPU: world_size-2 gloo processes shard a batch by image, run their shard (with the CPU
oracle standing in for the kernels -- this test is about the sharding, not the arithmetic), and rank 0 checks that the
merged result equals the unsharded one and that the timing reduction returns the slowest rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import importlib.util
    import torch.distributed as dist
    import oracle
    import roi3d_synth
    spec = importlib.util.spec_from_file_location("roi3d_sharding", os.path.join(ROOT, "3d-mask-r-cnn_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, vol, crop = 5, (32, 32, 16), (4, 4, 4)
    rng = np.random.default_rng(3)
    image = rng.standard_normal((B, 8, 8, 16, 8), dtype=np.float32)
    boxes = roi3d_synth.rois(40, vol, seed=4)
    bidx = rng.integers(0, B, 40).astype(np.int32)
    grads = rng.standard_normal((40,) + crop + (8,), dtype=np.float32)
    b_loc, i_loc, pos = sh.shard_rois(boxes, bidx, B, world, rank)
    img_loc = sh.shard_volume(image, world, rank)
    crops = oracle.crop_and_resize_3d(img_loc, b_loc, i_loc, crop)
    gimg = oracle.crop_and_resize_3d_grad_image(grads[pos], b_loc, i_loc, img_loc.shape)
    nms_keep = {i: oracle.non_max_suppression_3d(boxes[bidx == i], rng.random(int((bidx == i).sum())).astype(np.float32) * 0 + 1, 5, 0.5)
                for i in sh.images_of_rank(B, world, rank)}
    slowest = sh.max_over_ranks(10.0 + rank)
    assert sh.gather_over_ranks(10.0 + rank) == [10.0 + r for r in range(world)]
    gathered = [None] * world
    dist.all_gather_object(gathered, (crops, pos, gimg, sh.images_of_rank(B, world, rank), nms_keep, slowest))
    if rank == 0:
        full = oracle.crop_and_resize_3d(image, boxes, bidx, crop)
        merged = sh.merge_by_position([g[0] for g in gathered], [g[1] for g in gathered], 40)
        full_g = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape)
        merged_g = np.zeros_like(full_g)
        for g in gathered:
            merged_g[g[3]] = g[2]
        owners = sorted(i for g in gathered for i in g[3])
        q.put((bool(np.array_equal(full, merged)), bool(np.array_equal(full_g, merged_g)), owners == list(range(B)),
               sorted(k for g in gathered for k in g[4]) == list(range(B)), [g[5] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_matches_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fwd_ok, bwd_ok, all_images_owned_once, nms_per_image, slowest = res
    assert fwd_ok and bwd_ok and all_images_owned_once and nms_per_image
    assert slowest == [11.0, 11.0]              # every rank sees the max over ranks


def test_shard_helpers():
    import importlib.util
    spec = importlib.util.spec_from_file_location("roi3d_sharding", os.path.join(ROOT, "3d-mask-r-cnn_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    assert sh.images_of_rank(8, 8, 3) == [3] and sh.images_of_rank(8, 2, 1) == [1, 3, 5, 7] and sh.images_of_rank(2, 4, 3) == []
    boxes = np.arange(36, dtype=np.float32).reshape(6, 6)
    b, i, pos = sh.shard_rois(boxes, np.array([0, 1, 2, 3, 1, 3]), 4, 2, 1)
    assert pos.tolist() == [1, 3, 4, 5] and i.tolist() == [0, 1, 0, 1] and np.array_equal(b, boxes[pos])
    assert sh.max_over_ranks(3.5) == 3.5        # no process group -> identity
