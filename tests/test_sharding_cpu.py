"""CPU: clip sharding + output gather over a world_size-2 gloo group (the N>1 host logic of bench.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from axial_vs_b200 import sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))
            sizes = sharding.shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(n_items, rank, world)
    local = torch.arange(a, b, dtype=torch.float32)[:, None] * torch.ones(1, 3)
    full = sharding.gather_clip_outputs(local, n_items)
    assert full.shape == (n_items, 3)
    assert torch.equal(full[:, 0], torch.arange(n_items, dtype=torch.float32))
    dist.destroy_process_group()


def test_gather_world2_ragged():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 7), nprocs=2, join=True)


# ---- cross-clip module over a clip-sharded video: one exchange of the clip queries, redundant layers, local mask logits
def _cc_case():
    import torch
    from axial_vs_b200 import synth
    Q, T, V, H, W, L, K, seed = 8, 5, 2, 3, 4, 1, 10, 77
    p = synth.cross_clip_params(seed, L, K)
    cq = synth.randn(seed + 1, 1, Q, T, 256)
    pf = synth.randn(seed + 2, 1, 128, T * V, H, W)
    return Q, T, V, H, W, L, p, cq, pf


def _cc_fns(p, L, V, Q):
    import torch
    from oracle import traj_oracle as O

    def refine(cq_full):
        T = cq_full.shape[2]
        r = O.cross_clip_module(cq_full, torch.zeros(1, 128, T * V, 1, 1), p, L, V)      # class logits do not depend on the pixels
        vq = r["clip_query"].permute(0, 2, 3, 1).reshape(T, 256, Q)
        me = O.conv_bn_1d(vq, p, "_mask_embedding_projection", "syncbn", "gelu")
        mk = O.conv_bn_1d(me, O._sub(p, "_predictor"), "_transformer_mask_head", "syncbn", None)      # [T, 128, Q]
        return r["pred_logits"], mk.permute(0, 2, 1).reshape(T * Q, 128)

    def masks(mk_local, pf_local, t_local):
        mk = mk_local.reshape(t_local, Q, 128)
        lg = torch.einsum("tcp,tqc->qtp", pf_local.flatten(2), mk)
        return O.batch_norm_eval(lg.unsqueeze(0), O._sub(p, "_predictor"), "_pixel_space_mask_batch_norm").squeeze(0)

    return refine, masks


def _cc_worker(rank, world, port):
    import torch
    import torch.distributed as dist
    from axial_vs_b200 import sharding
    from oracle import traj_oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    Q, T, V, H, W, L, p, cq, pf = _cc_case()
    ref = O.cross_clip_module(cq, pf, p, L, V)
    t0, t1 = sharding.shard_range(T, rank, world)                         # 5 clips over 2 ranks: ragged shards
    pf_clips = pf.reshape(1, 128, T, V, H, W).permute(0, 2, 1, 3, 4, 5).reshape(T, 128, V * H, W)
    refine, masks = _cc_fns(p, L, V, Q)
    cls, ml = sharding.cross_clip_sharded(refine, masks, cq[:, :, t0:t1].contiguous(), pf_clips[t0:t1].contiguous(), T)
    assert torch.allclose(cls, ref["pred_logits"], atol=1e-5)
    assert torch.allclose(ml.reshape(1, Q, T * V, H, W), ref["pred_masks"], atol=1e-4)
    dist.destroy_process_group()


def test_cross_clip_sharded_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_cc_worker, args=(2, port), nprocs=2, join=True)
