"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on seeded inputs.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

Every fixture stores the seeds/shapes needed to regenerate inputs and weights with
``axial_vs_b200.synth`` plus the reference's fp32 outputs, and a checksum of the weights so a
generator drift is detected instead of silently mis-comparing.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from axial_vs_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle import traj_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name: str, **arrs):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in conv.items()})


@torch.no_grad()
def main():
    torch.manual_seed(0)
    TA = ref_loader.temporal_attention()
    PE = ref_loader.pos_embeddings()
    CC = ref_loader.cross_clip()

    # ---- 1. TrajectoryAttention alone (Vk, with attention maps) -------------------------------
    for tag, (Bp, F, n, seed) in {"a": (3, 2, 5, 11), "b": (2, 3, 7, 12)}.items():
        g = torch.Generator().manual_seed(seed)
        p = {}
        synth.traj_attn_params(g, "", 256, p)
        m = TA.TrajectoryAttention(256, 8, 0.0).eval()
        m.load_state_dict(p, strict=True)
        q = synth.randn(seed + 100, Bp, F * n, 256)
        v = synth.randn(seed + 200, Bp, F * n, 256)
        y, maps = m(q, q, v, num_frames=F)
        save(f"ta_vk_{tag}", Bp=Bp, F=F, n=n, seed=seed, y=y, maps=maps, wsum=synth.checksum(p))

    # ---- 2. axial layer, encoder, non-axial layer ---------------------------------------------
    for tag, (B, T, H, W, seed) in {"a": (1, 2, 6, 5, 21), "b": (2, 3, 4, 7, 22)}.items():
        p = synth.axial_layer_params(seed)
        m = TA.TemporalAxialTrajectoryAttentionLayer(256, 1024, 0.0, 0.0, "relu", 8).eval()
        m.load_state_dict(p, strict=True)
        src = synth.randn(seed + 100, B * T, H * W, 256)
        pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
        out, hm, wm = m(src, pos)
        save(f"axial_layer_{tag}", B=B, T=T, H=H, W=W, seed=seed, out=out, hmap=hm, wmap=wm,
             wsum=synth.checksum(p))

    B, T, H, W, seed = 1, 2, 5, 6, 31
    p = synth.encoder_params(seed, 2)
    m = TA.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 2).eval()
    m.load_state_dict(p, strict=True)
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[1])
    out, hm, wm = m(src, pos)
    save("encoder_axial", B=B, T=T, H=H, W=W, seed=seed, out=out, hmap=hm, wmap=wm, wsum=synth.checksum(p))

    B, T, H, W, seed = 1, 2, 4, 5, 41
    p = synth.encoder_params(seed, 1, axial=False)
    m = TA.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "trajectory", 1).eval()
    m.load_state_dict(p, strict=True)
    src = synth.randn(seed + 100, B * T, H * W, 256)
    pos = O.level_pos3d(B, T, H, W, synth.level_embed(seed + 200)[0])
    out, _, _ = m(src, pos)
    save("encoder_trajectory", B=B, T=T, H=H, W=W, seed=seed, out=out, wsum=synth.checksum(p))

    # ---- 3. 3-D sine positional table -----------------------------------------------------------
    pe = PE.PositionEmbeddingSine3D(128, normalize=True)
    for tag, (B, T, H, W) in {"a": (1, 2, 5, 7), "b": (2, 5, 3, 4)}.items():
        tab = pe(torch.zeros(B, T, 256, H, W), fmt="btchw")             # [B,T,C,H,W]
        save(f"pos3d_{tag}", B=B, T=T, H=H, W=W, table=tab.permute(0, 1, 3, 4, 2).contiguous())

    # ---- 4. cross-clip: attention alone, full module ----------------------------------------------
    Q, Tc, seed = 16, 3, 51
    g = torch.Generator().manual_seed(seed)
    p = {}
    synth.traj_attn_params(g, "", 256, p, fused_qkv=True)
    m = CC.TrajectoryAttention(256, 8, 0.0).eval()
    m.load_state_dict(p, strict=True)
    x = synth.randn(seed + 100, 2, Tc * Q, 256)
    save("cc_ta", b=2, Q=Q, T=Tc, seed=seed, y=m(x, seq_len=Q, num_frames=Tc), wsum=synth.checksum(p))

    Q, Tc, V, Hh, Ww, L, K, seed = 16, 4, 2, 6, 5, 2, 124, 61
    p = synth.cross_clip_params(seed, L, K)
    m = CC.CrossClipTrackingModule(num_layers=L, num_classes=K, attn_drop=0.0, aspp_drop=0.0,
                                   kernel_sizes=[3, 3, 3], atrous_rates=[1, 2, 3], norm_fn="ln",
                                   num_clip_frames=V).eval()
    m.load_state_dict(p, strict=True)
    cq = synth.randn(seed + 100, 1, Q, Tc, 256)
    pf = synth.randn(seed + 200, 1, 128, Tc * V, Hh, Ww)
    o = m(cq, pf)
    save("cc_module", Q=Q, T=Tc, V=V, H=Hh, W=Ww, L=L, K=K, seed=seed, pred_logits=o["pred_logits"],
         pred_masks=o["pred_masks"], aux_masks=o["aux_outputs"][0]["pred_masks"], wsum=synth.checksum(p))


    # ---- 5. clip-level decoder attention (row A11): tensors captured inside the reference kMaXTransformerLayer ---------
    DEC = sys.modules["maxtron_deeplab.modeling.transformer_decoder.maxtron_transformer_decoder"]
    for tag, (N, L, Cp, TH, Wd, seed) in {"a": (2, 128, 64, 10, 7, 71), "b": (1, 37, 32, 6, 5, 72)}.items():
        torch.manual_seed(seed)
        layer = DEC.kMaXTransformerLayer(num_classes=20, in_channel_pixel=Cp, in_channel_query=256, base_filters=128, num_heads=8,
                                         bottleneck_expansion=2, key_expansion=1, value_expansion=2).eval()
        g = torch.Generator().manual_seed(seed)
        for m in layer.modules():                       # non-trivial BN statistics and affine (norm_init=0 would zero the updates)
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(1.0 + 0.3 * torch.rand(m.running_var.shape, generator=g))
        cap = {}
        hooks = [
            layer._query_self_attention.register_forward_hook(lambda mod, i, o: cap.update(q=i[0], k=i[1], v=i[2], attn_out=o)),
            layer._pixel_v_conv_bn.register_forward_hook(lambda mod, i, o: cap.update(pixel_value=o.flatten(2))),
            layer._predictor.register_forward_hook(lambda mod, i, o: cap.update(mask_logits=o["mask_logits"].flatten(2))),
            layer._kmeans_query_batch_norm_retrieved_value.register_forward_hook(lambda mod, i, o: cap.update(kmeans_update=i[0])),
        ]
        layer(synth.randn(seed + 100, N, Cp, TH, Wd), synth.randn(seed + 200, N, 256, L))
        for h in hooks:
            h.remove()
        ao = layer._query_self_attention
        bn = {f"{nm}.{k}": v for nm in ("_batch_norm_similarity", "_batch_norm_retrieved_value")
              for k, v in getattr(ao, nm).state_dict().items()}
        save(f"decoder_attn_{tag}", N=N, L=L, seed=seed, **{k: v for k, v in cap.items()},
             **{"bn." + k: v for k, v in bn.items()})


    # ---- 6. within-clip input / output projections (row f1): the reference instantiates stock modules,
    #         nn.Sequential(nn.Conv2d(c_in, 256, 1), nn.GroupNorm(32, 256)) and the reverse (WC/msdeformattn.py:355-362)
    for tag, (n, c, H, W, seed) in {"a": (3, 128, 7, 5, 81), "b": (2, 512, 9, 11, 82)}.items():
        g = torch.Generator().manual_seed(seed)
        pin = {"0.weight": synth._xavier(g, 256, c, 1, 1), "0.bias": 0.1 * torch.randn(256, generator=g),
               "1.weight": 1.0 + 0.2 * torch.randn(256, generator=g), "1.bias": 0.1 * torch.randn(256, generator=g)}
        pout = {"0.weight": synth._xavier(g, 2 * c, 256, 1, 1), "0.bias": 0.1 * torch.randn(2 * c, generator=g),
                "1.weight": 1.0 + 0.2 * torch.randn(2 * c, generator=g), "1.bias": 0.1 * torch.randn(2 * c, generator=g)}
        mi = torch.nn.Sequential(torch.nn.Conv2d(c, 256, kernel_size=1), torch.nn.GroupNorm(32, 256)).eval()
        mo = torch.nn.Sequential(torch.nn.Conv2d(256, 2 * c, kernel_size=1), torch.nn.GroupNorm(32, 2 * c)).eval()
        mi.load_state_dict(pin, strict=True)
        mo.load_state_dict(pout, strict=True)
        x = synth.randn(seed + 100, n, c, H, W)
        tok = mi(x).flatten(2).transpose(1, 2).contiguous()                                       # :100-106
        y = mo(tok.transpose(1, 2).contiguous().view(n, -1, H, W))                                # :432-434
        save(f"proj_{tag}", n=n, c=c, H=H, W=W, seed=seed, tokens=tok, y=y, wsum=synth.checksum({**{"i." + k: v for k, v in pin.items()},
                                                                                                **{"o." + k: v for k, v in pout.items()}}))


    # ---- 7. MSDeformAttn spatial encoder layer (row f2): the reference module on CPU (its pure-PyTorch sampling branch)
    WCM = ref_loader.within_clip_module()
    for tag, (n, shapes, seed) in {"a": (2, [(3, 4), (5, 6), (7, 9)], 91), "b": (1, [(4, 4), (9, 7), (17, 13)], 92)}.items():
        p = synth.msda_layer_params(seed)
        m = WCM.MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4).eval()
        m.load_state_dict(p, strict=True)
        Len = sum(h * w for h, w in shapes)
        src = synth.randn(seed + 100, n, Len, 256)
        pos = synth.randn(seed + 200, n, Len, 256)
        ss = torch.tensor(shapes)
        lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
        ref = WCM.MSDeformAttnTransformerEncoder.get_reference_points(ss, torch.ones(n, 3, 2), "cpu")
        out = m(src, pos, ref, ss, lsi, None)
        save(f"msda_layer_{tag}", n=n, shapes=ss, seed=seed, out=out, ref_points=ref, wsum=synth.checksum(p))


    # ---- 8. the whole within-clip transformer encoder (rows A6 + f2): reference MSDeformAttnTransformerEncoder on CPU,
    #         2 stages x [spatial layer on 3 levels, TemporalEncoder(1 axial layer) on the first 2 levels]
    B, T, shapes, seed = 1, 2, [(3, 4), (5, 6), (7, 9)], 95
    sp = WCM.MSDeformAttnTransformerEncoderLayer(256, 1024, 0.0, "relu", 3, 8, 4)
    tp = WCM.TemporalEncoder(256, 1024, 0.0, 0.0, "relu", 8, "axial-trajectory", 1)
    enc = WCM.MSDeformAttnTransformerEncoder(sp, 2, 3, 2, tp).eval()
    state = {}
    for i in range(2):
        state.update({f"spatial_layers.{i}.{k}": v for k, v in synth.msda_layer_params(seed + i).items()})
        state.update({f"temporal_layers.{i}.{k}": v for k, v in synth.encoder_params(seed + 10 + i, 1).items()})
    enc.load_state_dict(state, strict=True)
    Len = sum(h * w for h, w in shapes)
    src = synth.randn(seed + 100, B * T, Len, 256)
    pos = synth.randn(seed + 200, B * T, Len, 256)
    le = synth.level_embed(seed + 300)
    pos3d = [O.level_pos3d(B, T, h, w, le[i]) for i, (h, w) in enumerate(shapes[:2])]
    ss = torch.tensor(shapes)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    out, _, _ = enc(src, ss, lsi, torch.ones(B * T, 3, 2), pos, torch.zeros(B * T, Len, dtype=torch.bool), pos3d)
    save("wc_encoder", B=B, T=T, shapes=ss, seed=seed, out=out, wsum=synth.checksum(state))


    # ---- 9. the whole within-clip tracking module (MSDeformAttnPixelDecoder.forward_features): projections, positional terms,
    #         2 stages x [spatial layer, temporal layer], output projections; B=1, T=2, three small levels
    D2L = sys.modules["detectron2.layers"]
    chans, sizes, seed = [512, 256, 256], [(3, 4), (5, 6), (7, 9)], 97            # res5, res4, res3
    shape = {"res2": D2L.ShapeSpec(channels=64, stride=4), "res3": D2L.ShapeSpec(channels=chans[2], stride=8),
             "res4": D2L.ShapeSpec(channels=chans[1], stride=16), "res5": D2L.ShapeSpec(channels=chans[0], stride=32)}
    mod = WCM.MSDeformAttnPixelDecoder(shape, transformer_dropout=0.0, transformer_attn_drop=0.0, transformer_nheads=8,
                                       transformer_dim_feedforward=1024, transformer_num_stages=2, transformer_spatial_layers=2,
                                       transformer_temporal_layers=2, transformer_temporal_attn_type="axial-trajectory", conv_dims=256,
                                       transformer_spatial_in_features=["res3", "res4", "res5"],
                                       transformer_temporal_in_features=["res4", "res5"], num_clip_frames=2, cross_clip_training=False).eval()
    p = synth.within_clip_module_params(seed, chans)
    mod.load_state_dict(p, strict=True)
    feats = {f"res{5 - i}": synth.randn(seed + 100 + i, 2, chans[i], *sizes[i]) for i in range(3)}
    out, _, _ = mod.forward_features(feats)
    save("wc_module", seed=seed, chans=chans, sizes=torch.tensor(sizes), res5=out["res5"], res4=out["res4"], res3=out["res3"], wsum=synth.checksum(p))
    panoptic_goldens()
    kmax_goldens()


@torch.no_grad()
def panoptic_goldens():
    # ---- 10. mask-wise panoptic post-processing (row f4): the unmodified MaXTronWCDeepLab.panoptic_mask_inference
    import types
    from oracle import panoptic_oracle as PO
    WCMODEL = ref_loader.wc_model().MaXTronWCDeepLab
    for tag, (seed, N, C, T, H, W, thr) in {"a": (11, 16, 6, 2, 12, 10, 0.3), "b": (12, 128, 124, 2, 33, 41, 0.3), "c": (13, 32, 19, 3, 20, 17, 0.4)}.items():
        thing, stuff, div = synth.panoptic_metadata(C)
        md = types.SimpleNamespace(thing_dataset_id_to_contiguous_id=thing, stuff_dataset_id_to_contiguous_id=stuff, label_divisor=div)
        ns = types.SimpleNamespace(class_threshold_thing=0.1, class_threshold_stuff=0.3, pixel_confidence_threshold=thr, overlap_threshold=0.8,
                                   reorder_class_weight=1.0, reorder_mask_weight=1.0, metadata=md)
        mc, mp, me = synth.panoptic_case(seed, N, C, T, H, W)
        seg, dic = WCMODEL.panoptic_mask_inference(ns, mc, mp, me)
        cats = sorted(dic.keys())
        embs = torch.cat([torch.stack(dic[c]) for c in cats]) if cats else torch.zeros(0, me.shape[1])
        mg = PO.margins(mc.numpy(), mp.numpy(), PO.Metadata(thing, stuff, div), thr, 0.1, 0.3)
        print(f"  panoptic_{tag}: {len(seg.unique())} ids, margins {mg}")
        save(f"panoptic_{tag}", seed=seed, N=N, C=C, T=T, H=H, W=W, thr=thr, seg=seg, cats=torch.tensor(cats, dtype=torch.int64),
             counts=torch.tensor([len(dic[c]) for c in cats], dtype=torch.int64), embs=embs)



@torch.no_grad()
def kmax_goldens():
    # ---- 11. kMaX pixel-decoder axial attention (row f3): the unmodified AxialAttention / AxialAttention2D in eval mode
    ref_loader.cross_clip()                                   # loads kmax_pixel_decoder.py
    KP = sys.modules["kmax_deeplab.modeling.pixel_decoder.kmax_pixel_decoder"]
    for tag, (N, C, L, seed) in {"a": (3, 128, 9, 31), "b": (2, 512, 41, 32)}.items():
        m = KP.AxialAttention(C, query_shape=L, total_key_depth=512, total_value_depth=1024, num_heads=8).eval()
        p = synth.kmax_axial_params(seed, C)
        m.load_state_dict(p, strict=True)
        y = m(synth.randn(seed + 100, N, C, L))
        save(f"kmax_axial_{tag}", N=N, C=C, L=L, seed=seed, y=y, wsum=synth.checksum(p))
    N, C, H, W, seed = 1, 256, 7, 10, 33
    m2 = KP.AxialAttention2D(C, query_shape=[H, W], filters=512, key_expansion=1, value_expansion=2, num_heads=8).eval()
    ph, pw = synth.kmax_axial_params(seed, C), synth.kmax_axial_params(seed + 1, 1024)
    m2._height_axis.load_state_dict(ph, strict=True)
    m2._width_axis.load_state_dict(pw, strict=True)
    y = m2(synth.randn(seed + 100, N, C, H, W))
    save("kmax_axial_2d", N=N, C=C, H=H, W=W, seed=seed, y=y, wsum=synth.checksum(ph) + synth.checksum(pw))


if __name__ == "__main__":
    if not ref_loader.available():
        raise SystemExit("reference tree not found; golden fixtures can only be generated where it is mounted")
    if sys.argv[1:] == ["panoptic"]:          # only the post-processing fixtures
        panoptic_goldens()
    elif sys.argv[1:] == ["kmax"]:
        kmax_goldens()
    else:
        main()
