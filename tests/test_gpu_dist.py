"""Multi-GPU paths on real devices (NCCL, 2 ranks): data-parallel Gram caching with the single all-reduce,
and tensor-sharded merges with the final all-gather.  Skipped on boxes with fewer than 2 GPUs; the same
host logic is covered on CPU with gloo in tests/test_dist_cpu.py."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    import vl_merging_b200 as vlm

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = vlm.vlmo_config("tiny")
        model = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1).cuda()
        cache = vlm.GramCache()
        cache.register(model)
        cache8 = vlm.GramCache(precision="int8x4")       # fp64 Gram buffers: the packed fp64 exchange
        cache8.register(model)
        # the same exchange as ONE kernel over NVSwitch multicast memory (symmetric arenas), fp32 and fp64
        cache_mc = vlm.GramCache(symmetric=True)
        cache_mc.register(model)
        cache_mc8 = vlm.GramCache(precision="fp64", symmetric=True)
        cache_mc8.register(model)
        with torch.no_grad():
            model(vlm.synthetic_batch(2, cfg, seed=50 + rank, device="cuda"))   # this rank's shard of the calibration set
        cache.all_reduce()
        cache8.all_reduce()
        grams = cache.state_dict()
        grams8 = cache8.state_dict()
        mc = {}
        try:
            cache_mc.all_reduce()
            cache_mc8.all_reduce()
            gm, gm8 = cache_mc.state_dict(), cache_mc8.state_dict()
            mc = {"err32": max(((gm[k] - grams[k]).norm() / grams[k].norm()).item() for k in grams),
                  "err64": max(((gm8[k] - grams8[k]).norm() / grams8[k].norm()).item() for k in grams8),
                  "sym": all(torch.equal(gm[k], gm[k].T) and torch.equal(gm8[k], gm8[k].T) for k in grams),
                  "pick32": gm["transformer.blocks.7.mlp.l.fc2"].numpy(), "pick64": gm8["transformer.blocks.0.attn.v"].numpy(),
                  "calls": dict(cache_mc.calls) == dict(cache.calls), "n": len(gm)}
        except RuntimeError as e:
            if "multicast" not in str(e):
                raise
            mc = {"unsupported": str(e)}
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        mcfg = dict(vlffn_start_layer_index=10, only_activate_used_experts=False, merge_ratio=0.5, sum_lambda=0.75,
                    scaling_for_non_diag=0.9, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0})
        merged = vlm.merge_weights(sd, mcfg, group=dist.group.WORLD)
        rm = vlm.regmean(sd, mcfg, gram_matrices=grams, group=dist.group.WORLD)
        ufo = vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)).eval().cuda()
        ufo.load_state_dict(merged, strict=False)
        ib = [vlm.synthetic_batch(3, cfg, seed=70 + i, device="cuda") for i in range(3)]
        tb = [vlm.synthetic_batch(4, cfg, seed=80 + i, device="cuda", pad=True) for i in range(5)]
        img_f, txt_f = vlm.irtr_features(ufo, ib, tb, group=dist.group.WORLD)     # batches sharded over the ranks
        img_1, txt_1 = vlm.irtr_features(ufo, ib, tb)                              # every batch on this rank
        irtr_err = max((img_f - img_1).abs().max().item(), (txt_f - txt_1).abs().max().item())
        pick = ["transformer.blocks.0.attn.qkv.weight", "transformer.blocks.11.mlp.fc2.weight", "transformer.blocks.4.norm2.bias"]
        ret[rank] = {
            "grams": {k: grams[k].numpy() for k in ("transformer.blocks.0.attn.v", "transformer.blocks.7.mlp.l.fc2")},
            "grams8": {k: grams8[k].numpy() for k in ("transformer.blocks.0.attn.v", "transformer.blocks.7.mlp.l.fc2")},
            "reduce_bytes": (cache.last_reduce_bytes, cache8.last_reduce_bytes),
            "mc": mc,
            "n_grams": len(grams),
            "merged": {k: merged[k].cpu().numpy() for k in pick},
            "regmean": {k: rm[k].cpu().numpy() for k in pick},
            "irtr_err": irtr_err, "irtr_shapes": (tuple(img_f.shape), tuple(txt_f.shape)),
        }
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_calibration_and_sharded_merge():
    import torch.multiprocessing as mp

    import vl_merging_b200 as vlm

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["n_grams"] == r1["n_grams"] == 96
    assert r0["irtr_shapes"] == ((9, 192), (20, 192)) and max(r0["irtr_err"], r1["irtr_err"]) < 1e-5
    # single-process oracle for the reduced result: one cache over both shards
    cfg = vlm.vlmo_config("tiny")
    model = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1).cuda()
    cache = vlm.GramCache()
    cache.register(model)
    cache8 = vlm.GramCache(precision="int8x4")
    cache8.register(model)
    with torch.no_grad():
        for rank in range(2):
            model(vlm.synthetic_batch(2, cfg, seed=50 + rank, device="cuda"))
    want, want8 = cache.state_dict(), cache8.state_dict()
    for k, g in r0["grams"].items():
        assert np.array_equal(g, r1["grams"][k])
        assert np.linalg.norm(g - want[k].numpy()) <= 1e-5 * np.linalg.norm(want[k].numpy())
    if "unsupported" not in r0["mc"]:                       # the multimem exchange: same sums, bit-identical on both ranks
        assert r0["mc"]["n"] == 96 and r0["mc"]["calls"] and r0["mc"]["sym"] and r1["mc"]["sym"]
        assert max(r0["mc"]["err32"], r1["mc"]["err32"]) < 1e-6 and max(r0["mc"]["err64"], r1["mc"]["err64"]) < 1e-14
        assert np.array_equal(r0["mc"]["pick32"], r1["mc"]["pick32"]) and np.array_equal(r0["mc"]["pick64"], r1["mc"]["pick64"])
    for k, g in r0["grams8"].items():                       # the RegMean-grade cache: packed fp64 upper triangles travel
        assert np.array_equal(g, r1["grams8"][k]) and np.array_equal(g, g.T)
        assert np.linalg.norm(g - want8[k].numpy()) <= 1e-13 * np.linalg.norm(want8[k].numpy())
    assert r0["reduce_bytes"][1] == 2 * r0["reduce_bytes"][0] == 2 * 4 * sum(d * (d + 1) // 2 for d in [192] * 72 + [768] * 24)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    mcfg = dict(vlffn_start_layer_index=10, only_activate_used_experts=False, merge_ratio=0.5, sum_lambda=0.75,
                scaling_for_non_diag=0.9, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0})
    single = vlm.merge_weights(sd, mcfg)
    single_rm = vlm.regmean(sd, mcfg, gram_matrices=want)
    for k in r0["merged"]:
        assert np.array_equal(r0["merged"][k], r1["merged"][k])
        assert np.array_equal(r0["merged"][k], single[k].cpu().numpy())          # sharding does not change a bit
        a, b = r0["regmean"][k], single_rm[k].cpu().numpy()
        assert np.array_equal(r0["regmean"][k], r1["regmean"][k])
        assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(b)
