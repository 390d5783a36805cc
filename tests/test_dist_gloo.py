"""world_size-2 `gloo` tests (CPU) of the host-side multi-GPU logic: the comm shims, per-rank data sharding, and the
FCOS normaliser semantics the kernels implement (num_pos / sum-ctr are SUM-all-reduced then divided by world:
fcos_outputs.py:319-321,362) — checked through the oracle: 2 ranks on per-rank shards == the formula applied by hand."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "unbiased-teacher-v2_b200")]
    from oracle import ut2_oracle as O
    from ubteacher.d2compat import comm
    from ubteacher.data.synthetic import SyntheticTwoCropLoader

    out = {}
    out["world"] = comm.get_world_size()
    out["sum"] = float(comm.reduce_sum(torch.tensor([float(rank + 1)])))
    g = comm.gather({"rank": rank, "loss": 1.0 + rank})
    out["gather"] = [x["loss"] for x in g] if comm.is_main_process() else g
    ld = SyntheticTwoCropLoader(1, 1, h=32, w=48, rank=rank, pin=False)
    lq, lk, uq, uk = next(ld)
    out["img_sum"] = int(lq[0]["image"].long().sum())
    # FCOS losses on this rank's shard with the world-averaged normalisers
    gen = torch.Generator().manual_seed(100 + rank)
    hw = [(8, 10), (4, 5), (2, 3), (1, 2), (1, 1)]
    strides = [8, 16, 32, 64, 128]
    locs = [O.compute_locations(h, w, s) for (h, w), s in zip(hw, strides)]
    mk = lambda c: [torch.randn(1, c, h, w, generator=gen) for h, w in hw]
    logits, reg, std, ctr = mk(80), mk(68), mk(4), mk(1)
    boxes = [torch.tensor([[4.0, 6.0, 60.0 + 10 * rank, 50.0], [20.0, 10.0, 70.0, 62.0]])[: 2 - rank]]
    classes = [torch.tensor([3, 5])[: 2 - rank]]

    def allreduce(t):
        t = t.clone()
        dist.all_reduce(t)
        return t
    losses, _ = O.fcos_losses_labeled(logits, reg, std, ctr, locs, boxes, classes, strides, world_size=world, allreduce=allreduce)
    local, _ = O.fcos_losses_labeled(logits, reg, std, ctr, locs, boxes, classes, strides)
    tg = O.fcos_assign_targets_fast(locs, boxes, classes, strides)
    npos = float((torch.cat(tg["labels"]) != 80).sum())
    out["losses"] = {k: float(v) for k, v in losses.items()}
    out["local"] = {k: float(v) for k, v in local.items()}
    out["npos"] = npos
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["world"] == 2 and res[0]["sum"] == 3.0 and res[1]["sum"] == 3.0
    assert res[0]["gather"] == [1.0, 2.0] and res[1]["gather"] == []
    assert res[0]["img_sum"] != res[1]["img_sum"]            # ranks draw different shards
    # world-averaged normaliser: cls loss scales as local_npos_norm / avg_npos_norm
    avg = max((res[0]["npos"] + res[1]["npos"]) / 2, 1.0)
    for r in (0, 1):
        local_norm = max(res[r]["npos"], 1.0)
        expect = res[r]["local"]["loss_fcos_cls"] * local_norm / avg
        assert abs(res[r]["losses"]["loss_fcos_cls"] - expect) <= 1e-5 * abs(expect)
        expect = res[r]["local"]["loss_fcos_ctr"] * local_norm / avg
        assert abs(res[r]["losses"]["loss_fcos_ctr"] - expect) <= 1e-5 * abs(expect) + 1e-7
