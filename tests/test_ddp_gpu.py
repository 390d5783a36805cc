"""Two-GPU data-parallel parity: two ranks on per-image shards (NCCL all-reduce of the gradient arena, mean
folded into SGD, world-averaged FCOS normalisers) must move the student exactly like one process on the whole batch.
Skipped on a single-GPU box."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_batch(step):
    from ubteacher.data.synthetic import synth_instances
    g = torch.Generator().manual_seed(500 + step)
    h, w = 128, 160

    def mk(n, with_gt):
        out = []
        for _ in range(n):
            d = {"image": torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8), "height": h, "width": w}
            if with_gt:
                d["instances"] = synth_instances(g, h, w, 3)
            out.append(d)
        return out
    lq = mk(2, True)
    lk = [dict(d, image=torch.flip(d["image"], [0])) for d in lq]
    return lq, lk, mk(2, False), mk(2, False)


class _ShardLoader:
    """mode "same": every rank (and the single process) trains on shard 0; mode "split": rank r on shard r."""

    def __init__(self, rank, mode):
        self.rank, self.mode, self.step = rank, mode, 0

    def __iter__(self):
        return self

    def __next__(self):
        parts = _global_batch(self.step)
        self.step += 1
        k = 0 if self.mode == "same" else self.rank
        return tuple([dict(p[k])] for p in parts)


def _run(rank, world, port, q, mode):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "unbiased-teacher-v2_b200"), os.path.join(ROOT, "tests")]
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from util_cfg import fcos_cfg
    from ubteacher.d2compat.events import EventStorage
    from ubteacher.engine import UBTeacherTrainer
    cfg = fcos_cfg(**{"MODEL.DEVICE": f"cuda:{rank}", "SOLVER.BASE_LR": 0.01})
    tr = UBTeacherTrainer(cfg, data_loader=_ShardLoader(rank, mode))
    s0 = tr.model.engine.arena.data.clone()
    with EventStorage(0) as tr.storage:
        for it in range(2):
            tr.iter = it
            tr.run_step_full_semisup()
            tr.scheduler.step()
    torch.cuda.synchronize()
    data = tr.model.engine.arena.data
    delta = (data - s0).cpu()
    digest = [float(data.double().sum()), float((data.double() * torch.arange(data.numel(), device=data.device).double().remainder(977.0)).sum())]
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    idx = torch.randperm(delta.numel(), generator=torch.Generator().manual_seed(1))[:200000]
    # numpy: pickled by value (a torch tensor in the queue would need this process to stay alive)
    q.put({"rank": rank, "norm": float(delta.norm()), "sample": delta[idx].numpy().copy(), "digest": digest})


def _spawn(world, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r["rank"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_data_parallel_step():
    # (a) two ranks on the SAME shard == one process on that shard: the summed gradient arena times 1 / world, the
    #     world-averaged normalisers and the initial broadcast leave the update unchanged
    same = _spawn(2, "same")
    one = _spawn(1, "same")[0]
    a, b = torch.from_numpy(same[0]["sample"]).double(), torch.from_numpy(one["sample"]).double()
    rel = float((a - b).norm() / b.norm())
    assert one["norm"] > 0 and rel < 2e-2, (rel, same[0]["norm"], one["norm"])     # fp32-atomic / bf16 summation order only
    assert same[0]["digest"] == same[1]["digest"]
    # (b) different shards: the replicas stay bit-identical, and the update differs from the single-shard one
    split = _spawn(2, "split")
    assert split[0]["digest"] == split[1]["digest"]
    c = torch.from_numpy(split[0]["sample"]).double()
    assert float((c - b).norm() / b.norm()) > 5 * rel
