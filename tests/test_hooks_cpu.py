"""Host logic of the training-loop hooks, the dataset catalog and the annotation box modes (no GPU needed)."""
import json

import pytest


class _Ckpt:
    def __init__(self):
        self.saved = []

    def save(self, name, **kw):
        self.saved.append((name, kw.get("iteration")))


class _Trainer:
    def __init__(self, max_iter):
        from ubteacher.d2compat.events import EventStorage
        self.max_iter, self.iter, self.storage = max_iter, 0, EventStorage(0)


def _run(trainer, hooks):
    for h in hooks:
        h.trainer = trainer
        h.before_train()
    for trainer.iter in range(trainer.max_iter):
        trainer.storage.put_scalar("loss_x", 1.0 / (trainer.iter + 1))
        for h in hooks:
            h.after_step()
        trainer.storage.step()
    for h in hooks:
        h.after_train()


def test_periodic_checkpointer_eval_hook_and_writer(tmp_path):
    from ubteacher.engine import hooks
    ck, tr, evals = _Ckpt(), _Trainer(10), []
    w = hooks.JSONWriter(str(tmp_path / "metrics.json"))
    hs = [hooks.PeriodicCheckpointer(ck, 4), hooks.EvalHook(5, lambda: evals.append(tr.iter) or {"bbox": {"AP": 12.5}}),
          hooks.PeriodicWriter([w], period=3)]
    _run(tr, hs)
    assert ck.saved == [("model_0000003", 3), ("model_0000007", 7), ("model_final", 9)]       # [D2] PeriodicCheckpointer naming
    assert evals == [4, 9]                                                                      # every 5 iterations + after the last
    assert tr.storage.latest()["bbox/AP"][0] == 12.5
    lines = [json.loads(x) for x in open(tmp_path / "metrics.json")]
    assert [r["iteration"] for r in lines] == [2, 5, 8, 9] and "loss_x" in lines[0]


def test_eval_hook_disabled_with_period_zero():
    from ubteacher.engine import hooks
    tr, calls = _Trainer(4), []
    _run(tr, [hooks.EvalHook(0, lambda: calls.append(1))])
    assert calls == []


def test_dataset_catalog_and_bbox_modes(tmp_path):
    from ubteacher.d2compat.catalog import DatasetCatalog, register_json
    from ubteacher.data.dataset_mapper import bbox_xyxy
    p = tmp_path / "d.json"
    p.write_text(json.dumps([{"file_name": "a.jpg", "height": 4, "width": 6, "image_id": 1,
                              "annotations": [{"bbox": [1, 2, 3, 4], "bbox_mode": 1, "category_id": 0}]}]))
    register_json("toy_train", str(p))
    assert "toy_train" in DatasetCatalog and DatasetCatalog.get("toy_train")[0]["image_id"] == 1
    with pytest.raises(KeyError):
        DatasetCatalog.get("missing")
    assert bbox_xyxy({"bbox": [1, 2, 3, 4], "bbox_mode": 1}) == [1.0, 2.0, 4.0, 6.0]          # XYWH_ABS -> XYXY_ABS
    assert bbox_xyxy({"bbox": [1, 2, 3, 4]}) == [1.0, 2.0, 3.0, 4.0] == bbox_xyxy({"bbox": [1, 2, 3, 4], "bbox_mode": 0})
    with pytest.raises(ValueError):
        bbox_xyxy({"bbox": [0, 0, 1, 1], "bbox_mode": 4})


def test_graph_cache_policy_first_sight_eager_then_entry_and_bounded(monkeypatch):
    """Trainer._graph_entry: a batch geometry gets a graph entry on its SECOND sight, geometries seen once are only keys in a
    bounded LRU (a dataset with free aspect ratios never repeats a batch), and the number of entries is capped."""
    import torch
    from ubteacher.engine.trainer import UBTeacherTrainer

    class Stub:
        _graphs, _seen_once, graph_padded_key = {}, {}, True
        _batch_key = UBTeacherTrainer._batch_key

        class model:
            class engine:
                padded_size = staticmethod(lambda sizes: (max(s[0] for s in sizes) + 31 & ~31, max(s[1] for s in sizes) + 31 & ~31))

    def batch(h, w):
        img = {"image": torch.empty(3, h, w, dtype=torch.uint8)}
        return ([img], [img], [img, img], [img, img])

    st = Stub()
    entry = lambda d, **kw: UBTeacherTrainer._graph_entry(st, d, **kw)
    a = batch(128, 160)
    assert entry(a) is None and len(st._seen_once) == 1 and not st._graphs          # first sight: eager
    assert entry(a, create=False) is None
    e = entry(a)
    assert e is not None and e["graph"] is None and not st._seen_once                # second sight: an entry to capture into
    assert entry(a) is e and entry(a, create=False) is e
    for i in range(1500):                                                             # never-repeating geometries
        assert entry(batch(64 + i, 96)) is None
    assert len(st._seen_once) == 1024 and len(st._graphs) == 1
    monkeypatch.setenv("UT2_GRAPH_CACHE", "2")
    b, c = batch(160, 192), batch(192, 224)
    entry(b); entry(c)
    assert entry(b) is not None and entry(c) is None and len(st._graphs) == 2         # cap reached: c stays eager
    # batches of mixed image sizes are keyed by the padded size of their forward groups (sizes travel as device data)
    def mixed(sizes):
        im = [{"image": torch.empty(3, h, w, dtype=torch.uint8)} for h, w in sizes]
        return ([im[0]], [im[1]], [im[2]], [im[3]])
    k1 = st._batch_key(mixed([(120, 150), (128, 160), (100, 130), (90, 140)]))
    k2 = st._batch_key(mixed([(128, 155), (110, 160), (97, 129), (96, 131)]))
    assert k1 == k2 == ("padded", (2, 128, 160), (1, 128, 160), (1, 96, 160))
    st.graph_padded_key = False
    assert st._batch_key(mixed([(120, 150), (128, 160), (100, 130), (90, 140)]))[0] == (3, 120, 150)


def test_static_size_table_dies_with_its_buffer():
    """ops.STATIC_SIZES maps the address of a static input buffer to its device-side image sizes; the entry must not outlive the
    buffer (the allocator may hand the address to an unrelated image)."""
    import gc
    import torch
    from ubteacher import ops
    buf, sizes = torch.empty(64, dtype=torch.uint8), torch.zeros(2, 2, dtype=torch.int32)
    ops.register_static(ops.STATIC_SIZES, buf, sizes)
    ptr = buf.data_ptr()
    assert ops.lookup_static(ops.STATIC_SIZES, ptr) is sizes
    assert ops.lookup_static(ops.STATIC_SIZES, buf[:12].view(3, 2, 2).data_ptr()) is sizes      # this batch's view of the buffer
    assert ops.lookup_static(ops.STATIC_SIZES, ptr + 1) is None
    del buf
    gc.collect()
    assert ops.lookup_static(ops.STATIC_SIZES, ptr) is None and ptr not in ops.STATIC_SIZES
