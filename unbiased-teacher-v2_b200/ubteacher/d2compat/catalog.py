"""Minimal ``DatasetCatalog`` ([D2] detectron2.data.DatasetCatalog): name -> callable returning a list of dataset dicts
({"file_name", "height", "width", "image_id", "annotations": [{"bbox", "bbox_mode", "category_id", "iscrowd"}]}). The
reference registers COCO splits in ubteacher/data/datasets/builtin.py; here a user registers whatever is on disk
(train_net.py --dataset-json), the trainer picks it up by cfg.DATASETS.TRAIN / TEST."""
import json


class _Catalog(dict):
    def register(self, name, func):
        assert callable(func), "register a function that returns the dataset dicts"
        self[name] = func

    def get(self, name):
        if name not in self:
            raise KeyError(f"Dataset '{name}' is not registered (DatasetCatalog.register(name, func) or train_net.py --dataset-json)")
        return self[name]()


DatasetCatalog = _Catalog()


def register_json(name, path):
    def load():
        with open(path) as f:
            return json.load(f)
    DatasetCatalog.register(name, load)
