from .box_ap import BoxAPEvaluator, coco_box_ap  # noqa: F401
from .evaluator import DatasetEvaluators, inference_context, inference_on_dataset  # noqa: F401
