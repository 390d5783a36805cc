"""String-keyed registries with the Detectron2 names the reference configs refer to
(META_ARCH_REGISTRY etc.). Uses detectron2's own registries when available."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, f"An object named '{name}' was already registered in '{self._name}' registry!"
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map


try:  # pragma: no cover
    from detectron2.modeling import (BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY,
                                     ROI_HEADS_REGISTRY)
except Exception:  # noqa: BLE001
    META_ARCH_REGISTRY = Registry("META_ARCH")
    BACKBONE_REGISTRY = Registry("BACKBONE")
    PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
    ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
