"""[D2] WarmupMultiStepLR: lr = base * warm(it) * gamma^{#milestones <= it}."""
from bisect import bisect_right


class WarmupMultiStepLR:
    def __init__(self, optimizer, milestones, gamma=0.1, warmup_factor=0.001, warmup_iters=1000, warmup_method="linear",
                 last_epoch=-1):
        self.optimizer = optimizer
        self.milestones = sorted(milestones)
        self.gamma, self.warmup_factor, self.warmup_iters, self.warmup_method = gamma, warmup_factor, warmup_iters, warmup_method
        self.base_lrs = [g["initial_lr"] for g in optimizer.param_groups]
        self.last_epoch = last_epoch
        self.step()

    def _warm(self, it):
        if it >= self.warmup_iters:
            return 1.0
        if self.warmup_method == "constant":
            return self.warmup_factor
        alpha = it / self.warmup_iters
        return self.warmup_factor * (1 - alpha) + alpha

    def get_lr(self):
        w = self._warm(self.last_epoch)
        return [b * w * self.gamma ** bisect_right(self.milestones, self.last_epoch) for b in self.base_lrs]

    def step(self):
        self.last_epoch += 1
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g["lr"] = lr

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "base_lrs": list(self.base_lrs)}

    def load_state_dict(self, sd):
        self.last_epoch = sd["last_epoch"] - 1
        self.step()
