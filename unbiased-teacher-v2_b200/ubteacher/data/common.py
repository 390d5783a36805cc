"""Aspect-ratio grouping of the two-crop streams (reference: ubteacher/data/common.py:93-167,
``AspectRatioGroupedSemiSupDatasetTwoCrop``): landscape and portrait images are batched separately so that a batch
needs little padding. Each stream element is a pair (strong-view dict, weak-view dict) of the same image.

Behaviour kept from the reference, including its quirk: the labeled and unlabeled streams are advanced in lock step, and
while one side's current bucket is already full its incoming pairs are consumed and dropped until the other side fills."""


class _Side:
    """One stream's two buckets (w > h, w <= h) and the bucket the side is currently filling."""

    def __init__(self, batch_size):
        self.batch_size = batch_size
        self.strong = ([], [])
        self.weak = ([], [])
        self.cur = None              # index of the bucket touched last (None before the first element)

    def full(self):
        return self.cur is not None and len(self.strong[self.cur]) == self.batch_size

    def offer(self, pair):
        if self.full():
            return                   # dropped (reference: the append is skipped while the bucket waits for the other side)
        q, k = pair
        self.cur = 0 if q["width"] > q["height"] else 1
        self.strong[self.cur].append(q)
        self.weak[self.cur].append(k)

    def take(self):
        out = (self.strong[self.cur][:], self.weak[self.cur][:])
        del self.strong[self.cur][:]
        del self.weak[self.cur][:]
        return out


class AspectRatioGroupedSemiSupDatasetTwoCrop:
    def __init__(self, dataset, batch_size):
        self.label_dataset, self.unlabel_dataset = dataset
        self.batch_size_label, self.batch_size_unlabel = batch_size[0], batch_size[1]

    def __iter__(self):
        lab, unl = _Side(self.batch_size_label), _Side(self.batch_size_unlabel)
        for d_label, d_unlabel in zip(self.label_dataset, self.unlabel_dataset):
            lab.offer(d_label)
            unl.offer(d_unlabel)
            if lab.full() and unl.full():
                lq, lk = lab.take()
                uq, uk = unl.take()
                yield lq, lk, uq, uk      # label_strong, label_weak, unlabel_strong, unlabel_weak
