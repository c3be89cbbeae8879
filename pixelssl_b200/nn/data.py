"""Rank-aware two-stream batch sampler.

The reference builds ONE global batch per step on the host — ``labeled_batch_size`` labeled indices
followed by ``unlabeled_batch_size`` unlabeled ones, both already multiplied by the GPU count
(task_template/proxy.py:252-261, 372-375) — and lets ``nn.DataParallel`` scatter it in contiguous
chunks, so replica 0 sees mostly labeled and the last replica only unlabeled samples
(pixelssl/nn/data.py:126-177).  With one process per GPU every rank loads its own batch, so the
sampler is made rank-aware instead:

  * every rank draws the SAME global permutations (same ``np.random`` consumption as the reference
    sampler, so with ``world_size == 1`` the index stream is identical to the reference's);
  * rank r keeps labeled[r*lbs:(r+1)*lbs] + unlabeled[r*ubs:(r+1)*ubs] of each global batch —
    labeled-first inside the rank, which is what every ``_train`` body slices on
    (e.g. ssl_mt.py:147-151) — so the union over ranks is exactly the reference's global batch and
    no sample is read twice.
"""

import numpy as np
from torch.utils.data.sampler import Sampler


class TwoStreamBatchSampler(Sampler):
    """pixelssl/nn/data.py:126-177 with ``rank`` / ``world_size``.

    ``labeled_batch_size`` / ``unlabeled_batch_size`` are PER-RANK sizes (the reference's values divided
    by its GPU count).  An 'epoch' goes through the longer index list once; the shorter one is
    re-shuffled as often as needed.  All ranks must seed ``np.random`` identically (or pass the same
    ``seed``) so they agree on the permutations.
    """

    def __init__(self, labeled_idxs, unlabeled_idxs, labeled_batch_size, unlabeled_batch_size,
                 rank=None, world_size=None, seed=None):
        if rank is None or world_size is None:
            rank, world_size = _dist_rank_world()
        if not 0 <= rank < world_size:
            raise ValueError('rank %r out of range for world_size %r' % (rank, world_size))
        self.labeled_idxs = labeled_idxs
        self.unlabeled_idxs = unlabeled_idxs
        self.labeled_batch_size = labeled_batch_size
        self.unlabeled_batch_size = unlabeled_batch_size
        self.rank, self.world_size = rank, world_size
        self.global_labeled_batch_size = labeled_batch_size * world_size
        self.global_unlabeled_batch_size = unlabeled_batch_size * world_size
        self._rng = np.random if seed is None else np.random.RandomState(seed)

        assert len(self.labeled_idxs) >= self.global_labeled_batch_size > 0
        assert len(self.unlabeled_idxs) >= self.global_unlabeled_batch_size > 0

        self.unlabeled_batchs = len(self.unlabeled_idxs) // self.global_unlabeled_batch_size
        self.labeled_batchs = len(self.labeled_idxs) // self.global_labeled_batch_size

    def __iter__(self):
        # the stream that defines the epoch is shuffled eagerly (like the reference's ``iterate_once``),
        # the other one lazily, permutation by permutation, as elements are requested
        unlabeled_once = self.unlabeled_batchs >= self.labeled_batchs
        once = self._rng.permutation(self.unlabeled_idxs if unlabeled_once else self.labeled_idxs)
        return self._rank_batches(once, unlabeled_once)

    def __len__(self):
        return max(self.unlabeled_batchs, self.labeled_batchs)

    def _rank_batches(self, once, unlabeled_once):
        r, lbs, ubs = self.rank, self.labeled_batch_size, self.unlabeled_batch_size
        endless = _Reshuffler(self._rng, self.labeled_idxs if unlabeled_once else self.unlabeled_idxs)
        n_once = self.global_unlabeled_batch_size if unlabeled_once else self.global_labeled_batch_size
        n_endless = self.global_labeled_batch_size if unlabeled_once else self.global_unlabeled_batch_size
        pos = 0
        while True:
            if unlabeled_once:
                # the reference's zip() asks the labeled stream first, also for the request that ends
                # the epoch; kept so that the np.random stream stays identical across epochs
                labeled = endless.take(n_endless)
                if pos + n_once > len(once):
                    return
                unlabeled = once[pos:pos + n_once]
            else:
                if pos + n_once > len(once):
                    return
                labeled = once[pos:pos + n_once]
                unlabeled = endless.take(n_endless)
            pos += n_once
            yield list(labeled[r * lbs:(r + 1) * lbs]) + list(unlabeled[r * ubs:(r + 1) * ubs])


class _Reshuffler:
    """Endless stream over ``indices``: a fresh permutation is drawn only when an element is needed
    and the previous permutation is used up (``iterate_eternally``, pixelssl/nn/data.py:167-172)."""

    def __init__(self, rng, indices):
        self._rng, self._indices = rng, indices
        self._buf, self._pos = (), 0

    def take(self, n):
        out = []
        while len(out) < n:
            if self._pos >= len(self._buf):
                self._buf, self._pos = self._rng.permutation(self._indices), 0
            k = min(n - len(out), len(self._buf) - self._pos)
            out.extend(self._buf[self._pos:self._pos + k])
            self._pos += k
        return out


def _dist_rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
