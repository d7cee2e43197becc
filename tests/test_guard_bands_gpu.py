"""Out-of-bounds-write check with canaries (compute-sanitizer is not available on the GPU pool).

Every output tensor the host side allocates while tests/kernel_exercise.py runs every kernel at ragged sizes is
carved out of a larger byte buffer pre-filled with a canary; after the run both guard bands of every buffer must
still hold the canary.  This sees stray writes up to GUARD bytes either side of an output -- the tail of a
vectorised store, a bulk copy rounded up to its 16-byte granule, an off-by-one row.  It does not see stray reads
or far-away writes; result parity with the oracle (test_kernels_gpu.py) covers wrong reads.
"""
import math

import pytest
import torch

import kernel_exercise

pytestmark = pytest.mark.gpu

GUARD = 8192          # bytes either side; a multiple of 256 so the interior keeps cudaMalloc-like alignment
CANARY = 0xA5


class GuardedTorch:
    """Stands in for the `torch` module inside the host-side modules: empty/zeros/empty_like/zeros_like on a CUDA
    device return the interior of a canary-filled buffer; everything else is torch's own."""

    def __init__(self):
        self.raw = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _carve(self, shape, dtype, device, zero):
        if isinstance(shape, (int,)):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        if device is None or torch.device(device).type != "cuda":
            return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
        itemsize = torch.empty((), dtype=dtype).element_size()
        nbytes = math.prod(shape) * itemsize
        span = (nbytes + 255) // 256 * 256  # the band after the data starts right at its last byte + 1
        raw = torch.full((GUARD + span + GUARD,), CANARY, dtype=torch.uint8, device=device)
        self.raw.append((raw, nbytes))
        inner = raw[GUARD:GUARD + nbytes]
        if zero:
            inner.zero_()
        if dtype == torch.bool:
            if not zero:
                inner.fill_(1)  # a bool tensor must not hold bytes other than 0/1
            return inner.view(torch.bool).view(shape)
        return inner.view(dtype).view(shape)

    def empty(self, *shape, dtype=torch.float32, device=None, **kw):
        if kw or device is None or torch.device(device).type != "cuda":  # host tensors (pinned staging buffers ...)
            return torch.empty(*shape, dtype=dtype, device=device, **kw)
        return self._carve(shape[0] if len(shape) == 1 else shape, dtype, device, False)

    def zeros(self, *shape, dtype=torch.float32, device=None, **kw):
        if kw or device is None or torch.device(device).type != "cuda":
            return torch.zeros(*shape, dtype=dtype, device=device, **kw)
        return self._carve(shape[0] if len(shape) == 1 else shape, dtype, device, True)

    def empty_like(self, t):
        return self._carve(tuple(t.shape), t.dtype, t.device, False)

    def zeros_like(self, t):
        return self._carve(tuple(t.shape), t.dtype, t.device, True)

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for i, (raw, nbytes) in enumerate(self.raw):
            head = raw[:GUARD]
            tail = raw[GUARD + nbytes:]
            if not bool((head == CANARY).all()) or not bool((tail == CANARY).all()):
                first_tail = int((tail != CANARY).nonzero()[0]) if bool((tail != CANARY).any()) else None
                n_head = int((head != CANARY).sum())
                bad.append(f"buffer #{i} ({nbytes} B): {n_head} head bytes overwritten, first tail offset {first_tail}")
        return bad


def test_guarded_allocator_sees_a_stray_write():
    g = GuardedTorch()
    x = g.empty(10, dtype=torch.float32, device="cuda")
    assert x.shape == (10,) and x.data_ptr() % 256 == 0
    assert g.check() == []
    raw, nbytes = g.raw[0]
    raw[GUARD + nbytes] = 0  # one byte past the end
    assert len(g.check()) == 1


def test_no_kernel_writes_outside_its_outputs(monkeypatch):
    import g2048
    from g2048 import engine as E
    from g2048.ppo import data_loader, rollout_buffer
    from g2048.runs import batch_runner

    guard = GuardedTorch()
    for mod in (E, data_loader, rollout_buffer, batch_runner):
        monkeypatch.setattr(mod, "torch", guard)
    kernel_exercise.run_all(E, g2048, T=guard)
    assert len(guard.raw) > 300, "the host side did not allocate through the guarded allocator"
    assert guard.check() == []
