"""jax.random-style keys on the device: ``key(seed)``, ``split(key, n)`` and the runner's chain.

A key is two uint32 words, stored in an int32 tensor of shape (2,) (a batch of keys: (n, 2)).
Replaces jax.random.key / jax.random.split as used at src/runs/batch_runner.py:32,105-106,118-119.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N
from . import engine as E


def key(seed: int, device=None) -> torch.Tensor:
    """jax.random.key(seed)."""
    device = N.require_cuda() if device is None else device
    return E.words_tensor(list(E.key_words(seed)), device)


def as_key_tensor(k, device=None) -> torch.Tensor:
    """Accepts a tensor / numpy array / sequence of uint32 words; returns an int32 device tensor."""
    device = N.require_cuda() if device is None else device
    if isinstance(k, torch.Tensor):
        if k.dtype != torch.int32:
            k = torch.from_numpy(k.detach().cpu().numpy().astype(np.uint32).view(np.int32))
        return k.to(device).contiguous()
    return E.words_tensor(np.asarray(k, dtype=np.uint64).astype(np.uint32), device)


def split(k, num: int = 2, rng_mode=None) -> torch.Tensor:
    """jax.random.split(key, num) -> (num, 2)."""
    k = as_key_tensor(k)
    return E.split_keys(k, num, 0, num, E.resolve_rng_mode(rng_mode))


_SIDE_STREAMS: dict = {}  # device index -> the stream every KeyChain of that device generates on


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    """A stream is not free the first time it is used (the driver sets it up at its first launch, a millisecond or
    two), and torch hands out a different pooled stream per request: one per device, created once, serves all."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    stream = _SIDE_STREAMS.get(index)
    if stream is None:
        stream = _SIDE_STREAMS[index] = torch.cuda.Stream(device=device)
    return stream


class KeyChain:
    """The runner's ``key, sub = split(key)`` chain, generated ahead of time on the device.

    Sub keys are produced in blocks by one small kernel (the chain is sequential by
    construction) and consumed by the rollout kernels straight from device memory, so a run
    never waits on the host for keys.  ``peek(n)`` returns the next n sub keys without
    consuming them; ``consume(n)`` advances (a run consumes 1 + 2*T of them, T known only
    after the run).  ``key`` is the chain key the reference's runner would hold now.
    """

    COLD_BLOCK = 2048  # least number generated when a peek has to wait for them (1 + 2 * 1 023 loop steps: ~0.3 ms)
    TOP_UP = 256       # least number generated behind a consume

    def __init__(self, seed_or_key, rng_mode: int, device=None):
        self.device = N.require_cuda() if device is None else torch.device(device)
        self.rng_mode = rng_mode
        if isinstance(seed_or_key, (int, np.integer)):
            words = np.array(E.key_words(int(seed_or_key)), dtype=np.uint32)
        else:
            words = np.asarray(seed_or_key, dtype=np.uint32).reshape(2).copy()
        self._base_key = words  # chain key at absolute position self._base_pos - self._replay
        self._replay = 0  # splits consumed since _base_key was last materialised (see `key`)
        self._base_pos = 0  # number of splits consumed so far
        self._subs = torch.empty((0, 2), dtype=torch.int32, device=self.device)  # subs from _base_pos on
        self._tip = E.words_tensor(words, self.device)  # chain key after all generated subs
        # Generation runs on a stream of its own (one per device, shared by every chain): the chain is one thread
        # working through ~0.14 us per split -- for a batch of a thousand envs that is a fifth of the play kernel's own
        # time -- so the keys a batch consumes are replaced while the next batches run: every consume(k) queues the
        # generation of k more, two consumes ahead of the peek that will need them, and only a chain that starts cold is
        # waited for (plus a quarter more, queued behind it, that covers the first consume).  _ahead: blocks generated
        # (or being generated) but not yet joined to _subs, oldest first, each with the event that marks it complete.
        self._side = None
        self._window = 0  # the largest peek since the last consume: what the next batch will want to see again
        self._ahead: list[tuple[torch.Tensor, torch.cuda.Event]] = []

    def _generate(self, count: int) -> None:
        """Queue `count` more sub keys on the side stream (ordered behind every earlier block: they share _tip)."""
        if self._side is None:
            self._side = _side_stream(self.device)
            self._side.wait_stream(torch.cuda.current_stream(self.device))  # _tip was written on the caller's stream
            # ... and from now on is read and written by kernels of the side stream only: the allocator must not hand
            # its memory to somebody else while one of them is still running (a chain that is dropped right after its
            # last batch leaves a generation in flight)
            self._tip.record_stream(self._side)
        with torch.cuda.stream(self._side):
            block = E.chain_advance(self._tip, self.rng_mode, count)
            done = torch.cuda.Event()
            done.record(self._side)
        self._ahead.append((block, done))

    def _pending(self) -> int:
        return sum(block.shape[0] for block, _ in self._ahead)

    def peek(self, n: int) -> torch.Tensor:
        capturing = torch.cuda.is_current_stream_capturing()
        current = torch.cuda.current_stream(self.device)
        self._window = max(self._window, n)
        while n > self._subs.shape[0]:
            if not self._ahead:
                if capturing:
                    raise RuntimeError("KeyChain.peek inside a CUDA graph capture needs keys that were not generated yet: peek them before the capture")
                count = max(n - self._subs.shape[0], self.COLD_BLOCK)
                self._generate(count)       # the caller waits for this one ...
                self._generate(count // 4)  # ... and this one is ready by the time it has consumed some
            block, done = self._ahead.pop(0)
            current.wait_event(done)
            block.record_stream(current)
            self._subs = torch.cat([self._subs, block]) if self._subs.shape[0] else block
        return self._subs[:n]

    def consume(self, n: int) -> None:
        """Advance by n splits.  Nothing runs on the caller's stream: the sub keys were generated ahead, the chain key
        itself (``key``) is only replayed when somebody asks for it, and the keys that replace the consumed ones are
        queued on the generation stream."""
        if n <= 0:
            return
        window = self._window
        self.peek(n)
        self._replay += n
        self._base_pos += n
        self._subs = self._subs[n:]
        self._window = 0
        # top up behind the consumer, two consumes deep: the block the next batch has to join was then queued a whole
        # batch ago and is complete (one deep, it would have been queued microseconds before that batch asks for it)
        short = window + 2 * n - (self._subs.shape[0] + self._pending())
        if short > 0 and not torch.cuda.is_current_stream_capturing():
            self._generate(max(short, self.TOP_UP))

    def next_sub(self) -> torch.Tensor:
        sub = self.peek(1)[0].clone()
        self.consume(1)
        return sub

    @property
    def key(self) -> np.ndarray:
        """The chain key after everything consumed so far (the reference runner's ``self.key``): replayed from the last
        materialised key over the pending splits -- one small kernel and one 8-byte read, off the rollout's path."""
        if self._replay:
            base = E.words_tensor(self._base_key, self.device)
            E.chain_advance(base, self.rng_mode, self._replay)
            self._base_key = E.words_numpy(base).copy()
            self._replay = 0
        return self._base_key.copy()

    @property
    def position(self) -> int:
        return self._base_pos
