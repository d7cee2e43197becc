"""Time-major GAE kernel alone (preallocated outputs, L2 flushed before each launch).
python tools/probes/tm_bench.py [variant.so ...]: also times other builds of the library (tools/probes/build_gae4_variant.sh)."""
import ctypes as C
import sys
sys.path[:0] = ["/root/repo", "/root/repo/2048-ppo-agent_b200"]
import torch
from g2048 import engine as E
from g2048 import _native as N


class _Lib:  # same call convention as g2048._native for one entry point of another build
    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.g2048_gae_time_major.restype = C.c_int
        self.lib.g2048_gae_time_major.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_int64, C.c_void_p, C.c_double, C.c_double] + [C.c_void_p] * 4

    def call(self, name, *args):
        assert getattr(self.lib, name)(*args) == 0


for path in [None] + sys.argv[1:]:
  NN = N if path is None else _Lib(path)
  print(path or "libg2048.so")
  for t_steps, b in ((128, 1 << 16), (128, 1 << 18), (37, 3000)):
      rr = torch.rand((t_steps, b), device="cuda"); vv = torch.rand((t_steps, b), device="cuda")
      mm = ((torch.rand((t_steps, b), device="cuda") < 1 / 300).to(torch.uint8) << 6)
      flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
      adv = torch.empty_like(rr); ret = torch.empty_like(rr); mom = torch.zeros(6, dtype=torch.float64, device="cuda")
      run = lambda: NN.call("g2048_gae_time_major", N.ptr(rr), N.ptr(vv), N.ptr(mm), t_steps, b, None, 0.99, 0.95, N.ptr(adv), N.ptr(ret), N.ptr(mom), N.stream_ptr())
      run()
      ts = []
      for _ in range(8):
          flush.fill_(1)
          a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record(); run(); c.record(); torch.cuda.synchronize()
          ts.append(a.elapsed_time(c) * 1e3)
      ts.sort()
      print(t_steps, b, "median us", round(ts[4], 1), "GB/s", round(t_steps * b * 17 / ts[4] / 1e3))
