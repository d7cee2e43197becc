"""A/B timing of builds of libg2048.so on ONE box: every library given on the command line (default: the shipped one
plus tools/ab/*.so) runs the same play workloads in its own process (G2048_LIB selects the library).

    make -C 2048-ppo-agent_b200/csrc OUT=../../tools/ab/libg2048_x.so BUILD=build_x EXTRA=-DG2048_PLAY3_TAIL=0
    python tools/ab_variants.py            # on the GPU box; prints one JSON line per library
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def one():
    sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]
    import torch

    from g2048 import engine as E

    dev = torch.device("cuda", 0)
    subs = E.chain_advance(E.words_tensor([0, 2048], dev), 1, 1 + 2 * 2048)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts), sum(ts) / len(ts)

    for policy, pname in ((0, "random"), (1, "drul")):
        for logn, reps in ((18, 7), (21, 5), (24, 3)):
            if policy == 1 and logn == 24:
                continue
            n = 1 << logn
            st = {}
            best, mean = timed(lambda: st.update(s=E.play(policy, subs, n, 0, n, 1, per_env=False)["stats"]), reps)
            steps = E.play_stats_dict(st["s"])["env_steps"]
            out[f"{pname}_2^{logn}"] = {"ms_min": round(best, 4), "ms_mean": round(mean, 4), "G_env_steps_per_s": round(steps / best / 1e6, 3)}
        n = 1 << 18
        best, mean = timed(lambda: E.play_record(policy, subs, n, 0, n, 1), 5)
        out[f"{pname}_record_2^18"] = {"ms_min": round(best, 4), "ms_mean": round(mean, 4)}
    # the opt-in GAE scan at the bench's size
    from g2048 import _native as N

    n_g = 1 << 26
    r, v = torch.rand(n_g, device=dev), torch.rand(n_g, device=dev)
    d = (torch.rand(n_g, device=dev) < 1 / 300).to(torch.uint8)
    adv, ret = torch.empty(n_g, device=dev), torch.empty(n_g, device=dev)
    scratch = torch.zeros(int(N.lib.g2048_gae_scan_scratch_bytes(n_g)), dtype=torch.uint8, device=dev)
    mom = torch.zeros(6, dtype=torch.float64, device=dev)
    best, mean = timed(lambda: N.call("g2048_gae_flat_scan", N.ptr(r), N.ptr(v), N.ptr(d), n_g, 0.99, 0.95, N.ptr(adv), N.ptr(ret),
                                      N.ptr(scratch), N.ptr(mom), N.stream_ptr()), 5)
    out["gae_scan_2^26"] = {"ms_min": round(best, 4), "ms_mean": round(mean, 4), "TBs": round(n_g * 17 / best / 1e9, 3)}
    print(json.dumps(out))


def main():
    libs = [Path(p) for p in sys.argv[1:]] or [ROOT / "2048-ppo-agent_b200" / "libg2048.so", *sorted((ROOT / "tools" / "ab").glob("*.so"))]
    for rnd in range(2):  # two rounds, interleaved, so that drift of the box shows
        for lib in libs:
            env = dict(os.environ, G2048_LIB=str(lib))
            res = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
            line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-500:]
            print(json.dumps({"lib": lib.name, "round": rnd, "result": json.loads(line) if line.startswith("{") else line}))


if __name__ == "__main__":
    one() if "--one" in sys.argv else main()
