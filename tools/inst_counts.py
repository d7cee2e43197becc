"""Measured instruction counts of the shipped play kernels -> profiles/inst_counts.json (what bench.py's integer
roofline is computed from, instead of SURVEY 8(d)'s paper count).

On the GPU box (the plain run first, as the profiling recipe asks):

    python tools/inst_counts.py run gpurun_out/inst_counts_steps.json &&
    ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,\
sm__inst_executed_pipe_lsu.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --print-units base -k regex:play3_kernel --csv --log-file gpurun_out/inst_counts.csv \
        python tools/inst_counts.py run gpurun_out/inst_counts_steps_ncu.json

Here:  python tools/inst_counts.py parse gpurun_out/inst_counts.csv gpurun_out/inst_counts_steps.json profiles/inst_counts.json

`run` launches, in this order, the table kernel for: random policy 2^21 envs, DRUL policy 2^20 envs, and the
recording form of both at 2^18 envs; it writes the env-steps of each launch (from the statistics block) as JSON.
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "2048-ppo-agent_b200")]

LAUNCHES = [("random", 0, 1 << 21, False), ("drul", 1, 1 << 20, False), ("random_record", 0, 1 << 18, True),
            ("drul_record", 1, 1 << 18, True)]


def run(out_path):
    import torch

    from g2048 import engine as E

    dev = torch.device("cuda", 0)
    subs = E.chain_advance(E.words_tensor(list(E.key_words(2048)), dev), 1, 1 + 2 * 2048)
    E.play(0, subs, 40000, 0, 40000, 1, per_env=False)  # builds the row tables (not a play3 launch of interest: skipped below)
    torch.cuda.synchronize()
    rows = []
    for name, policy, n, record in LAUNCHES:
        if record:
            st = E.play_record(policy, subs, n, 0, n, 1)["stats"]
        else:
            st = E.play(policy, subs, n, 0, n, 1, per_env=False, entry="g2048_play_tables")["stats"]
        d = E.play_stats_dict(st)
        assert d["episodes"] == n and d["cut_short"] == 0, d
        rows.append({"name": name, "envs": n, "env_steps": d["env_steps"]})
    Path(out_path).write_text(json.dumps(rows))
    print(json.dumps(rows))


def parse(csv_path, steps_path, out_path):
    steps = json.loads(Path(steps_path).read_text())
    lines = [ln for ln in Path(csv_path).read_text().splitlines() if not ln.startswith("==")]
    recs = list(csv.DictReader(lines))
    by_id = {}
    for r in recs:  # long format: one row per (launch ID, metric)
        by_id.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    launches = [by_id[k] for k in sorted(by_id)][1:]  # the first play3 launch is the warm-up that builds the tables
    assert len(launches) == len(steps), (len(launches), len(steps))
    out = {}
    for lch, st in zip(launches, steps):
        ti, wi = lch["smsp__thread_inst_executed.sum"], lch["smsp__inst_executed.sum"]
        out[st["name"]] = {
            "thread_inst_per_env_step": ti / st["env_steps"], "warp_inst_per_env_step_x32": 32 * wi / st["env_steps"],
            "active_lanes_per_warp_inst": ti / wi, "alu_pipe_warp_inst_per_env_step_x32": 32 * lch.get("sm__inst_executed_pipe_alu.sum", 0) / st["env_steps"],
            "lsu_pipe_warp_inst_per_env_step_x32": 32 * lch.get("sm__inst_executed_pipe_lsu.sum", 0) / st["env_steps"],
            "dram_bytes_per_launch": lch.get("dram__bytes_read.sum", 0) + lch.get("dram__bytes_write.sum", 0),
            "dram_bytes_read": lch.get("dram__bytes_read.sum", 0), "dram_bytes_written": lch.get("dram__bytes_write.sum", 0),
            "envs": st["envs"], "env_steps": st["env_steps"], "ncu_duration_ms": lch.get("gpu__time_duration.sum", 0) / 1e6,
            "kernel": lch["kernel"][:80],
            "source": "ncu smsp__thread_inst_executed.sum / env-steps of the launch (tools/inst_counts.py, shipped build)",
        }
    Path(out_path).write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        parse(*sys.argv[2:5])
