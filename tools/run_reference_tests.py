"""Runs the reference's OWN test files against the drop-in package -- the ones that need neither jax nor pgx:
tests/ppo/test_rollout_buffer.py, tests/ppo/test_data_loader.py, tests/test_running_stats_vec.py (59 tests).

The files are never added to this repository: `--stage` copies them from the reference checkout into the
git-ignored scratch directory refcheck/ (which travels to the GPU box with the gpurun snapshot), and a plain run
executes whatever is staged there with `src.*` resolving to 2048-ppo-agent_b200/src, the re-export of this package.

    python tools/run_reference_tests.py --stage /root/reference      # here, where the reference is checked out
    gpurun -- python tools/run_reference_tests.py                     # on a B200
"""
import argparse
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SCRATCH = ROOT / "refcheck"
FILES = ("tests/ppo/test_rollout_buffer.py", "tests/ppo/test_data_loader.py", "tests/test_running_stats_vec.py")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", metavar="REFERENCE_ROOT", help="copy the test files from this checkout into refcheck/")
    args = ap.parse_args()
    if args.stage:
        for rel in FILES:
            dst = SCRATCH / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(Path(args.stage) / rel, dst)
        print(f"staged {len(FILES)} files under {SCRATCH} (git-ignored)")
        return 0
    if not (SCRATCH / "tests").is_dir():
        print("nothing staged: run with --stage <reference checkout> first")
        return 2
    env = dict(os.environ, PYTHONPATH=str(ROOT / "2048-ppo-agent_b200"))
    return subprocess.call([sys.executable, "-m", "pytest", "tests", "-q", "--no-header", "-p", "no:cacheprovider", "--tb=short"],
                           cwd=SCRATCH, env=env)


if __name__ == "__main__":
    sys.exit(main())
