"""Runs any script of the repo (bench.py, tools/bench_configs.py, ...) on the CPU functional simulator:
    python tests/cpusim/run_sim.py bench.py --n 512 --steps 1 --warmup 1
TEST INFRASTRUCTURE: checks the scripts' own logic (argument handling, byte accounting, JSON contract) without a GPU.  The
numbers such a run prints are NOT measurements."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import simtorch  # noqa: E402

simtorch.install()
script = sys.argv[1]
sys.argv = sys.argv[1:] + os.environ.get("CPUSIM_ARGS", "").split()   # torchrun's own parser chokes on e.g. `--n` after the script
runpy.run_path(os.path.join(ROOT, script) if not os.path.isabs(script) else script, run_name="__main__")
