"""Runs a few P=1 all-five-plugin cycles (for an ncu launch list of the latency path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from scheduler_plugins_b200 import engine as E, synth  # noqa: E402

print(bench.cycle_latency(E, synth, 0, 50_000, cycles=int(sys.argv[1]) if len(sys.argv) > 1 else 3))
