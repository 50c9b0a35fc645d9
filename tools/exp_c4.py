"""Developer tool: config-4 timings (device resident, CUDA events) for the lengths served by the 13*11*7 kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_suite  # noqa: F401  (its argv[1] selects cases)
