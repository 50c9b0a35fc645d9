"""The N > 1 host path on CPU: world_size-2 gloo, product kernels under the thread emulator (tools/sim).
Covers lot sharding (no collective) and the sharded 2-D transform (all-to-all transposes)."""
import os
import subprocess
import sys

import pytest

import fftlibs as fl


@pytest.mark.parametrize("world", [2])
def test_sharded_paths_world2_gloo(world):
    subprocess.check_call(["make", "-s", "-C", os.path.join(fl.ROOT, "tools", "sim")])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(fl.ROOT, "tests", "dist_worker.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
