"""N>1 host logic on the CPU: the engine's model-parallel scheme restated over the oracle's kernels with gloo
collectives (tests/mp_gloo_worker.py), world sizes 2 and 3 (3 = uneven unit ranges), against the single-process oracle."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,port", [(2, 29611), (3, 29612)])
def test_sharded_training_equals_single_process(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mp_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert res["world"] == world and res["outgoing"] == [False, True, True, True]
    for got, want in zip(res["losses"], res["want"]):
        assert abs(got - want) <= 1e-5 * abs(want)
    assert max(res["errs"]) < 1e-5, res["errs"]
