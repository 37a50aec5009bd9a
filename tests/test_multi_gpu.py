"""Model-parallel (hot-path rows a15 / 8e) parity on real GPUs: the engine on 2 (and 4) ranks with NCCL
reduce-scatter / all-gather at the layer boundaries must reproduce the single-process CPU oracle.
Skipped when the box has fewer GPUs than ranks."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_world(world, extra_env=None, port=29611):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=env, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("MP_RESULT ")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-4000:]
    return json.loads(lines[-1][len("MP_RESULT "):])


@pytest.mark.parametrize("p2p", [1, 0], ids=["peer-memory", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_model_parallel_sgd_matches_oracle(world, p2p):
    r = run_world(world, {"MP_MODE": "0", "MP_P2P": str(p2p)}, port=29611 + world + 10 * p2p)
    assert r["loss_err"] < 1e-5, r
    assert r["stream_loss_err"] < 1e-5, r                     # LoadSparseData while model parallel (the e2e path of bench.py at N > 1)
    for k, v in r["errs"].items():
        assert v < (5e-5 if k.startswith("b") else 1e-5), (k, v, r)


@pytest.mark.parametrize("p2p", [1, 0], ids=["peer-memory", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_model_parallel_momentum_odd_and_uneven_units(world, p2p):
    # world 2: odd local strides (65 | 65, 33 | 33, 1025 | 1025) -> scalar kernel paths, no 128-bit alignment;
    # world 4: 130 and 66 and 2050 do not divide by 4 -> uneven unit ranges (E/NNLayer.cpp:108-112): the peer-memory kernels
    # address them in place, the NCCL path takes its all-reduce + slice / grouped-broadcast fallbacks
    r = run_world(world, {"MP_MODE": "1", "MP_SIZES": "[2050, 130, 66, 130, 2050]", "MP_P2P": str(p2p), "MP_TOPK": "20"}, port=29631 + world + 10 * p2p)
    assert r["loss_err"] < 1e-5, r
    assert r["topk_ok"] is True, r
    for k, v in r["errs"].items():
        assert v < (5e-5 if k.startswith("b") else 1e-5), (k, v, r)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_model_parallel_config3_shape_tensor_core_path(world):
    """BASELINE config 3: the ML-20M-shape network sharded 2 / 4 / 8 ways (27,278 / 8 = 3,409.75: uneven shards), 3xTF32 mode, i.e. the
    fused output-layer forward and the streamed dW / dX kernels on every shard, peer-memory exchange with cached gathers."""
    r = run_world(world, {"MP_MODE": "0", "MP_SIZES": "[27278, 128, 128, 128, 27278]", "MP_BATCH": "1024", "MP_DATA": "ml20m", "MP_GEMM": "2",
                          "MP_STEPS": "2"}, port=29651 + world)
    assert r["loss_err"] < 3e-5, r
    for k, v in r["errs"].items():
        assert v < 5e-5, (k, v, r)


@pytest.mark.parametrize("world", [2, 4])
def test_model_parallel_topk_global_matches_single_process(world):
    r = run_world(world, {"MP_MODE": "0", "MP_TOPK": "50"}, port=29691 + world)
    assert r["topk_ok"] is True, r
