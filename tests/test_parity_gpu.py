"""GPU parity tests (-m gpu) of the SHIPPED precision at the SHIPPED configurations (BASELINE.json configs[2], [3], [4]) against
the fp64 oracle, under the one tolerance table of oracle/parity.py:

  C3  the bench batch itself: 64 Oxford-shape clouds x 16384 points (synth.make_batch(64, 16384, seed0=1000)), 512 clusters x 64,
      seed-0 weights in the TF random-init state (what bench.py times) and with randomised BN statistics;
  C4  the training batch shape (6 triplets = 18 clouds x 4096 points, 512 clusters x 64), forward in eval mode;
  C5  one KITTI-shape scan: 131 072 points, 1024 clusters.

Both device paths are held to their own row of the table: "bf16x3" (tcgen05, the default and the one the bench measures) and "fp32"
(CUDA-core FFMA).  FPS / ball-query indices must equal the oracle's bit for bit.  Matches models/feat3dnet.py:258-313.
"""
import numpy as np
import pytest
import torch

from oracle import net as onet
from oracle import parity
from tests.conftest import pkg
from tests.test_model_gpu import run_pipeline

pytestmark = pytest.mark.gpu

CONFIGS = {"C3": dict(B=64, N=16384, M=512, kind="oxford", seed0=1000),
           "C4": dict(B=18, N=4096, M=512, kind="oxford", seed0=2000),
           "C5": dict(B=1, N=131072, M=1024, kind="kitti", seed0=3000)}


@pytest.mark.parametrize("randomize_bn", [False, True])
@pytest.mark.parametrize("config", ["C3", "C4", "C5"])
def test_shipped_precisions_at_shipped_configs_vs_fp64_oracle(cuda, config, randomize_bn):
    cfg = CONFIGS[config]
    xyz = pkg("synth").make_batch(cfg["B"], cfg["N"], seed0=cfg["seed0"], kind=cfg["kind"])
    params = onet.init_params(seed=0, randomize_bn=randomize_bn)
    P64 = onet.to_torch(params, torch.float64)
    ref = parity.oracle_forward(onet, xyz, P64, cfg["M"])
    for precision in ("bf16x3", "fp32"):
        out, _ = run_pipeline(xyz, params, cfg["M"], precision=precision)
        assert np.array_equal(out["fps_idx"].cpu().numpy(), ref["fps_idx"]), "FPS indices differ from the oracle"
        assert np.array_equal(out["idx"].cpu().numpy(), ref["idx"]), "ball-query indices differ from the oracle"
        assert np.array_equal(out["pts_cnt"].cpu().numpy(), ref["pts_cnt"])
        assert np.array_equal(out["xyz"].cpu().numpy(), ref["xyz"])
        own = parity.oracle_descriptor_at(onet, xyz, P64, ref["xyz"], out["orientation"])
        e = parity.check(parity.errors(out, ref, own), precision, "%s %s rb=%d" % (config, precision, randomize_bn))
        print(config, precision, "rb=%d" % randomize_bn, {k: "%.2e" % v for k, v in e.items()})
