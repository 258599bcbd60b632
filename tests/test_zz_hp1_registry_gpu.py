"""HP1 step kernel on the GPU against 13 more fixtures recorded from the REFERENCE'S OWN BaseMultirotor.step
(tests/golden/make_golden_registry.py): magpie, x500, lmf1, lmf2, tinyprop, base_random (eight arbitrarily rotated rotors),
morphy_stiff, octarotor; the velocity-steering-angle controller.  The spec is built by the product's registries / config mirror / URDF
pipeline.  Written after the round's GPU budget was spent: the same text passes on the host shadow (tests/test_hp1_shadow_cpu.py)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from aerial_gym_simulator_b200.hp1 import Hp1Engine
from tests import _hp1_common as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REG_FILES = sorted(glob.glob(os.path.join(GOLD, "hp1_regstep_*.npz")))


@pytest.mark.parametrize("path", REG_FILES, ids=[os.path.basename(p)[12:-4] for p in REG_FILES])
def test_step_kernel_matches_reference_golden_registry_specs(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    spec = H.spec_from_registry(meta["robot"], meta["controller"])
    eng = Hp1Engine(spec, meta["N"], DEV, debug_wrench=True, per_env_params="all")  # the fixtures randomise gains / motor constants
    H.check_engine_against_step_fixture(eng, spec, H.oracle_model_from_spec(spec), z, meta, sync=torch.cuda.synchronize)
    eng.close()
