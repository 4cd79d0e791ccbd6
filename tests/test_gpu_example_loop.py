"""The example training loop (examples/mbrl_loop.py): RandomMpc warm-up -> TransitionModel.fit ->
batched SafeCemMpc collection, all on libsimba_b200.so."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_mbrl_loop_runs_and_the_model_improves():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'mbrl_loop.py')
    spec = importlib.util.spec_from_file_location('mbrl_loop', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rep = mod.run(iterations=2, n_envs=4, warmup_steps=300, interaction_steps=120, episode_length=30,
                  training_steps=300, log=lambda s: None)
    assert len(rep) == 2
    for r in rep:
        assert np.isfinite(r['loss_last']) and r['loss_last'] < r['loss_first']
        assert np.isfinite(r['mean_return']) and r['env_steps'] >= 120
    # the first fit must beat the randomly initialised ensemble on held-in transitions
    assert rep[0]['one_step_error'] < 0.5 * rep[0]['one_step_error_before_fit']
