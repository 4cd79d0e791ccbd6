"""Ensemble training step on the GPU against oracle/train_oracle.py.

Tolerances (fp32 on both sides, different summation order): loss and gradients 1e-4 relative
(+1e-7 absolute for gradients that are sums of near-cancelling terms); Adam slots and weights
after one step 1e-4 relative; after a run of steps the weights are compared with an absolute
bound of 4 * lr, because m / (sqrt(v) + eps) turns ulp-level differences of a tiny gradient
into sign-level differences of its update."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as T
from simba_b200.models import MlpEnsemble, TransitionModel

pytestmark = pytest.mark.gpu


def _ensemble(n_in, n_out, members, batch, layers, units, lr=0.00025, schedule=False, steps=5000,
              epochs=1, seed=0):
    return MlpEnsemble(n_in, n_out, members, batch_size=batch, learning_rate=lr,
                       learning_rate_schedule=schedule, training_steps=steps, train_epochs=epochs,
                       mlp_params=dict(n_layers=layers, units=units), seed=seed)


def _oracle(ens):
    return T.EnsembleTrainer([m.get_weights() for m in ens.ensemble], batch_size=ens.batch_size,
                             learning_rate=ens.learning_rate,
                             learning_rate_schedule=ens.learning_rate_schedule,
                             training_steps=ens.training_steps, train_epochs=ens.train_epochs)


def _data(rng, members, rows, n_in, n_out, scale=1.0):
    x = rng.uniform(0, 1, (members, rows, n_in)).astype(np.float32)
    y = (rng.normal(0, 0.1, (members, rows, n_out)) * scale).astype(np.float32)
    return x, y


@pytest.mark.parametrize('shape', [(62, 60, 5, 64, 4, 128), (12, 10, 3, 16, 2, 40), (7, 5, 2, 9, 1, 33),
                                   (62, 60, 2, 32, 2, 400), (20, 18, 2, 24, 3, 200), (9, 7, 1, 5, 2, 130)])
def test_one_step_matches_oracle(shape):
    n_in, n_out, E, B, L, U = shape
    ens = _ensemble(n_in, n_out, E, B, L, U)
    ora = _oracle(ens)
    x, y = _data(np.random.default_rng(1), E, B, n_in, n_out)
    loss = float(ens.training_step(x, y))
    want = float(ora.training_step(x, y))
    assert loss == pytest.approx(want, rel=1e-4)
    n_var = 2 * L + 4
    for e in range(E):
        for got, ref in zip(ens.trainer_arrays('grads', e), ora.last_grads[e * n_var:(e + 1) * n_var]):
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-7)
        for got, ref in zip(ens.trainer_arrays('m', e), ora.optimizer.m[e * n_var:(e + 1) * n_var]):
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-8)
        for got, ref in zip(ens.trainer_arrays('weights', e), ora.nets[e].arrays):
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=ens.learning_rate * 0.02)
    assert ens.iterations == 1


def test_clipvalue_and_schedule_over_several_steps():
    """Large targets push |g| past clipvalue 1.0; steps_per_epoch 3 of 4 epochs walks the
    EpochLearningRateSchedule (mlp_ensemble.py:80-83) down to zero."""
    n_in, n_out, E, B, L, U = 12, 10, 3, 16, 2, 40
    ens = _ensemble(n_in, n_out, E, B, L, U, lr=1e-3, schedule=True, steps=3, epochs=4)
    ora = _oracle(ens)
    rng = np.random.default_rng(2)
    clipped = False
    for s in range(14):
        x, y = _data(rng, E, B, n_in, n_out, scale=300.0 if s < 4 else 1.0)
        got = float(ens.training_step(x, y))
        want = float(ora.training_step(x, y))
        clipped |= any(np.abs(g).max() > 1.0 for g in ora.last_grads)
        assert got == pytest.approx(want, rel=2e-4), s
    assert clipped
    for e in range(E):
        for got, ref in zip(ens.trainer_arrays('weights', e), ora.nets[e].arrays):
            np.testing.assert_allclose(got, ref, rtol=0, atol=4e-3)
    # after 12 steps the schedule is at 0: the last two steps must not have moved anything
    before = [a.copy() for a in ens.trainer_arrays('weights', 0)]
    x, y = _data(rng, E, B, n_in, n_out)
    ens.training_step(x, y)
    for a, b in zip(before, ens.trainer_arrays('weights', 0)):
        np.testing.assert_array_equal(a, b)
    assert ens.iterations == 15


def test_fit_loop_with_uneven_batches_matches_oracle():
    n_in, n_out, E, B, L, U = 12, 10, 3, 16, 2, 40
    ens = _ensemble(n_in, n_out, E, B, L, U, lr=5e-4, steps=25)
    ens.validation_split = 0.0
    ora = _oracle(ens)
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1, (50, n_in)).astype(np.float32)          # 50 rows / 16 -> batches 13,13,12,12
    y = rng.normal(0, 0.1, (50, n_out)).astype(np.float32)
    np.random.seed(11)
    losses = ens.fit(x, y)
    np.random.seed(11)
    train_idx, _ = ens._split_indices(50)
    index, rows = ens.batch_schedule(50, 25)
    assert sorted(set(rows.tolist())) == [12, 13]
    want = ora.fit_batches(x, y, [train_idx[index[s, :, :rows[s]]] for s in range(25)])
    np.testing.assert_allclose(losses, want, rtol=2e-4)
    for e in range(E):
        for got, ref in zip(ens.ensemble[e].get_weights(), ora.nets[e].arrays):
            np.testing.assert_allclose(got, ref, rtol=0, atol=4 * 5e-4)
    assert ens.iterations == 25


def test_validation_step_chunks_and_matches_oracle():
    n_in, n_out, E, L, U = 12, 10, 3, 2, 40
    ens = _ensemble(n_in, n_out, E, 16, L, U)
    ora = _oracle(ens)
    rng = np.random.default_rng(4)
    for rows in (5, 64, MlpEnsemble.EVAL_ROWS + 777):
        x = rng.uniform(0, 1, (rows, n_in)).astype(np.float32)
        y = rng.normal(0, 0.1, (rows, n_out)).astype(np.float32)
        assert float(ens.validation_step(x, y)) == pytest.approx(float(ora.validation_step(x, y)), rel=1e-4)


def test_fit_learns_and_planner_sees_the_new_weights():
    """TransitionModel.fit (transition_model.py:34-40) on a linear system, then the trained
    weights must reach forward() (model handle re-commit) and ensemble[e].get_weights()."""
    from simba_b200.spaces import Box
    O, A = 6, 2
    obs_space, act_space = Box([-2.0] * O, [2.0] * O), Box([-1.0] * A, [1.0] * A)
    tm = TransitionModel('mlp_ensemble', obs_space, act_space, True, True, ensemble_size=3,
                         batch_size=32, learning_rate=2e-3, learning_rate_schedule=False,
                         training_steps=300, mlp_params=dict(n_layers=2, units=64), train_epochs=1)
    rng = np.random.default_rng(5)
    obs = rng.uniform(-1, 1, (600, O)).astype(np.float32)
    act = rng.uniform(-1, 1, (600, A)).astype(np.float32)
    nxt = obs + 0.1 * np.tanh(obs) + 0.05 * np.pad(act, ((0, 0), (0, O - A)))
    inputs = np.concatenate([obs, act], axis=1)
    w_before = tm.model.ensemble[0].get_weights()
    mu_before, _ = tm.model.forward(tm.scale(inputs[:6]))
    np.random.seed(0)
    losses = tm.fit(inputs, nxt)
    assert losses.shape == (300,) and np.isfinite(losses).all()
    assert losses[-20:].mean() < losses[:20].mean() - 0.5
    assert len(tm.model.validation_losses) == 10
    assert tm.model.validation_losses[-1][1] < tm.model.validation_losses[0][1]
    w_after = tm.model.ensemble[0].get_weights()
    assert any(np.abs(a - b).max() > 1e-3 for a, b in zip(w_before, w_after))
    mu_after, var_after = tm.model.forward(tm.scale(inputs[:6]))
    assert np.abs(mu_after - mu_before).max() > 1e-3
    ora = T.MemberNet(w_after)
    mu_ref, var_ref, _ = ora.forward(tm.scale(inputs[:2]))            # rows 0-1 belong to member 0
    np.testing.assert_allclose(mu_after[:2], mu_ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(var_after[:2], var_ref, rtol=1e-4, atol=1e-7)
    # prediction error on the deltas went down to the noise floor of this toy system
    err = np.abs(tm.predict(inputs[:90])[:, 1, :] - nxt[:90]).mean()
    assert err < 0.05


def test_training_twice_gives_identical_results():
    n_in, n_out, E, B, L, U = 12, 10, 3, 16, 2, 40
    rng = np.random.default_rng(6)
    x, y = _data(rng, E, B, n_in, n_out)
    out = []
    for _ in range(2):
        ens = _ensemble(n_in, n_out, E, B, L, U, seed=3)
        for _ in range(5):
            ens.training_step(x, y)
        out.append(ens.trainer_arrays('weights', 1))
    for a, b in zip(*out):
        np.testing.assert_array_equal(a, b)


from hypothesis import HealthCheck, assume, given, settings  # noqa: E402
from hypothesis import strategies as st                      # noqa: E402


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(n_out=st.integers(1, 70), n_act=st.integers(1, 6), E=st.integers(1, 6), B=st.integers(1, 70),
       rows_frac=st.floats(0.0, 1.0), L=st.integers(1, 5), U=st.integers(1, 300),
       steps=st.integers(1, 3), seed=st.integers(0, 2 ** 31 - 1))
def test_training_step_random_shapes(n_out, n_act, E, B, rows_frac, L, U, steps, seed):
    """Random layer widths (aligned and unaligned, one and several weight chunks), batch sizes that
    do not fill a row tile, fewer rows than batch_size: loss and weights follow the oracle."""
    n_in = n_out + n_act
    rows = max(1, int(round(rows_frac * B)))
    ens = _ensemble(n_in, n_out, E, B, L, U, lr=1e-3, seed=seed % 1000)
    ora = _oracle(ens)
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        x, y = _data(rng, E, rows, n_in, n_out)
        # a hidden pre-activation within rounding distance of the ReLU kink may be masked differently
        # by two correct fp32 implementations (and then a whole unit's gradient differs): skip those
        for e, net in enumerate(ora.nets):
            h = x[e]
            for l in range(L):
                z = h @ net.arrays[2 * l] + net.arrays[2 * l + 1]
                assume(np.abs(z).min() > 1e-5)
                h = np.maximum(z, 0)
        got, want = float(ens.training_step(x, y)), float(ora.training_step(x, y))
        assert got == pytest.approx(want, rel=2e-4, abs=1e-6)
    n_var = 2 * L + 4
    for e in range(E):
        for g, r in zip(ens.trainer_arrays('grads', e), ora.last_grads[e * n_var:(e + 1) * n_var]):
            np.testing.assert_allclose(g, r, rtol=2e-3, atol=2e-6)
        for g, r in zip(ens.trainer_arrays('weights', e), ora.nets[e].arrays):
            np.testing.assert_allclose(g, r, rtol=0, atol=4e-3 * steps)
    assert ens.iterations == steps


@pytest.mark.parametrize('shape,rate', [((62, 60, 3, 64, 4, 128), 0.1), ((12, 10, 2, 16, 2, 40), 0.5),
                                        ((20, 18, 2, 24, 3, 200), 0.25)])
def test_dropout_training_steps_match_oracle(shape, rate):
    """dropout_rate > 0 (mlp_ensemble.py:17-22, training=True): the device draws the keep masks from
    Philox stream 4; the oracle is fed oracle/philox.dropout_keep's restatement of the same counters.
    Loss and weights over three steps within the training tolerance; validation_step (training=False)
    is unaffected by the rate."""
    n_in, n_out, E, B, L, U = shape
    ens = MlpEnsemble(n_in, n_out, E, batch_size=B, learning_rate=1e-3, learning_rate_schedule=False,
                      mlp_params=dict(n_layers=L, units=U, dropout_rate=rate), seed=7)
    ora = T.EnsembleTrainer([m.get_weights() for m in ens.ensemble], batch_size=B, learning_rate=1e-3,
                            learning_rate_schedule=False, dropout_rate=rate, dropout_seed=7)
    rng = np.random.default_rng(12)
    x, y = _data(rng, E, B, n_in, n_out)
    plain = _ensemble(n_in, n_out, E, B, L, U, seed=7)
    assert float(ens.validation_step(x[0], y[0])) == pytest.approx(float(plain.validation_step(x[0], y[0])), rel=1e-6)
    first = None
    for s in range(3):
        x, y = _data(rng, E, B, n_in, n_out)
        got, want = float(ens.training_step(x, y)), float(ora.training_step(x, y))
        assert got == pytest.approx(want, rel=2e-4, abs=1e-6), s
        first = got if first is None else first
    n_var = 2 * L + 4
    for e in range(E):
        for g, r in zip(ens.trainer_arrays('grads', e), ora.last_grads[e * n_var:(e + 1) * n_var]):
            np.testing.assert_allclose(g, r, rtol=2e-3, atol=2e-6)
        for g, r in zip(ens.trainer_arrays('weights', e), ora.nets[e].arrays):
            np.testing.assert_allclose(g, r, rtol=0, atol=1.2e-2)
    # dropout changes the step: the same data without dropout gives a different loss
    assert abs(float(plain.training_step(x, y)) - first) > 1e-4
