"""The training-step oracle (oracle/train_oracle.py) pinned by calculus and hand-computed steps.

No TensorFlow here, so these are the known-answer tests that stand in for running the reference
(mlp_ensemble.py:64-67, :70-83, :113-117, :134-146)."""
import numpy as np
import pytest

from oracle import train_oracle as T


def _member(rng, n_in, n_out, units, layers, dtype):
    arrays = []
    fan = n_in
    for _ in range(layers):
        arrays += [rng.normal(0, 0.4, (fan, units)), rng.normal(0, 0.1, units)]
        fan = units
    for _ in range(2):
        arrays += [rng.normal(0, 0.4, (units, n_out)), rng.normal(0, 0.1, n_out)]
    return [a.astype(dtype) for a in arrays]


def test_nll_known_answer():
    y = np.zeros((1, 2))
    mu = np.array([[1.0, -2.0]])
    var = np.array([[1.0, 4.0]])
    want = 0.5 * (np.log(2 * np.pi) + np.log(8 * np.pi)) / 2 + 0.5 * (1.0 + 1.0) / 2
    assert T.negative_log_likelihood(y, mu, var) == pytest.approx(want, rel=1e-12)


def test_gradients_match_central_differences():
    rng = np.random.default_rng(0)
    net = T.MemberNet(_member(rng, 5, 3, 8, 2, np.float64), np.float64)
    x = rng.normal(size=(7, 5))
    y = rng.normal(size=(7, 3))
    loss, grads = net.loss_and_grads(x, y, 0.2)
    h = 1e-6
    for a, g in zip(net.arrays, grads):
        flat = a.reshape(-1)
        for j in rng.choice(flat.size, size=min(flat.size, 6), replace=False):
            keep = flat[j]
            flat[j] = keep + h
            up = T.negative_log_likelihood(y, *net.forward(x)[:2]) * 0.2
            flat[j] = keep - h
            dn = T.negative_log_likelihood(y, *net.forward(x)[:2]) * 0.2
            flat[j] = keep
            assert g.reshape(-1)[j] == pytest.approx((up - dn) / (2 * h), rel=2e-5, abs=1e-9)
    assert loss == pytest.approx(T.negative_log_likelihood(y, *net.forward(x)[:2]) * 0.2)


def test_schedule():
    # mlp_ensemble.py:80-83 with steps_per_epoch 10, 4 epochs
    assert T.lr_schedule(0, 1e-3, 10, 4) == np.float32(1e-3)
    assert T.lr_schedule(9, 1e-3, 10, 4) == np.float32(1e-3)
    assert T.lr_schedule(10, 1e-3, 10, 4) == pytest.approx(0.75e-3, rel=1e-6)
    assert T.lr_schedule(39, 1e-3, 10, 4) == pytest.approx(0.25e-3, rel=1e-6)
    assert T.lr_schedule(400, 1e-3, 10, 4) == 0.0
    assert T.lr_schedule(400, 1e-3, 10, 4, enabled=False) == np.float32(1e-3)


def test_adam_hand_computed_two_steps():
    opt = T.Adam([(3,)], 0.1, 100, 10, schedule=False, dtype=np.float64)
    w = [np.array([1.0, 1.0, 1.0])]
    g = [np.array([0.5, -3.0, 0.0])]                      # -3 is clipped to -1
    opt.apply(w, g)
    # t = 1: m = 0.1 g, v = 0.001 g^2, lr_t = 0.1 sqrt(0.001) / 0.1
    gc = np.array([0.5, -1.0, 0.0])
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = 1.0 - lr_t * (0.1 * gc) / (np.sqrt(0.001 * gc ** 2) + 1e-5)
    np.testing.assert_allclose(w[0], want, rtol=1e-7)      # lr0 is an fp32 constant
    opt.apply(w, g)
    m2 = 0.1 * gc + 0.9 * 0.1 * gc
    v2 = 0.001 * gc ** 2 + 0.999 * 0.001 * gc ** 2
    lr_t2 = 0.1 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    np.testing.assert_allclose(w[0], want - lr_t2 * m2 / (np.sqrt(v2) + 1e-5), rtol=1e-7)
    assert opt.iterations == 2


def test_training_reduces_loss_and_batches_cover_rows():
    rng = np.random.default_rng(1)
    members = [_member(rng, 4, 2, 16, 2, np.float32) for _ in range(3)]
    tr = T.EnsembleTrainer(members, batch_size=8, learning_rate=3e-3, learning_rate_schedule=False,
                           training_steps=60)
    x = rng.normal(size=(40, 4)).astype(np.float32)
    y = (x[:, :2] * 0.5 + 0.1).astype(np.float32)
    idx = T.make_batch_index(rng, 40, 3, 8, 60)
    assert len(idx) == 60 and idx[0].shape == (3, 8)
    first_pass = np.concatenate([b[0] for b in idx[:5]])
    assert sorted(first_pass.tolist()) == list(range(40))
    before = tr.validation_step(x, y)
    losses = tr.fit_batches(x, y, idx)
    assert np.isfinite(losses).all()
    assert tr.validation_step(x, y) < before


def test_uneven_batches_follow_array_split():
    idx = T.make_batch_index(np.random.default_rng(2), 10, 2, 4, 3)     # ceil(10/4) = 3 batches
    assert [b.shape[1] for b in idx] == [4, 3, 3]


def test_dropout_gradients_and_mask_statistics():
    """Dropout (training only): gradients with a given keep mask against central differences, and the
    Philox mask contract keeps a fraction 1 - rate of the units, independently per (member, layer, step)."""
    from oracle import philox
    rng = np.random.default_rng(3)
    net = T.MemberNet(_member(rng, 5, 3, 8, 2, np.float64), np.float64)
    x, y = rng.normal(size=(7, 5)), rng.normal(size=(7, 3))
    keep = [philox.dropout_keep(1, 0, 0, l, 7, 8, 0.3).astype(np.float64) / 0.7 for l in range(2)]
    loss, grads = net.loss_and_grads(x, y, 1.0, keep)
    h = 1e-6
    for a, g in zip(net.arrays, grads):
        flat = a.reshape(-1)
        for j in rng.choice(flat.size, size=min(flat.size, 5), replace=False):
            old = flat[j]
            flat[j] = old + h
            up = T.negative_log_likelihood(y, *net.forward(x, keep)[:2])
            flat[j] = old - h
            dn = T.negative_log_likelihood(y, *net.forward(x, keep)[:2])
            flat[j] = old
            assert g.reshape(-1)[j] == pytest.approx((up - dn) / (2 * h), rel=2e-5, abs=1e-9)
    m = philox.dropout_keep(9, 3, 1, 2, 256, 400, 0.25)
    assert m.shape == (256, 400) and abs(m.mean() - 0.75) < 0.01
    assert not np.array_equal(m, philox.dropout_keep(9, 4, 1, 2, 256, 400, 0.25))      # new mask every step
    assert not np.array_equal(m, philox.dropout_keep(9, 3, 0, 2, 256, 400, 0.25))      # and per member
    assert np.array_equal(philox.dropout_keep(9, 3, 1, 2, 256, 400, 0.0), np.ones((256, 400), bool))


def test_training_oracle_matches_committed_golden():
    """Self-golden (tests/golden/make_golden.py): three training steps of a tiny ensemble with the epoch
    schedule, with and without dropout — freezes today's oracle so later edits cannot drift silently."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'oracle_tiny_train.npz'))
    for tag, rate in (('plain', 0.0), ('dropout', 0.25)):
        rng = np.random.default_rng(7)
        members = []
        for _ in range(2):
            arrays, fan = [], 6
            for _ in range(2):
                arrays += [rng.normal(0, 0.4, (fan, 8)).astype(np.float32), rng.normal(0, 0.1, 8).astype(np.float32)]
                fan = 8
            for _ in range(2):
                arrays += [rng.normal(0, 0.4, (8, 4)).astype(np.float32), rng.normal(0, 0.1, 4).astype(np.float32)]
            members.append(arrays)
        tr = T.EnsembleTrainer(members, batch_size=5, learning_rate=1e-2, learning_rate_schedule=True,
                               training_steps=2, train_epochs=3, dropout_rate=rate, dropout_seed=11)
        losses = []
        for s in range(3):
            x = rng.uniform(0, 1, (2, 5, 6)).astype(np.float32)
            y = rng.normal(0, 0.2, (2, 5, 4)).astype(np.float32)
            losses.append(tr.training_step(x, y))
        np.testing.assert_allclose(np.asarray(losses, np.float32), gold[tag + '_losses'], rtol=1e-6)
        for e, net in enumerate(tr.nets):
            for i, a in enumerate(net.arrays):
                np.testing.assert_allclose(a, gold['%s_member%d_var%d' % (tag, e, i)], rtol=1e-6, atol=1e-7)
