"""TEST INFRASTRUCTURE ONLY — CPU restatement of the ensemble training step.

PARITY UNPINNED against TensorFlow: the reference's step (`MlpEnsemble.training_step`,
simba/models/mlp_ensemble.py:134-146) is a `tf.GradientTape` over Keras layers followed by
`tf.keras.optimizers.Adam(clipvalue=1.0, epsilon=1e-5)` (mlp_ensemble.py:113-117); TensorFlow is
not installable here, so the forward/backward below restate the published op semantics and are
pinned by calculus instead: tests/test_train_oracle.py checks every analytic gradient against
float64 central differences and the Adam update against a hand-computed step.

Restated pieces (file:line of what each follows):
  * forward: h_l = dropout(relu(h_{l-1} W_l + b_l)) (mlp_ensemble.py:17-22); mu = h_L W_mu + b_mu;
    var = softplus(h_L W_v + b_v) + 1e-4 (mlp_ensemble.py:28-34). Dropout (training only, shipped rate
    0.0, config/models.yaml:13) is tf.nn.dropout: x * keep / (1 - rate); the keep masks are inputs of
    the oracle (TensorFlow's generator cannot be reproduced), produced by oracle/philox.dropout_keep
    when the device trainer's masks are to be matched
  * negative_log_likelihood (mlp_ensemble.py:64-67):
    0.5 * mean(log(2 pi var)) + 0.5 * mean((mu - y)^2 / var), means over batch x outputs
  * loss = sum_e nll_e / E (mlp_ensemble.py:139-141)
  * EpochLearningRateSchedule (mlp_ensemble.py:70-83):
    lr(step) = max(lr0 * (1 - floor(step / steps_per_epoch) / train_epochs), 0)
  * Adam as documented for tf.keras.optimizers.Adam (non-amsgrad), t = iterations + 1:
      g   <- clip(g, -clipvalue, +clipvalue)            (element-wise, `clipvalue`)
      lr_t = lr(iterations) * sqrt(1 - beta2^t) / (1 - beta1^t)
      m   <- m + (1 - beta1) (g - m);  v <- v + (1 - beta2) (g^2 - v)
      w   <- w - lr_t * m / (sqrt(v) + epsilon)
  * fit's batching (mlp_ensemble.py:167-186): per pass, one permutation of the training rows per
    member, `np.array_split` into ceil(n / batch_size) batches (sizes may differ by one).
"""
import numpy as np

from .simba_oracle import tf_softplus


def negative_log_likelihood(y, mu, var):
    """mlp_ensemble.py:64-67."""
    dt = mu.dtype.type
    return dt(0.5) * np.mean(np.log(dt(2.0 * np.pi) * var)) + \
        dt(0.5) * np.mean(np.square(mu - y) / var)


def lr_schedule(step, lr0, steps_per_epoch, train_epochs, enabled=True):
    """mlp_ensemble.py:80-83 (step = optimizer.iterations before the update)."""
    if not enabled:
        return np.float32(lr0)
    epochs = np.float32(np.floor(int(step) / int(steps_per_epoch)))
    return np.float32(max(np.float32(lr0) * (np.float32(1.0) - epochs / np.float32(train_epochs)),
                          np.float32(0.0)))


class MemberNet:
    """One GaussianDistMlp's variables in Keras order: L x (W[in,U], b[U]), (W_mu, b_mu), (W_v, b_v)."""

    def __init__(self, arrays, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.arrays = [np.array(a, dtype=self.dtype) for a in arrays]
        self.n_layers = len(arrays) // 2 - 2

    def forward(self, x, keep=None):
        """Returns (mu, var, cache) — cache holds what backward needs. keep: per hidden layer the
        dropout multiplier [B, U] (0 or 1 / (1 - rate)), or None (no dropout)."""
        L = self.n_layers
        hs = [np.asarray(x, dtype=self.dtype)]
        for l in range(L):
            h = np.maximum(hs[-1] @ self.arrays[2 * l] + self.arrays[2 * l + 1], 0)
            if keep is not None:
                h = h * keep[l].astype(self.dtype)
            hs.append(h)
        mu = hs[-1] @ self.arrays[2 * L] + self.arrays[2 * L + 1]
        raw = hs[-1] @ self.arrays[2 * L + 2] + self.arrays[2 * L + 3]
        var = tf_softplus(raw) + self.dtype.type(1e-4)
        return mu, var, (hs, raw)

    def loss_and_grads(self, x, y, loss_scale=1.0, keep=None):
        """d(loss_scale * nll)/d(variables), in Keras variable order."""
        dt = self.dtype.type
        L = self.n_layers
        mu, var, (hs, raw) = self.forward(x, keep)
        y = np.asarray(y, dtype=self.dtype)
        loss = negative_log_likelihood(y, mu, var) * dt(loss_scale)
        c = dt(loss_scale) / dt(mu.size)
        diff = mu - y
        d_mu = c * diff / var
        d_var = dt(0.5) * c * (dt(1.0) / var - np.square(diff) / np.square(var))
        sig = dt(1.0) / (dt(1.0) + np.exp(-raw))                     # d softplus / d raw
        d_raw = d_var * sig
        grads = [None] * len(self.arrays)
        h = hs[L]
        grads[2 * L] = h.T @ d_mu
        grads[2 * L + 1] = d_mu.sum(axis=0)
        grads[2 * L + 2] = h.T @ d_raw
        grads[2 * L + 3] = d_raw.sum(axis=0)
        d_h = d_mu @ self.arrays[2 * L].T + d_raw @ self.arrays[2 * L + 2].T
        for l in range(L - 1, -1, -1):
            # relu'(z) and the dropout multiplier: hs[l + 1] > 0 <=> active and kept
            d_z = d_h * (hs[l + 1] > 0) * (keep[l].astype(self.dtype) if keep is not None else dt(1.0))
            grads[2 * l] = hs[l].T @ d_z
            grads[2 * l + 1] = d_z.sum(axis=0)
            if l > 0:
                d_h = d_z @ self.arrays[2 * l].T
        return loss, grads


class Adam:
    def __init__(self, shapes, lr0, steps_per_epoch, train_epochs, schedule=True, beta1=0.9,
                 beta2=0.999, epsilon=1e-5, clipvalue=1.0, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.m = [np.zeros(s, self.dtype) for s in shapes]
        self.v = [np.zeros(s, self.dtype) for s in shapes]
        self.iterations = 0
        self.lr0, self.steps_per_epoch, self.train_epochs = lr0, steps_per_epoch, train_epochs
        self.schedule = schedule
        self.beta1, self.beta2, self.epsilon, self.clipvalue = beta1, beta2, epsilon, clipvalue

    def lr_t(self):
        dt = self.dtype.type
        t = self.iterations + 1
        lr = dt(lr_schedule(self.iterations, self.lr0, self.steps_per_epoch, self.train_epochs,
                            self.schedule))
        return lr * np.sqrt(dt(1.0) - dt(self.beta2) ** dt(t)) / (dt(1.0) - dt(self.beta1) ** dt(t))

    def apply(self, variables, grads):
        dt = self.dtype.type
        lr_t = self.lr_t()
        for w, g, m, v in zip(variables, grads, self.m, self.v):
            g = np.clip(g, -dt(self.clipvalue), dt(self.clipvalue))
            m += (dt(1.0) - dt(self.beta1)) * (g - m)
            v += (dt(1.0) - dt(self.beta2)) * (np.square(g) - v)
            w -= lr_t * m / (np.sqrt(v) + dt(self.epsilon))
        self.iterations += 1


class EnsembleTrainer:
    """training_step / validation_step / fit of mlp_ensemble.py:134-187 on numpy arrays."""

    def __init__(self, members, batch_size=64, learning_rate=0.00025, learning_rate_schedule=True,
                 training_steps=5000, train_epochs=1, dtype=np.float32, dropout_rate=0.0, dropout_seed=0):
        self.dropout_rate, self.dropout_seed = float(dropout_rate), int(dropout_seed)
        self.nets = [MemberNet(m, dtype) for m in members]
        self.batch_size = batch_size
        self.training_steps = training_steps
        flat = [a for n in self.nets for a in n.arrays]
        self.optimizer = Adam([a.shape for a in flat], learning_rate, training_steps, train_epochs,
                              learning_rate_schedule, dtype=dtype)
        self.last_grads = None

    @property
    def ensemble_size(self):
        return len(self.nets)

    def training_step(self, inputs, targets):
        """inputs [E, B, in], targets [E, B, O] -> scalar loss; updates the variables in place."""
        E = self.ensemble_size
        loss = self.nets[0].dtype.type(0.0)
        grads = []
        for e, net in enumerate(self.nets):
            keep = None
            if self.dropout_rate > 0.0:
                from . import philox
                scale = np.float32(1.0) / (np.float32(1.0) - np.float32(self.dropout_rate))
                keep = [philox.dropout_keep(self.dropout_seed, self.optimizer.iterations, e, l,
                                            inputs[e].shape[0], net.arrays[2 * l].shape[1],
                                            self.dropout_rate).astype(np.float32) * scale
                        for l in range(net.n_layers)]
            l, g = net.loss_and_grads(inputs[e], targets[e], 1.0 / E, keep)
            loss = loss + l
            grads += g
        self.last_grads = grads
        self.optimizer.apply([a for n in self.nets for a in n.arrays], grads)
        return loss

    def validation_step(self, inputs, targets):
        E = self.ensemble_size
        loss = self.nets[0].dtype.type(0.0)
        for net in self.nets:
            mu, var, _ = net.forward(inputs)
            loss = loss + negative_log_likelihood(np.asarray(targets, net.dtype), mu, var) / \
                net.dtype.type(E)
        return loss

    def fit_batches(self, inputs, targets, batch_index):
        """batch_index: list of [E, B_s] row-index arrays, one per step (see `make_batch_index`)."""
        losses = np.empty(len(batch_index))
        for s, idx in enumerate(batch_index):
            losses[s] = self.training_step(inputs[idx], targets[idx])
        return losses


def make_batch_index(rng, n_rows, ensemble_size, batch_size, steps):
    """The batch schedule of `fit` (mlp_ensemble.py:167-186) as explicit index arrays."""
    n_batches = int(np.ceil(n_rows / batch_size))
    out = []
    while len(out) < steps:
        shuffles = np.array([rng.permutation(n_rows) for _ in range(ensemble_size)])
        for b in np.array_split(shuffles, n_batches, axis=1):
            out.append(b)
            if len(out) == steps:
                break
    return out
