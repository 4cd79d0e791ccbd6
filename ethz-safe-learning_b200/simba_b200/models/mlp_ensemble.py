"""MlpEnsemble — the inference half of simba/models/mlp_ensemble.py behind the same interface.

The E Gaussian MLPs (mlp_ensemble.py:37-61, :111-112) live as packed fp32 + bf16 images inside a
`simba_model` handle of libsimba_b200.so; `forward` / `__call__` (mlp_ensemble.py:122-132,
:189-193) run the fp32 CUDA kernel. Training (`fit`, `training_step`, mlp_ensemble.py:134-187) is
outside the planning path and raises NotImplementedError.
"""
import ctypes as C

import numpy as np
import torch

from .. import _device, _lib


class MemberWeights(object):
    """One GaussianDistMlp's variables in Keras order (mlp_ensemble.py:46-50, :28-30):
    L x (kernel[in, U], bias[U]), mu head (kernel[U, O], bias[O]), var head (kernel[U, O], bias[O])."""

    def __init__(self, owner, index, arrays):
        self._owner = owner
        self._index = index
        self._arrays = arrays

    def get_weights(self):
        return [a.copy() for a in self._arrays]

    def set_weights(self, arrays):
        if len(arrays) != len(self._arrays):
            raise ValueError("expected %d arrays, got %d" % (len(self._arrays), len(arrays)))
        new = []
        for old, a in zip(self._arrays, arrays):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
            if a.shape != old.shape:
                raise ValueError("shape %s does not match %s" % (a.shape, old.shape))
            new.append(a)
        self._arrays = new
        self._owner._dirty = True

    @property
    def trainable_variables(self):
        return self.get_weights()


def glorot_uniform(rng, fan_in, fan_out):
    """Keras Dense default initialiser (mlp_ensemble.py:13, :28-30)."""
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=(fan_in, fan_out)).astype(np.float32)


class MlpEnsemble(object):
    def __init__(self, inputs_dim, outputs_dim, ensemble_size, batch_size=64, validation_split=0.2,
                 learning_rate=0.00025, learning_rate_schedule=True, training_steps=5000,
                 mlp_params=None, train_epochs=1, seed=0):
        mlp_params = dict(mlp_params or {})
        self.inputs_dim = inputs_dim
        self.outputs_dim = outputs_dim
        self.ensemble_size = ensemble_size
        self.batch_size = batch_size
        self.validation_split = validation_split
        self.training_steps = training_steps
        self.mlp_params = mlp_params
        self.n_layers = int(mlp_params.get('n_layers', 4))
        self.units = int(mlp_params.get('units', 128))
        activation = mlp_params.get('activation', 'tf.nn.relu')
        if activation not in ('tf.nn.relu', 'relu'):
            raise _lib.SimbaError(-6, "only ReLU hidden activations are fused (got %r)" % (activation,))
        if float(mlp_params.get('dropout_rate', 0.0)) != 0.0:
            # inference runs with training=False, so dropout is the identity either way
            pass
        rng = np.random.default_rng(seed)
        self.ensemble = []
        for e in range(ensemble_size):
            arrays = []
            fan_in = inputs_dim
            for _ in range(self.n_layers):
                arrays += [glorot_uniform(rng, fan_in, self.units), np.zeros(self.units, np.float32)]
                fan_in = self.units
            for _ in range(2):
                arrays += [glorot_uniform(rng, self.units, outputs_dim),
                           np.zeros(outputs_dim, np.float32)]
            self.ensemble.append(MemberWeights(self, e, arrays))
        self._handle = None
        self._dirty = True
        self._scaler = None          # (min, max, scale_features) set by TransitionModel
        self._lib = _lib.load()

    # -- device image ------------------------------------------------------------------------
    def _ensure_handle(self):
        if self._handle is None:
            _device.require_cuda()
            cfg = _lib.ModelConfig(self.outputs_dim, self.inputs_dim - self.outputs_dim,
                                   self.ensemble_size, self.n_layers, self.units)
            h = C.c_void_p()
            _lib.check(self._lib.simba_model_create(C.byref(cfg), C.byref(h)))
            self._handle = h
            self._dirty = True
        if self._dirty:
            for e, member in enumerate(self.ensemble):
                arrays = member._arrays
                for l in range(self.n_layers + 2):
                    _lib.check(self._lib.simba_model_set_layer(
                        self._handle, e, l, _device.ptr(arrays[2 * l]), _device.ptr(arrays[2 * l + 1])))
            if self._scaler is None:
                _lib.check(self._lib.simba_model_set_scaler(self._handle, None, None, 0))
            else:
                mn, mx, on = self._scaler
                _lib.check(self._lib.simba_model_set_scaler(
                    self._handle, _device.ptr(mn), _device.ptr(mx), int(on)))
            _lib.check(self._lib.simba_model_commit(self._handle))
            self._dirty = False
        return self._handle

    def _set_scaler(self, inputs_min, inputs_max, scale_features):
        self._scaler = (np.ascontiguousarray(inputs_min, dtype=np.float32),
                        np.ascontiguousarray(inputs_max, dtype=np.float32), bool(scale_features))
        self._dirty = True

    def __del__(self):
        try:
            if self._handle is not None:
                self._lib.simba_model_destroy(self._handle)
        except Exception:
            pass

    # -- reference interface --------------------------------------------------------------------
    def build(self):
        self._ensure_handle()

    def _run(self, inputs, eps):
        x, kind = _device.to_device(inputs)
        if x.dim() != 2 or x.shape[1] != self.inputs_dim:
            raise ValueError("inputs must be [B, %d]" % self.inputs_dim)
        b = x.shape[0]
        h = self._ensure_handle()
        mu = torch.empty((b, self.outputs_dim), dtype=torch.float32, device=x.device)
        var = torch.empty_like(mu)
        smp, e = None, None
        if eps is not None:
            e, _ = _device.to_device(eps)
            smp = torch.empty_like(mu)
        _lib.check(self._lib.simba_ensemble_forward(h, _device.ptr(x), _device.ptr(e), b,
                                                    _device.ptr(mu), _device.ptr(var),
                                                    _device.ptr(smp), _device.stream_ptr()))
        return mu, var, smp, kind

    def forward(self, inputs):
        """mlp_ensemble.py:122-132: rows split into E contiguous chunks -> (cat_mus, cat_vars)."""
        mu, var, _, kind = self._run(inputs, None)
        return _device.like_input(mu, kind), _device.like_input(var, kind)

    def __call__(self, inputs, eps=None, *args, **kwargs):
        """mlp_ensemble.py:189-193 -> (mean, stddev, sample). `eps` are the N(0,1) draws of
        Normal.sample(); if omitted they are drawn with torch on the device."""
        x, _ = _device.to_device(inputs)
        if eps is None:
            eps = torch.randn((x.shape[0], self.outputs_dim), dtype=torch.float32, device=x.device)
        mu, var, smp, kind = self._run(inputs, eps)
        return (_device.like_input(mu, kind), _device.like_input(torch.sqrt(var), kind),
                _device.like_input(smp, kind))

    def fit(self, inputs, targets):
        raise NotImplementedError(
            "ensemble training (mlp_ensemble.py:134-187) is outside the accelerated planning "
            "path (SURVEY.md section 8 f1); load trained weights with ensemble[e].set_weights().")
