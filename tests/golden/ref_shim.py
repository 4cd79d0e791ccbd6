"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference sources (/root/reference/simba/...) in this
container, where TensorFlow, TensorFlow-Probability, gym and tensorboardX cannot be installed.

`install()` puts minimal stand-ins for those four packages into sys.modules — just the ~60 stock ops
the planning path calls, each restated on torch-CPU fp32 following the TensorFlow documentation — and
then the reference's own Python (CemMpc / SafeCemMpc / MpcPolicy / TransitionModel / MlpEnsemble /
GaussianDistMlp / SafetyGymStateScorer: loops, done-mask orders, reshapes, tiling, Beta test, top-k /
best-so-far / refit logic) executes as written. What is pinned by fixtures generated this way is the
reference's CODE; what stays a restatement is the arithmetic of each stock TF op (listed below).
tests/golden/make_reference_golden.py is the generator; the fixtures travel, /root/reference does not.

Semantics restated here (TF docs / TF source where the docs are silent):
  * tensors are immutable: `x += y` rebinds (a torch in-place add would also change the copy the
    reference already wrote into its TensorArray, transition_model.py:69-76);
  * tf.random.normal(shape, mean, stddev) = z * stddev + mean and tfp Normal(loc, scale).sample() =
    z * scale + loc, with z taken from the queue the caller fills (parity is "given identical draws");
  * tf.nn.top_k(sorted=False): the k largest, ties -> lower index (the order of the result is
    unspecified in TF; returned here in descending score order, so argmax picks the lowest index
    among equal best scores); tf.argmax: first maximum;
  * tf.nn.moments: mean, then mean of squared differences (population variance);
  * tf.clip_by_value = maximum(minimum(t, hi), lo); tf.math.softplus: TF's thresholded kernel
    (core/kernels/softplus_op.h); tf.split(x, n): n equal chunks, error unless divisible;
  * Dense: x @ kernel[in, out] + bias, glorot_uniform / zeros; Dropout: identity unless training.
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get('SIMBA_REFERENCE_ROOT', '/root/reference')


class T(torch.Tensor):
    """torch tensor with TensorFlow's value semantics for augmented assignment."""

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o


_DTYPES = {}


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        out = x
    else:
        a = np.asarray(x)
        if a.dtype == np.float64:
            a = a.astype(np.float32)          # TF converts python floats / float64 defaults to float32 here
        out = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        out = out.to(_DTYPES.get(dtype, dtype))
    return out.as_subclass(T)


class DrawQueue:
    """Standard-normal draws the stand-in RNG ops consume, in call order."""

    def __init__(self):
        self.items = []

    def push(self, z):
        self.items.append(np.asarray(z, dtype=np.float32))

    def pop(self, shape):
        z = self.items.pop(0)
        assert tuple(z.shape) == tuple(int(s) for s in shape), (z.shape, tuple(shape))
        return _t(z)


DRAWS = DrawQueue()


class _TensorArray:
    def __init__(self, dtype, size):
        self.items = [None] * int(size)

    def write(self, i, v):
        self.items[int(i)] = v
        return self

    def stack(self):
        return torch.stack(self.items).as_subclass(T)


def _function(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def _tf_softplus(x):
    thr = float(np.log(np.finfo(np.float32).eps) + 2.0)
    ex = torch.exp(x)
    return torch.where(x > -thr, x, torch.where(x < thr, ex, torch.log1p(ex)))


def _top_k(scores, k, sorted=True):
    order = torch.argsort(-scores, stable=True)[:int(k)]
    return scores[order], order.to(torch.int32)


def _moments(x, axes):
    mean = x.mean(dim=axes)
    return mean, ((x - mean) ** 2).mean(dim=axes)


class _Layer:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class _Dense(_Layer):
    def __init__(self, units, activation=None):
        self.units, self.activation = units, activation
        self.kernel = self.bias = None

    def build(self, in_dim):
        lim = float(np.sqrt(6.0 / (in_dim + self.units)))
        self.kernel = _t(np.random.uniform(-lim, lim, (in_dim, self.units)).astype(np.float32))
        self.bias = _t(np.zeros((self.units,), np.float32))

    def call(self, x):
        if self.kernel is None:
            self.build(x.shape[-1])
        y = torch.matmul(x, self.kernel) + self.bias
        return self.activation(y) if self.activation is not None else y

    @property
    def trainable_variables(self):
        return [self.kernel, self.bias]


class _Dropout(_Layer):
    def __init__(self, rate):
        self.rate = rate

    def call(self, x, training=None):
        assert not (bool(training) and self.rate > 0.0), "dropout draws are outside this stand-in"
        return x


class _InputLayer(_Layer):
    def __init__(self, input_shape=None):
        self.input_shape = input_shape

    def call(self, x, *a, **k):
        return x


class _Sequential(_Layer):
    def __init__(self, layers):
        self.layers = list(layers)

    def call(self, x, training=None):
        for l in self.layers:
            x = l(x, training) if not isinstance(l, (_InputLayer, _Dense)) else l(x)
        return x


def install(reference_root=REFERENCE_ROOT):
    """Install the stand-ins and make `import simba...` resolve to the unmodified reference tree."""
    if 'tensorflow' in sys.modules and getattr(sys.modules['tensorflow'], '_simba_stand_in', False):
        return sys.modules['tensorflow']
    tf = types.ModuleType('tensorflow')
    tf._simba_stand_in = True
    tf.float32, tf.bool, tf.int32 = 'float32', 'bool', 'int32'
    _DTYPES.update({'float32': torch.float32, 'bool': torch.bool, 'int32': torch.int32, bool: torch.bool,
                    float: torch.float32, int: torch.int32})
    tf.constant = lambda v, dtype=None: _t(v, dtype)
    tf.convert_to_tensor = lambda v, dtype=None: _t(v, dtype)
    tf.broadcast_to = lambda x, shape: torch.broadcast_to(_t(x), tuple(int(s) for s in shape)).as_subclass(T)
    tf.zeros = lambda shape, dtype='float32': _t(torch.zeros(tuple(int(s) for s in shape), dtype=_DTYPES[dtype]))
    tf.ones = lambda shape, dtype='float32': _t(torch.ones(tuple(int(s) for s in shape), dtype=_DTYPES[dtype]))
    tf.zeros_like = lambda x, dtype=None: _t(torch.zeros_like(x, dtype=_DTYPES.get(dtype)))
    tf.range = lambda n: range(int(n))
    tf.shape = lambda x: tuple(int(s) for s in x.shape)
    tf.cast = lambda x, dtype: _t(x, dtype)
    tf.reshape = lambda x, shape: x.reshape(tuple(int(s) for s in shape))
    tf.tile = lambda x, multiples: x.repeat(*[int(m) for m in multiples])
    tf.concat = lambda xs, axis: torch.cat([_t(x, 'float32') for x in xs], dim=axis).as_subclass(T)
    tf.split = lambda x, n, axis=0: _split(x, n, axis)
    tf.transpose = lambda x, perm: x.permute(*perm)
    tf.gather = lambda x, idx, axis=0: x.index_select(axis, idx.to(torch.int64))
    tf.where = lambda c, a, b: torch.where(c, _t(a, b.dtype if isinstance(b, torch.Tensor) else None), _t(b))
    tf.linspace = lambda a, b, n: _t(torch.linspace(float(a), float(b), int(n), dtype=torch.float32))
    tf.squeeze = lambda x: x.squeeze()
    tf.sqrt, tf.square, tf.floor = torch.sqrt, torch.square, torch.floor
    tf.maximum = lambda a, b: torch.maximum(_t(a), _t(b, 'float32'))
    tf.clip_by_value = lambda t, lo, hi: torch.maximum(torch.minimum(t, _t(hi, t.dtype)), _t(lo, t.dtype))
    tf.greater = lambda a, b: a > b
    tf.less = lambda a, b: a < b
    tf.less_equal = lambda a, b: a <= b
    tf.logical_or, tf.logical_and, tf.logical_not = torch.logical_or, torch.logical_and, torch.logical_not
    tf.argmax = lambda x, axis=None: torch.argmax(x) if axis is None else torch.argmax(x, dim=axis)
    tf.reduce_mean = lambda x, axis=None: x.mean() if axis is None else x.mean(dim=axis)
    tf.reduce_sum = lambda x, axis=None: x.sum() if axis is None else x.sum(dim=axis)
    tf.reduce_min = lambda x, axis=None: x.min() if axis is None else x.min(dim=axis).values
    tf.TensorArray = _TensorArray
    tf.TensorSpec = lambda shape=None, dtype=None: None
    tf.function = _function
    tf.Module = type('Module', (), {'__init__': lambda self, *a, **k: None})

    def _split(x, n, axis):
        if x.shape[axis] % n:
            raise ValueError("tf.split: dimension %d not divisible by %d" % (x.shape[axis], n))
        return list(torch.split(x, x.shape[axis] // n, dim=axis))

    tf.math = types.ModuleType('tensorflow.math')
    tf.math.softplus, tf.math.log = _tf_softplus, torch.log
    tf.math.divide = lambda a, b: a / b
    tf.math.reduce_any = lambda x, axis=None: x.any() if axis is None else x.any(dim=axis)
    tf.nn = types.ModuleType('tensorflow.nn')
    tf.nn.relu, tf.nn.top_k, tf.nn.moments = torch.relu, _top_k, _moments
    tf.random = types.ModuleType('tensorflow.random')
    tf.random.normal = lambda shape, mean=0.0, stddev=1.0: DRAWS.pop(shape) * _t(stddev, 'float32') + _t(mean, 'float32')
    tf.linalg = types.ModuleType('tensorflow.linalg')
    keras = types.ModuleType('tensorflow.keras')
    keras.layers = types.ModuleType('tensorflow.keras.layers')
    keras.layers.Layer, keras.layers.Dense, keras.layers.Dropout = _Layer, _Dense, _Dropout
    keras.layers.InputLayer = _InputLayer
    keras.Sequential, keras.Model = _Sequential, _Layer
    keras.optimizers = types.ModuleType('tensorflow.keras.optimizers')
    keras.optimizers.Adam = lambda *a, **k: None       # construction only: training is outside this stand-in
    keras.optimizers.schedules = types.ModuleType('tensorflow.keras.optimizers.schedules')
    keras.optimizers.schedules.LearningRateSchedule = type('LearningRateSchedule', (), {'__init__': lambda self: None})
    tf.keras = keras
    compat = types.ModuleType('tensorflow.compat')
    compat.v1 = types.ModuleType('tensorflow.compat.v1')
    tf.compat = compat

    tfp = types.ModuleType('tensorflow_probability')
    tfp.distributions = types.ModuleType('tensorflow_probability.distributions')

    class Normal:
        def __init__(self, loc, scale):
            self.loc, self.scale = loc, scale

        def mean(self):
            return self.loc

        def stddev(self):
            return self.scale

        def sample(self):
            return DRAWS.pop(self.loc.shape) * self.scale + self.loc

    tfp.distributions.Normal = Normal

    gym = types.ModuleType('gym')
    gym.spaces = types.ModuleType('gym.spaces')

    class Box:
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.shape = self.low.shape

        def is_bounded(self):
            return bool(np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)))

    gym.spaces.Box = Box
    gym.Wrapper = type('Wrapper', (), {})
    tbx = types.ModuleType('tensorboardX')
    tbx.SummaryWriter = type('SummaryWriter', (), {})

    for name, mod in (('tensorflow', tf), ('tensorflow.math', tf.math), ('tensorflow.nn', tf.nn),
                      ('tensorflow.random', tf.random), ('tensorflow.keras', keras),
                      ('tensorflow.compat', compat), ('tensorflow.compat.v1', compat.v1),
                      ('tensorflow_probability', tfp), ('gym', gym), ('gym.spaces', gym.spaces),
                      ('tensorboardX', tbx)):
        sys.modules[name] = mod
    # simba.environment_utils/__init__.py registers MuJoCo environments (needs safety_gym): give the
    # package an empty body so that its submodule safety_gym.py (the scorer) imports unmodified
    pkg = types.ModuleType('simba.environment_utils')
    pkg.__path__ = [os.path.join(reference_root, 'simba', 'environment_utils')]
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import simba  # noqa: F401  (the reference package itself)
    sys.modules['simba.environment_utils'] = pkg
    return tf


def set_member_weights(mlp, arrays):
    """Load one member's variables, Keras order (mlp_ensemble.py:46-50, :28-30): L x (kernel, bias), mu head,
    var head."""
    denses = [l._dense for l in mlp.forward.layers if hasattr(l, '_dense')] + [mlp.head._mu, mlp.head._var]
    assert len(arrays) == 2 * len(denses)
    for i, d in enumerate(denses):
        d.kernel = _t(np.asarray(arrays[2 * i], np.float32))
        d.bias = _t(np.asarray(arrays[2 * i + 1], np.float32))
