"""Minimal eager stand-in for the TensorFlow-1.x API surface that qmcnn uses.

TEST INFRASTRUCTURE ONLY (lives under ``oracle/``; see ``oracle/__init__.py``).

Why it exists: the reference is TF-1 graph-mode Python and TensorFlow is not
installable here, so the reference cannot run as shipped.  Putting this directory
on ``sys.path`` lets ``/root/reference/{helpers,models,sampler,mcmc_tf}.py`` be
imported and executed UNMODIFIED: every ``tf.*`` call they make lands here and is
carried out immediately on torch-CPU tensors (float32/complex64 like TF, or
float64/complex128 with ``set_precision('double')`` for a ground-truth run).
``tests/golden/make_golden.py`` uses that to record golden vectors of the
reference's own Python - its indexing, window order, bookkeeping, control flow,
operator order - which pin ``oracle/`` and the CUDA path.

What it does NOT pin: TensorFlow's own numerical kernels (conv2d, exp, log, tanh
rounding; its RNG streams).  Each op below restates the *documented* TF-1
semantics (NHWC x HWIO 'VALID' cross-correlation, floor division/modulo,
``gather_nd``/``scatter_update`` index conventions, TF-1 Adam with epsilon outside
the square root, ...).

Execution model: one Python call of a graph-building function == one
``Session.run`` of the tensor it returns.  ``tf.control_dependencies`` is a no-op
because program order already is the dependency order the reference asks for
(checked op by op against ``sampler.py:72-177``); tensors are values (every op
returns fresh storage and variable updates replace storage, never mutate it), so
a tensor read before an assign keeps the old value exactly as in a TF graph.
Optimiser slots live in a module-level table keyed by variable, standing in for
the one persistent ``AdamOptimizer`` of the reference's single graph.
"""
import contextlib
import math

import numpy as np
import torch

__version__ = "1.shim"

# --------------------------------------------------------------------------
# dtypes and precision mode
# --------------------------------------------------------------------------
_PRECISION = "single"


def set_precision(mode):
    """'single': tf.float32/complex64 are what they say (TF's arithmetic types).
    'double': they are carried in float64/complex128 (ground truth)."""
    global _PRECISION
    assert mode in ("single", "double")
    _PRECISION = mode


class DType(object):
    def __init__(self, name, single, double):
        self.name, self._s, self._d = name, single, double

    @property
    def torch(self):
        return self._s if _PRECISION == "single" else self._d

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", torch.float32, torch.float64)
float64 = DType("float64", torch.float64, torch.float64)
complex64 = DType("complex64", torch.complex64, torch.complex128)
int32 = DType("int32", torch.int32, torch.int32)
int64 = DType("int64", torch.int64, torch.int64)
bool = DType("bool", torch.bool, torch.bool)   # noqa: A001 (TF exports tf.bool)

import builtins as _b  # noqa: E402  (tf.bool shadows the builtin below this line)


def _torch_dtype(dtype):
    return dtype.torch if isinstance(dtype, DType) else dtype


# --------------------------------------------------------------------------
# deterministic, logged randomness (tf.random_uniform / initializers)
# --------------------------------------------------------------------------
_rng = np.random.Generator(np.random.Philox(0))
random_log = []     # every draw, in program order: (kind, ndarray)


def set_random_seed(seed):
    global _rng
    _rng = np.random.Generator(np.random.Philox(seed))
    del random_log[:]


# --------------------------------------------------------------------------
# Tensor / Variable
# --------------------------------------------------------------------------
def _as_torch(x, like=None):
    """Python / numpy / Tensor -> torch tensor.  A non-tensor operand takes the
    dtype of the tensor it meets (``ops.convert_to_tensor(y, dtype=x.dtype)``)."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if like is not None and t.dtype != like.dtype:
            t = t.to(like.dtype)
        elif like is None and t.dtype == torch.float32:
            t = t.to(float32.torch)
        return t
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, (Tensor, torch.Tensor)) for e in x):
            return torch.stack([_as_torch(e, like) for e in x])
        return _as_torch(np.asarray(x), like)
    # scalars
    if like is not None:
        return torch.tensor(x, dtype=like.dtype)
    if isinstance(x, (_b.bool, np.bool_)):
        return torch.tensor(_b.bool(x))
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x), dtype=torch.int32)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=float32.torch)
    if isinstance(x, (_b.complex, np.complexfloating)):
        return torch.tensor(_b.complex(x), dtype=complex64.torch)
    raise TypeError("cannot convert %r" % (x,))


def _pair(a, b):
    """Binary-op operand conversion with TF's 'constant adopts tensor dtype'."""
    if isinstance(a, Tensor):
        return a.t, _as_torch(b, like=a.t)
    return _as_torch(a, like=b.t), b.t


def _index(idx):
    """Index objects may contain Tensors (e.g. ``var[i]`` with a loop counter)."""
    if isinstance(idx, tuple):
        return tuple(_index(i) for i in idx)
    if isinstance(idx, Tensor):
        return int(idx.t) if idx.t.ndim == 0 else idx.t.long()
    if isinstance(idx, np.integer):
        return int(idx)
    return idx


class Tensor(object):
    __array_priority__ = 1000.0
    __array_ufunc__ = None           # numpy operands defer to our reflected operators
    __hash__ = object.__hash__

    def __init__(self, t):
        self.t = t

    # -- introspection
    @property
    def dtype(self):
        return self.t.dtype

    @property
    def shape(self):
        return tuple(self.t.shape)

    def get_shape(self):
        return self.shape

    def numpy(self):
        return self.t.detach().resolve_conj().numpy().copy()

    def __repr__(self):
        return "tf1shim.Tensor(%r)" % (self.t,)

    def __bool__(self):
        return _b.bool(self.t)

    __nonzero__ = __bool__

    def __index__(self):
        return int(self.t)

    def __getitem__(self, idx):
        return Tensor(self.t[_index(idx)])

    # -- arithmetic
    def __add__(self, o):
        a, b = _pair(self, o); return Tensor(a + b)          # noqa: E702

    def __radd__(self, o):
        a, b = _pair(o, self); return Tensor(a + b)          # noqa: E702

    def __sub__(self, o):
        a, b = _pair(self, o); return Tensor(a - b)          # noqa: E702

    def __rsub__(self, o):
        a, b = _pair(o, self); return Tensor(a - b)          # noqa: E702

    def __mul__(self, o):
        a, b = _pair(self, o); return Tensor(a * b)          # noqa: E702

    def __rmul__(self, o):
        a, b = _pair(o, self); return Tensor(a * b)          # noqa: E702

    def __truediv__(self, o):
        a, b = _pair(self, o); return Tensor(a / b)          # noqa: E702

    def __rtruediv__(self, o):
        a, b = _pair(o, self); return Tensor(a / b)          # noqa: E702

    __div__, __rdiv__ = __truediv__, __rtruediv__

    def __floordiv__(self, o):
        a, b = _pair(self, o); return Tensor(torch.div(a, b, rounding_mode="floor"))   # noqa: E702

    def __mod__(self, o):
        a, b = _pair(self, o); return Tensor(torch.remainder(a, b))   # noqa: E702  (floormod)

    def __neg__(self):
        return Tensor(-self.t)

    def __pow__(self, o):
        a, b = _pair(self, o); return Tensor(torch.pow(a, b))   # noqa: E702

    # -- comparisons
    def __lt__(self, o):
        a, b = _pair(self, o); return Tensor(a < b)          # noqa: E702

    def __le__(self, o):
        a, b = _pair(self, o); return Tensor(a <= b)         # noqa: E702

    def __gt__(self, o):
        a, b = _pair(self, o); return Tensor(a > b)          # noqa: E702

    def __ge__(self, o):
        a, b = _pair(self, o); return Tensor(a >= b)         # noqa: E702


class Variable(Tensor):
    """``tf.Variable`` / ``tf.get_variable`` result.  ``self.t`` is the current value."""

    def __init__(self, initial_value, name=None, trainable=True, dtype=None):
        t = _as_torch(initial_value)
        if dtype is not None:
            t = t.to(_torch_dtype(dtype))
        self.trainable = _b.bool(trainable) and t.dtype.is_floating_point
        Tensor.__init__(self, t.clone().requires_grad_(self.trainable))
        self.name = name or "Variable_%d" % len(_all_variables)
        _all_variables.append(self)

    def _set(self, value):
        t = _as_torch(value, like=self.t).detach().to(self.t.dtype)
        assert tuple(t.shape) == tuple(self.t.shape), \
            "assign shape %s to variable %s of shape %s" % (tuple(t.shape), self.name, tuple(self.t.shape))
        self.t = t.clone().requires_grad_(self.trainable)

    def load(self, value, session=None):
        self._set(value)


_all_variables = []
_var_store = {}
_scope_stack = []     # [(name, reuse)]


def reset_default_graph():
    """Forget every variable, scope and optimiser slot."""
    del _all_variables[:]
    _var_store.clear()
    del _scope_stack[:]
    train._slots.clear()


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope_stack.append((name, reuse))
    try:
        yield
    finally:
        _scope_stack.pop()


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    full = "/".join([s for s, _ in _scope_stack] + [name])
    reuse = any(r for _, r in _scope_stack)
    if reuse:
        if full not in _var_store:
            raise ValueError("Variable %s does not exist" % full)
        return _var_store[full]
    if full in _var_store:
        raise ValueError("Variable %s already exists" % full)
    dtype = float32 if dtype is None else dtype
    shape = [int(s) for s in shape]
    value = initializer(shape, dtype)
    v = Variable(value, name=full, trainable=trainable, dtype=dtype)
    _var_store[full] = v
    return v


def random_normal_initializer(mean=0.0, stddev=1.0):
    def init(shape, dtype):
        x = (mean + stddev * _rng.standard_normal(shape)).astype(np.float32)
        random_log.append(("normal", x))
        return torch.from_numpy(x).to(_torch_dtype(dtype))
    return init


def constant_initializer(value=0):
    def init(shape, dtype):
        return torch.full(shape, value, dtype=_torch_dtype(dtype))
    return init


def global_variables_initializer():
    return None      # variables are initialised on creation


def trainable_variables():
    return [v for v in _all_variables if v.trainable]


# --------------------------------------------------------------------------
# graph plumbing that is a no-op under eager execution
# --------------------------------------------------------------------------
@contextlib.contextmanager
def _null(*args, **kwargs):
    yield


device = _null
name_scope = _null
control_dependencies = _null


def group(*ops):
    return None


def Print(x, data, *args, **kwargs):       # noqa: N802
    return x


class ConfigProto(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class Graph(object):
    @contextlib.contextmanager
    def as_default(self):
        yield self


class Session(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "tf1 shim is eager: call the graph-building function, there is no Session")


# --------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------
def constant(value, dtype=None, shape=None):
    t = _as_torch(value)
    if dtype is not None:
        t = t.to(_torch_dtype(dtype))
    return Tensor(t.clone())


def cast(x, dtype):
    return Tensor(_as_torch(x).to(_torch_dtype(dtype)))


def shape(x):
    """Static shapes only: a tuple of Python ints (slices and arithmetic on it
    behave like the reference expects of the shape tensor)."""
    return tuple(int(s) for s in _as_torch(x).shape)


def reshape(x, new_shape):
    new_shape = [int(s) for s in new_shape]
    return Tensor(_as_torch(x).reshape(new_shape).clone())


def transpose(x, perm=None):
    t = _as_torch(x)
    perm = list(range(t.ndim))[::-1] if perm is None else [int(p) for p in perm]
    return Tensor(t.permute(perm).contiguous())


def expand_dims(x, axis):
    return Tensor(_as_torch(x).unsqueeze(int(axis)))


def stack(values, axis=0):
    ts = [_as_torch(v) for v in values]
    return Tensor(torch.stack(ts, int(axis)))


def tile(x, multiples):
    return Tensor(_as_torch(x).repeat([int(m) for m in multiples]))


def slice(x, begin, size):       # noqa: A001 (TF exports tf.slice)
    t = _as_torch(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        b, s = int(b), int(s)
        idx.append(_b.slice(b, t.shape[d] if s == -1 else b + s))
    return Tensor(t[tuple(idx)].clone())


def ones(shape_, dtype=float32):
    shape_ = [int(shape_)] if np.ndim(shape_) == 0 else [int(s) for s in shape_]
    return Tensor(torch.ones(shape_, dtype=_torch_dtype(dtype)))


def zeros_like(x, dtype=None):
    t = _as_torch(x)
    return Tensor(torch.zeros_like(t, dtype=_torch_dtype(dtype) if dtype is not None else t.dtype))


def range(*args, **kwargs):      # noqa: A001 (TF exports tf.range)
    dtype = _torch_dtype(kwargs.pop("dtype", int32))
    return Tensor(torch.arange(*[int(a) for a in args], dtype=dtype))


def _axes(axis):
    if axis is None:
        return None
    if np.ndim(axis) == 0:
        return int(axis)
    return tuple(int(a) for a in axis)


def reduce_sum(x, axis=None):
    t = _as_torch(x)
    ax = _axes(axis)
    return Tensor(t.sum() if ax is None else t.sum(dim=ax))


def reduce_mean(x, axis=None):
    t = _as_torch(x)
    ax = _axes(axis)
    return Tensor(t.mean() if ax is None else t.mean(dim=ax))


def reduce_prod(x, axis=None):
    t = _as_torch(x)
    ax = _axes(axis)
    if ax is None:
        return Tensor(t.prod())
    assert isinstance(ax, int)
    return Tensor(t.prod(dim=ax).to(t.dtype))


def exp(x):
    return Tensor(torch.exp(_as_torch(x)))


def log(x):
    return Tensor(torch.log(_as_torch(x)))


def tanh(x):
    return Tensor(torch.tanh(_as_torch(x)))


def abs(x):      # noqa: A001
    return Tensor(torch.abs(_as_torch(x)))


def pow(x, y):   # noqa: A001
    a, b = _pair(x if isinstance(x, Tensor) else Tensor(_as_torch(x)), y)
    return Tensor(torch.pow(a, b))


def complex(real, imag):     # noqa: A001
    return Tensor(torch.complex(_as_torch(real), _as_torch(imag)))


def real(x):
    return Tensor(torch.real(_as_torch(x)).clone())


def conj(x):
    return Tensor(torch.conj(_as_torch(x)).resolve_conj())


def stop_gradient(x):
    return Tensor(_as_torch(x).detach())


def logical_and(a, b):
    return Tensor(torch.logical_and(_as_torch(a), _as_torch(b)))


def greater_equal(a, b):
    return (a if isinstance(a, Tensor) else Tensor(_as_torch(a))) >= b


def equal(a, b):
    x, y = _pair(a if isinstance(a, Tensor) else Tensor(_as_torch(a)), b)
    return Tensor(x == y)


def gather(params, indices):
    """Axis-0 gather: out[i...] = params[indices[i...]]."""
    p = _as_torch(params)
    return Tensor(p[_as_torch(indices).long()])


def gather_nd(params, indices):
    """indices[..., :K] address the first K axes of params."""
    p = _as_torch(params)
    idx = _as_torch(indices).long()
    k = idx.shape[-1]
    return Tensor(p[tuple(idx[..., j] for j in _b.range(k))])


def boolean_mask(x, mask):
    return Tensor(_as_torch(x)[_as_torch(mask).to(torch.bool)])


def scatter_update(ref, indices, updates):
    """ref[indices[i]] = updates[i] (axis 0); returns ref."""
    assert isinstance(ref, Variable)
    new = ref.t.detach().clone()
    idx = _as_torch(indices)
    upd = _as_torch(updates, like=new).detach().to(new.dtype)
    if idx.ndim == 0:
        new[int(idx)] = upd
    else:
        new[idx.long()] = upd
    ref._set(new)
    return ref


def scatter_nd_update(ref, indices, updates):
    assert isinstance(ref, Variable)
    new = ref.t.detach().clone()
    idx = _as_torch(indices).long()
    k = idx.shape[-1]
    new[tuple(idx[..., j] for j in _b.range(k))] = _as_torch(updates, like=new).detach().to(new.dtype)
    ref._set(new)
    return ref


def assign(ref, value):
    assert isinstance(ref, Variable)
    ref._set(value)
    return ref


def cond(pred, true_fn, false_fn):
    p = pred.t if isinstance(pred, Tensor) else pred
    return true_fn() if _b.bool(p) else false_fn()


def while_loop(cond, body, loop_vars, parallel_iterations=10, back_prop=True):   # noqa: A002
    vars_ = list(loop_vars)
    ctx = torch.no_grad() if not back_prop else _null()
    with ctx:
        while _b.bool(cond(*vars_)):
            out = body(*vars_)
            vars_ = list(out) if isinstance(out, (list, tuple)) else [out]
    return vars_[0] if len(vars_) == 1 else vars_


def map_fn(fn, elems, dtype=None, parallel_iterations=10, back_prop=True):
    t = _as_torch(elems)
    ctx = torch.no_grad() if not back_prop else _null()
    with ctx:
        outs = [_as_torch(fn(Tensor(t[i]))) for i in _b.range(t.shape[0])]
    return Tensor(torch.stack(outs, 0))


def random_uniform(shape_, minval=0, maxval=None, dtype=float32):
    shape_ = [int(s) for s in shape_]
    td = _torch_dtype(dtype)
    if td in (torch.int32, torch.int64):
        x = _rng.integers(int(minval), int(maxval), size=shape_).astype(np.int32)
        random_log.append(("uniform_int", x))
        return Tensor(torch.from_numpy(x).to(td))
    maxval = 1.0 if maxval is None else maxval
    # drawn as float32 in [0,1) in both precision modes: identical uniforms
    x = _rng.random(size=shape_, dtype=np.float32)
    x = (np.float32(minval) + x * np.float32(maxval - minval)).astype(np.float32)
    random_log.append(("uniform_float", x))
    return Tensor(torch.from_numpy(x).to(td))


# --------------------------------------------------------------------------
# tf.nn
# --------------------------------------------------------------------------
class _NN(object):
    """'VALID' cross-correlations, channels-last input, filters [*k, C_in, C_out]."""

    @staticmethod
    def _conv(x, filters, n_dims, padding):
        assert padding == "VALID"
        import torch.nn.functional as F
        t, w = _as_torch(x), _as_torch(filters)
        sp = list(_b.range(1, n_dims + 1))
        t = t.permute([0, n_dims + 1] + sp)                               # N, C, *spatial
        w = w.permute([n_dims + 1, n_dims] + list(_b.range(n_dims)))     # O, I, *k
        out = (F.conv1d, F.conv2d, F.conv3d)[n_dims - 1](t, w)
        return Tensor(out.permute([0] + [d + 1 for d in sp] + [1]).contiguous())

    def conv1d(self, value, filters, stride, padding):
        assert stride == 1
        return self._conv(value, filters, 1, padding)

    def conv2d(self, input, filter, strides, padding):        # noqa: A002
        assert list(strides) == [1] * 4
        return self._conv(input, filter, 2, padding)

    def conv3d(self, input, filter, strides, padding):        # noqa: A002
        assert list(strides) == [1] * 5
        return self._conv(input, filter, 3, padding)


nn = _NN()


# --------------------------------------------------------------------------
# tf.train
# --------------------------------------------------------------------------
class _AdamOptimizer(object):
    """``tf.train.AdamOptimizer`` (TF-1): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g^2; p -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss):
        vs = trainable_variables()
        grads = torch.autograd.grad(_as_torch(loss), [v.t for v in vs], allow_unused=True)
        train.last_gradients = {}
        for v, g in zip(vs, grads):
            if g is None:
                continue
            train.last_gradients[v.name] = g.detach().clone()
            m, s, t = train._slots.get(v.name, (torch.zeros_like(g), torch.zeros_like(g), 0))
            t += 1
            m = self.b1 * m + (1 - self.b1) * g
            s = self.b2 * s + (1 - self.b2) * g * g
            lr_t = self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
            v._set(v.t.detach() - lr_t * m / (torch.sqrt(s) + self.eps))
            train._slots[v.name] = (m, s, t)
        return None


class _Train(object):
    AdamOptimizer = _AdamOptimizer
    _slots = {}
    last_gradients = {}


train = _Train()


class _Summary(object):
    class FileWriter(object):
        def __init__(self, *a, **k):
            pass

        def flush(self):
            pass


summary = _Summary()
