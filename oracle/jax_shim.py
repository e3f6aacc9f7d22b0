"""NumPy-backed stand-in for the handful of JAX symbols the reference's CAVIaR path imports
(oracle side; TEST INFRASTRUCTURE, build container only).

JAX is not installed and not installable here (no wheel, no network).  `install()` registers
fake `jax`, `jax.numpy`, `jax.lax`, `jax.nn`, `jax.random`, `jax.scipy.special`,
`jax.scipy.integrate` modules so that the UNMODIFIED reference sources

    circuitmap/optimise/caviar.py   (imports at :5-12)
    circuitmap/optimise/pava.py     (imports at :4-7)
    circuitmap/simulation.py        (imports at :4-7)
    circuitmap/model.py

can be imported and executed from /root/reference by oracle/make_golden.py, exactly as the
`pytorch_lightning` shim does for the demixer.  The fixtures this produces are therefore outputs
of the reference's own source text (its expression order, its quirks, its control flow); what the
shim supplies underneath is

  * jnp.*            NumPy fp64 (x64 is on in the reference, caviar.py:12) on an ndarray subclass
                     that adds the functional `.at[idx].set(v)` update (out-of-place, OOB scalar
                     indices dropped as XLA scatter does);
  * lax.fori_loop / while_loop / scan, vmap   plain Python loops (the semantics of the traced
                     versions: carries are converted to arrays, vmap maps and stacks pytrees);
  * jit              identity;
  * jax.nn.sigmoid   1 / (1 + exp(-x))  (jax 0.3.x `expit`);
  * ndtr / ndtri     SciPy's (Cephes, as JAX's are);
  * jax.random.*     oracle/prng.py -- the RESTATED threefry stream (pinned to Random123 KATs and
                     to the known answers of JAX's own test-suite, tests/test_oracle_prng.py).

So a fixture made through this shim pins the restatement in oracle/caviar.py against the
reference's own code; it does not (cannot) pin XLA's floating-point reduction order or a live
jax.random.  Nothing outside oracle/make_golden.py and tests/test_jax_shim.py imports this file.
"""
import sys
import types

import numpy as np
import scipy.special as _sps

from . import prng as _prng


# ----------------------------------------------------------------------------- arrays
class JArr(np.ndarray):
    """ndarray with JAX's functional update syntax; in-place operators rebind (arrays are immutable)."""
    __array_priority__ = 100.0

    @property
    def at(self):
        return _At(self)

    def __iadd__(self, o):
        return np.add(self, o)

    def __isub__(self, o):
        return np.subtract(self, o)

    def __imul__(self, o):
        return np.multiply(self, o)

    def __itruediv__(self, o):
        return np.true_divide(self, o)

    def __hash__(self):
        return id(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def _oob(self):
        i = self.idx
        if isinstance(i, (int, np.integer)) or (isinstance(i, np.ndarray) and i.ndim == 0 and i.dtype.kind in "iu"):
            n = self.a.shape[0]
            return not (-n <= int(i) < n)
        return False

    def set(self, v):
        out = np.array(self.a, copy=True).view(JArr)
        if not self._oob():
            out[self.idx] = v
        return out

    def add(self, v):
        out = np.array(self.a, copy=True).view(JArr)
        if not self._oob():
            np.add.at(out, self.idx, v)
        return out


def _wrap_out(x):
    if isinstance(x, np.ndarray) and not isinstance(x, JArr):
        return x.view(JArr)
    if isinstance(x, tuple):
        return tuple(_wrap_out(e) for e in x)
    return x


def _wrap_fn(f):
    def g(*a, **k):
        return _wrap_out(f(*a, **k))
    g.__name__ = getattr(f, "__name__", "wrapped")
    return g


def _asarray(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(JArr) if not isinstance(x, JArr) or dtype is not None \
        else x


def _array(x, dtype=None, copy=True):
    return np.array(x, dtype=dtype, copy=True).view(JArr)


def _unique(x, size=None, fill_value=None, **kw):
    u = np.unique(x, **kw)
    if size is not None:
        if u.size >= size:
            u = u[:size]
        else:
            fv = u.min() if fill_value is None else fill_value
            u = np.concatenate([u, np.full(size - u.size, fv, dtype=u.dtype)])
    return u.view(JArr)


class _NumpyFacade(types.ModuleType):
    """jax.numpy: every NumPy function, results viewed as JArr."""

    def __init__(self, name, base, overrides=None):
        super().__init__(name)
        self._base = base
        for k, v in (overrides or {}).items():
            setattr(self, k, v)

    def __getattr__(self, name):
        obj = getattr(self._base, name)
        if isinstance(obj, type) or not callable(obj):
            return obj
        w = _wrap_fn(obj)
        setattr(self, name, w)
        return w


# ----------------------------------------------------------------------------- pytrees / control flow
def _leaf_to_array(x):
    """What tracing does to a loop carry: Python scalars become (x64) array scalars."""
    if isinstance(x, (tuple, list)):
        return type(x)(_leaf_to_array(e) for e in x) if isinstance(x, tuple) else [_leaf_to_array(e) for e in x]
    if isinstance(x, bool):
        return np.bool_(x)
    if isinstance(x, int):
        return np.int64(x)
    if isinstance(x, float):
        return np.float64(x)
    if isinstance(x, np.ndarray) and not isinstance(x, JArr):
        return x.view(JArr)
    return x


def _tree_stack(items):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return tuple(_tree_stack([it[j] for it in items]) for j in range(len(first)))
    if first is None:
        return None
    return np.stack([np.asarray(it) for it in items]).view(JArr)


def fori_loop(lower, upper, body_fun, init_val):
    val = _leaf_to_array(init_val)
    for i in range(int(lower), int(upper)):
        val = _leaf_to_array(body_fun(np.int64(i), val))
    return val


def while_loop(cond_fun, body_fun, init_val):
    val = _leaf_to_array(init_val)
    while bool(cond_fun(val)):
        val = _leaf_to_array(body_fun(val))
    return val


def scan(f, init, xs, length=None):
    carry = _leaf_to_array(init)
    ys = []
    n = len(xs) if xs is not None else int(length)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        carry = _leaf_to_array(carry)
        ys.append(y)
    return carry, _tree_stack(ys)


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), "vmap shim: in_axes / args mismatch"
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = np.shape(a)[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else _wrap_out(np.take(np.asarray(a), i, axis=ax)) for a, ax in zip(args, axes)]
            outs.append(fun(*sl))
        return _tree_stack(outs)
    return mapped


def grad(fun, *a, **k):
    def g(*args, **kwargs):
        raise NotImplementedError("jax shim: grad is not provided (not on the CAVIaR path)")
    return g


# ----------------------------------------------------------------------------- nn / special / random
def sigmoid(x):
    with np.errstate(over="ignore"):
        return _wrap_out(np.asarray(1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))))


def _random_choice(key, a, shape=(), replace=True, p=None):
    assert not replace and p is None, "jax shim: only choice(..., replace=False) is provided"
    n = int(a) if np.ndim(a) == 0 else len(a)
    shape = tuple(np.atleast_1d(shape).astype(int))
    perm = _prng.permutation(np.asarray(key, dtype=np.uint32), n)
    assert shape == (n,), "jax shim: choice(key, N, [N], replace=False) only"
    return (perm if np.ndim(a) == 0 else np.asarray(a)[perm]).view(JArr)


def _random_uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
    u = _prng.uniform_f64(np.asarray(key, dtype=np.uint32), tuple(int(s) for s in shape))
    return (u * (maxval - minval) + minval).view(JArr)


def install():
    """Register the fake modules in sys.modules (idempotent).  Refuses to shadow a real JAX."""
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__circuitmap_shim__", False):
        raise RuntimeError("a real jax is importable; use it instead of the shim")
    jax = types.ModuleType("jax")
    jax.__circuitmap_shim__ = True

    linalg = _NumpyFacade("jax.numpy.linalg", np.linalg)
    jnp = _NumpyFacade("jax.numpy", np, dict(array=_array, asarray=_asarray, unique=_unique, linalg=linalg))

    lax = types.ModuleType("jax.lax")
    lax.fori_loop, lax.while_loop, lax.scan = fori_loop, while_loop, scan

    nn = types.ModuleType("jax.nn")
    nn.sigmoid = sigmoid

    random = types.ModuleType("jax.random")
    random.PRNGKey = lambda seed: _prng.prng_key(seed).view(JArr)
    random.split = lambda key, num=2: _prng.split(np.asarray(key, dtype=np.uint32), num).view(JArr)
    random.uniform = _random_uniform
    random.choice = _random_choice

    jsp = types.ModuleType("jax.scipy")
    special = types.ModuleType("jax.scipy.special")
    special.ndtr = _wrap_fn(_sps.ndtr)
    special.ndtri = _wrap_fn(_sps.ndtri)
    integrate = types.ModuleType("jax.scipy.integrate")
    integrate.trapezoid = _wrap_fn(np.trapezoid)
    jsp.special, jsp.integrate = special, integrate

    class _Config:
        def update(self, *a, **k):
            pass

    jax.numpy, jax.lax, jax.nn, jax.random, jax.scipy = jnp, lax, nn, random, jsp
    jax.jit, jax.vmap, jax.grad = jit, vmap, grad
    jax.config = _Config()
    jax.devices = lambda *a, **k: ["cpu:0"]
    jax.device_put = lambda x, device=None: _asarray(x)
    for name, mod in [("jax", jax), ("jax.numpy", jnp), ("jax.numpy.linalg", linalg), ("jax.lax", lax),
                      ("jax.nn", nn), ("jax.random", random), ("jax.scipy", jsp),
                      ("jax.scipy.special", special), ("jax.scipy.integrate", integrate)]:
        sys.modules[name] = mod
    return jax


def install_lightning_shim():
    """`pytorch_lightning` stand-in (LightningModule := nn.Module + load_from_checkpoint) so that
    circuitmap/__init__.py imports (neural_waveform_demixing.py:4)."""
    import torch
    if "pytorch_lightning" in sys.modules:
        return sys.modules["pytorch_lightning"]
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        @classmethod
        def _load(cls, path):
            m = cls()
            ck = torch.load(path, map_location="cpu", weights_only=True)
            m.load_state_dict(ck["state_dict"])
            return m

        def load_from_checkpoint(self, path):       # the reference calls it on an instance (nwd.py:27)
            return type(self)._load(path)

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.Trainer = object
    sys.modules["pytorch_lightning"] = pl
    return pl


def import_reference(ref_root="/root/reference"):
    """Import the unmodified reference package `circuitmap` from the read-only checkout."""
    install()
    install_lightning_shim()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import circuitmap                                  # noqa: E402  (the reference, not circuitmap_b200)
    assert circuitmap.__file__.startswith(ref_root)
    return circuitmap
