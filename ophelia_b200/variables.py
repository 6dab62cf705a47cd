"""Variable scopes and the parameter store (the slice of tf.variable_scope / tf.get_variable the path needs).

Variable names and layouts equal the reference's checkpoint names (`train.py:194`):
`Text2Mel/TextEnc/C_2/conv1d/kernel [k,Cin,Cout]`, `.../normalize/{beta,gamma}`, `.../HC_4/H1/gamma`,
`SSRN/D_4/conv2d_transpose/kernel [1,3,Cout,Cin]`, `.../embed_1/lookup_table`.

All trainable values of a store live in ONE flat fp32 device buffer (plus flat grad / Adam m / v buffers of the
same layout): one memset, one allreduce and one fused clip+Adam launch per step.
"""
import contextlib
import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

_scope_stack = []
_current_store = None


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope_stack.append(name)
    try:
        yield
    finally:
        _scope_stack.pop()


def current_scope():
    return "/".join(_scope_stack)


def scoped(name):
    s = current_scope()
    return s + "/" + name if s else name


@contextlib.contextmanager
def use_store(store):
    global _current_store
    prev, _current_store = _current_store, store
    saved_stack = list(_scope_stack)
    del _scope_stack[:]
    try:
        yield store
    finally:
        _current_store = prev
        _scope_stack[:] = saved_stack


def get_store():
    if _current_store is None:
        raise RuntimeError("no active VariableStore (use `with use_store(store):`)")
    return _current_store


def _trunc_normal(rng, shape, std):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def init_value(kind, shape, rng):
    """Reference initialisers: truncated_normal(0.1) embeddings (modules.py:37); variance_scaling_initializer()
    = truncated normal, sigma = sqrt(1.3 * 2 / fan_in) (modules.py:132, 250); zeros / ones for bias, beta / gamma."""
    if kind == "embed":
        return _trunc_normal(rng, shape, 0.1)
    if kind == "kernel":                      # [k, Cin, Cout]
        return _trunc_normal(rng, shape, math.sqrt(1.3 * 2.0 / (shape[0] * shape[1])))
    if kind == "kernel_t":                    # [1, 3, Cout, Cin]
        return _trunc_normal(rng, shape, math.sqrt(1.3 * 2.0 / (shape[0] * shape[1] * shape[2])))
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    if kind == "ones":
        return np.ones(shape, np.float32)
    raise ValueError(kind)


class VariableStore(object):
    def __init__(self, device="cuda:0", seed=0):
        self.device = torch.device(device)
        self.rng = np.random.default_rng(seed)
        self.specs = OrderedDict()     # name -> (shape, kind)
        self._host = OrderedDict()     # name -> np.ndarray until finalize()
        self.vars = OrderedDict()      # name -> view into self.flat
        self.grads = OrderedDict()
        self.offsets = OrderedDict()
        self.flat = self.grad_flat = self.m_flat = self.v_flat = None
        self.version = 0               # bumped whenever values change (packed weight images re-pack lazily)
        self.packed = {}
        self.dropout_seed = mix_dropout_seed(seed, 0)

    # ---- declaration
    def declare(self, name, shape, kind):
        shape = tuple(int(s) for s in shape)
        if name in self.specs:
            assert self.specs[name][0] == shape, "variable %s re-declared with another shape" % name
            return
        self.specs[name] = (shape, kind)
        self._host[name] = init_value(kind, shape, self.rng)

    def declare_all(self, specs):
        for name, shape, kind in specs:
            self.declare(name, shape, kind)
        return self

    def finalize(self, with_optimizer=False):
        """(Re)build the flat buffers so that they hold every declared variable.  Graph constructors declare the
        whole inventory first, so this normally runs once; ad-hoc layer calls may grow the store later."""
        if self._host:
            old = self.state_dict() if self.flat is not None else {}
            assert self.m_flat is None or int(self.global_step.item()) == 0, "cannot add variables after training started"
            off = 0
            self.offsets = OrderedDict()
            for name, (shape, _k) in self.specs.items():
                self.offsets[name] = off
                off += (int(np.prod(shape)) + 3) // 4 * 4          # keep every variable 16-byte aligned
            self.numel = off
            host = np.zeros(off, np.float32)
            for name, (shape, _k) in self.specs.items():
                o = self.offsets[name]
                src = self._host[name] if name in self._host else old[name]
                host[o:o + int(np.prod(shape))] = src.reshape(-1)
            self.flat = torch.from_numpy(host).to(self.device)
            self.grad_flat = torch.zeros_like(self.flat)
            self.vars, self.grads = OrderedDict(), OrderedDict()
            for name, (shape, _k) in self.specs.items():
                o, n = self.offsets[name], int(np.prod(shape))
                self.vars[name] = self.flat[o:o + n].view(shape)
                self.grads[name] = self.grad_flat[o:o + n].view(shape)
            self._host = OrderedDict()
            self.packed = {}
            self._pack_plan = None
            self.version += 1
            if self.m_flat is not None:
                self._alloc_opt()
        if with_optimizer and self.m_flat is None:
            self._alloc_opt()
        return self

    def _alloc_opt(self):
        self.m_flat = torch.zeros_like(self.flat)
        self.v_flat = torch.zeros_like(self.flat)
        self.global_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.lr_t = torch.zeros(2, dtype=torch.float32, device=self.device)

    # ---- packed conv kernels: one launch re-packs all of them after an optimiser step
    def repack_all(self):
        import ctypes
        from . import _lib
        if not self.packed:
            return
        key = tuple(self.packed)
        plan = getattr(self, "_pack_plan", None)
        if plan is None or plan[0] != key:
            lib = _lib.load()
            jb = int(lib.oph_pack_job_bytes())
            cap = 3 * len(self.packed)
            host = ctypes.create_string_buffer(cap * jb)
            njobs, nblocks = ctypes.c_int(0), ctypes.c_longlong(0)
            for pk in self.packed.values():
                _lib.call("oph_pack_plan_add", ctypes.cast(host, ctypes.c_void_p), cap, ctypes.byref(njobs),
                          ctypes.byref(nblocks), pk.w.data_ptr(), pk.k, pk.cin, pk.cout, int(pk.deconv),
                          pk.fwd.data_ptr(), pk.bwd.data_ptr() if pk.bwd is not None else None)
            dev = torch.frombuffer(bytearray(host.raw[:njobs.value * jb]), dtype=torch.uint8).to(self.device)
            plan = (key, dev, njobs.value, nblocks.value)
            self._pack_plan = plan
        _lib.call("oph_pack_run", plan[1].data_ptr(), plan[2], plan[3], torch.cuda.current_stream().cuda_stream)
        for pk in self.packed.values():
            pk.version = self.version

    # ---- access
    def get(self, name):
        return self.vars[name]

    def grad(self, name):
        return self.grads[name]

    def num_parameters(self):
        return sum(int(np.prod(s)) for s, _ in self.specs.values())

    def names(self, prefix=""):
        return [n for n in self.specs if n.startswith(prefix)]

    # ---- (de)serialisation: plain name -> array dict, the exchange format for a TF-checkpoint importer
    def state_dict(self):
        return OrderedDict((n, v.detach().cpu().numpy().copy()) for n, v in self.vars.items())

    def load_state_dict(self, values, strict=True):
        self.finalize()
        for n, v in values.items():
            if n not in self.vars:
                if strict:
                    raise KeyError(n)
                continue
            t = torch.as_tensor(np.asarray(v, dtype=np.float32))
            assert tuple(t.shape) == tuple(self.vars[n].shape), (n, t.shape, self.vars[n].shape)
            self.vars[n].copy_(t.to(self.device))
        self.version += 1

    def save(self, path):
        np.savez(path, **self.state_dict())

    def load(self, path, strict=True):
        with np.load(path) as z:
            self.load_state_dict({k: z[k] for k in z.files}, strict=strict)


def layer_seed(name, store=None):
    """Per-layer dropout seed: stable per layer name, offset by the store's `dropout_seed` (hp.seed and the data-parallel
    rank mixed by the Graph constructor) so that runs with another seed and the replicas of one run draw different masks;
    the kernels add global_step on top."""
    return (zlib.crc32(name.encode()) * 2654435761 + (getattr(store, "dropout_seed", 0) if store is not None else 0)) % (1 << 62)


def mix_dropout_seed(seed, rank):
    return (int(seed) * 0x9E3779B97F4A7C15 + int(rank) * 0xBF58476D1CE4E5B9) % (1 << 62)
