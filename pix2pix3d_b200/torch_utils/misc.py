"""Small helpers the hot-path modules call (reference torch_utils/misc.py:84-107, 150-176)."""
import contextlib
import re
import warnings

import torch


def assert_shape(tensor, ref_shape):
    """Check rank and every non-None entry of `ref_shape` (reference torch_utils/misc.py:84)."""
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref) in enumerate(zip(tensor.shape, ref_shape)):
        if ref is None:
            continue
        if int(size) != int(ref):
            raise AssertionError(f'Wrong size for dimension {idx}: got {int(size)}, expected {int(ref)}')


@contextlib.contextmanager
def suppress_tracer_warnings():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=torch.jit.TracerWarning)
        yield


def profiled_function(fn):
    """Wrap `fn` in a profiler scope named after it."""
    def wrapped(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


def params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.parameters()) + list(module.buffers())


def named_params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.named_parameters()) + list(module.named_buffers())


def copy_params_and_buffers(src_module, dst_module, require_all=False, allow_mismatch=False):
    """Name-matched tensor copy between modules (reference torch_utils/misc.py:157-176)."""
    src = dict(named_params_and_buffers(src_module))
    for name, tensor in named_params_and_buffers(dst_module):
        # pix2pix3D addition (:163-165): `superresolution_semantic.*` is initialised from `superresolution.*` when the source
        # (an EG3D checkpoint) has no semantic branch
        name_src = name if name in src else name.replace('_semantic', '')
        if name_src not in src:
            if require_all:
                raise AssertionError(f'{name} missing in source module')
            print(f'Warning: {name} not found in source module')
            continue
        s = src[name_src].detach()
        if s.shape != tensor.shape and allow_mismatch:
            print(f'Warning: {name_src} shape mismatch: {tuple(s.shape)} vs {tuple(tensor.shape)}')
            continue
        tensor.copy_(s).requires_grad_(tensor.requires_grad)


# ---------------------------------------------------------------------------------------------
# Host-side helpers of the reference's training loop (torch_utils/misc.py:24-46, 113-144, 183-262): cached constants,
# NaN scrubbing of the flat gradient, the infinite data-order sampler, DDP consistency check, module summary table.
# ---------------------------------------------------------------------------------------------
import numpy as np

nan_to_num = torch.nan_to_num
symbolic_assert = torch._assert            # pylint: disable=protected-access

_constants = {}


def constant(value, shape=None, dtype=None, device=None, memory_format=None):
    """Tensor for a Python / numpy constant, built once per (value, shape, dtype, device, layout) and reused."""
    value = np.asarray(value)
    shape = tuple(shape) if shape is not None else None
    dtype = dtype or torch.get_default_dtype()
    device = device or torch.device('cpu')
    memory_format = memory_format or torch.contiguous_format
    key = (value.shape, value.dtype, value.tobytes(), shape, dtype, device, memory_format)
    t = _constants.get(key)
    if t is None:
        t = torch.as_tensor(value.copy(), dtype=dtype, device=device)
        if shape is not None:
            t = t.expand(torch.broadcast_shapes(t.shape, shape))
        t = t.contiguous(memory_format=memory_format)
        _constants[key] = t
    return t


class InfiniteSampler(torch.utils.data.Sampler):
    """Endless, per-rank strided walk over a shuffled index order that keeps re-shuffling locally: after visiting position i
    the entry is swapped with one up to `window_size * N` positions back. The RandomState call sequence (one `shuffle`, then one
    `randint(window)` per visited position, on every rank) is what fixes the data order for a seed (reference misc.py:113-144)."""

    def __init__(self, dataset, rank=0, num_replicas=1, shuffle=True, seed=0, window_size=0.5):
        assert len(dataset) > 0 and num_replicas > 0 and 0 <= rank < num_replicas and 0 <= window_size <= 1
        super().__init__()
        self.dataset, self.rank, self.num_replicas = dataset, rank, num_replicas
        self.shuffle, self.seed, self.window_size = shuffle, seed, window_size

    def __iter__(self):
        n = len(self.dataset)
        order = np.arange(n)
        rnd, window = None, 0
        if self.shuffle:
            rnd = np.random.RandomState(self.seed)
            rnd.shuffle(order)
            window = int(np.rint(n * self.window_size))
        pos = 0
        while True:
            i = pos % n
            if pos % self.num_replicas == self.rank:
                yield order[i]
            if window >= 2:
                j = (i - rnd.randint(window)) % n
                order[i], order[j] = order[j], order[i]
            pos += 1


@contextlib.contextmanager
def ddp_sync(module, sync):
    assert isinstance(module, torch.nn.Module)
    if sync or not isinstance(module, torch.nn.parallel.DistributedDataParallel):
        yield
    else:
        with module.no_sync():
            yield


def check_ddp_consistency(module, ignore_regex=None):
    """Every parameter and buffer equals rank 0's copy (one broadcast per tensor, as the reference)."""
    assert isinstance(module, torch.nn.Module)
    for name, tensor in named_params_and_buffers(module):
        fullname = type(module).__name__ + '.' + name
        if ignore_regex is not None and re.fullmatch(ignore_regex, fullname):
            continue
        tensor = tensor.detach()
        if tensor.is_floating_point():
            tensor = nan_to_num(tensor)
        other = tensor.clone()
        torch.distributed.broadcast(tensor=other, src=0)
        assert (tensor == other).all(), fullname


def print_module_summary(module, inputs, max_nesting=3, skip_redundant=True):
    """Run `module(*inputs)` once and print one row per sub-module call (parameters, buffers, output shape, dtype)."""
    assert isinstance(module, torch.nn.Module) and isinstance(inputs, (tuple, list))
    calls, depth = [], [0]

    def enter(_m, _i):
        depth[0] += 1

    def leave(m, _i, out):
        depth[0] -= 1
        if depth[0] <= max_nesting:
            outs = [t for t in (list(out) if isinstance(out, (tuple, list)) else [out]) if isinstance(t, torch.Tensor)]
            calls.append((m, outs))

    hooks = [m.register_forward_pre_hook(enter) for m in module.modules()] + [m.register_forward_hook(leave) for m in module.modules()]
    outputs = module(*inputs)
    for h in hooks:
        h.remove()
    seen, rows = set(), []
    names = {m: n for n, m in module.named_modules()}
    tot_p = tot_b = 0
    for m, outs in calls:
        fresh_p = [t for t in m.parameters() if id(t) not in seen]
        fresh_b = [t for t in m.buffers() if id(t) not in seen]
        fresh_o = [t for t in outs if id(t) not in seen]
        seen |= {id(t) for t in fresh_p + fresh_b + fresh_o}
        if skip_redundant and not (fresh_p or fresh_b or fresh_o):
            continue
        n_p, n_b = sum(t.numel() for t in fresh_p), sum(t.numel() for t in fresh_b)
        tot_p, tot_b = tot_p + n_p, tot_b + n_b
        name = '<top-level>' if m is module else names[m]
        shapes = [str(list(t.shape)) for t in outs] or ['-']
        dtypes = [str(t.dtype).split('.')[-1] for t in outs] or ['-']
        rows.append([name + (':0' if len(outs) >= 2 else ''), str(n_p) if n_p else '-', str(n_b) if n_b else '-', shapes[0], dtypes[0]])
        for k in range(1, len(outs)):
            rows.append([f'{name}:{k}', '-', '-', shapes[k], dtypes[k]])
    head = [type(module).__name__, 'Parameters', 'Buffers', 'Output shape', 'Datatype']
    rule = ['---'] * len(head)
    table = [head, rule] + rows + [rule, ['Total', str(tot_p), str(tot_b), '-', '-']]
    widths = [max(len(r[c]) for r in table) for c in range(len(head))]
    print()
    for r in table:
        print('  '.join(cell + ' ' * (w - len(cell)) for cell, w in zip(r, widths)))
    print()
    return outputs
