"""`persistent_class` decorator with the reference's surface (torch_utils/persistence.py:37-128).

Classes keep a record of their constructor arguments (`init_args`, `init_kwargs`) so that callers such as
training_loop.py / legacy.py can rebuild them. Objects of this package pickle by reference to the package (no source is
embedded). Pickles WRITTEN BY THE REFERENCE embed the source of each persistent class's module and name
`torch_utils.persistence._reconstruct_persistent_obj` as their constructor (reference persistence.py:119-128, 181-204);
`_reconstruct_persistent_obj` below accepts that format and rebuilds the object as the mirror class of the same name from
its recorded constructor arguments, then loads the pickled parameters and buffers -- what the reference's callers do by hand
with `reload_modules` (`generate_samples.py`, `misc.copy_params_and_buffers`) -- so a released checkpoint comes up on the
kernels of this package without executing the code stored inside it.
"""
import copy
import importlib
import re
import sys

_version = 6                 # pickle format version of the reference's persistence module (persistence.py:29)
_import_hooks = []
# set to True to fall back to executing the module source stored in the pickle (the reference's behaviour) for classes this
# package has no mirror of
allow_embedded_source = False
_MIRROR_MODULES = ('training.networks_stylegan2', 'training.triplane', 'training.triplane_cond', 'training.superresolution',
                   'training.dual_discriminator')

_decorated = set()


def persistent_class(orig_class):
    assert isinstance(orig_class, type)
    if orig_class in _decorated:
        return orig_class

    class Decorator(orig_class):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            # only the outermost decorated class of an instance records the arguments
            if getattr(self, '_init_args', None) is None or type(self) is Decorator:
                self._init_args = copy.deepcopy(args)
                self._init_kwargs = copy.deepcopy(kwargs)

        @property
        def init_args(self):
            return copy.deepcopy(self._init_args)

    # returns an EasyDict like the reference; dnnlib is imported lazily
    def _init_kwargs(self):
        from .. import dnnlib
        return dnnlib.EasyDict(copy.deepcopy(self._init_kwargs))
    Decorator.init_kwargs = property(_init_kwargs)

    Decorator.__name__ = orig_class.__name__
    Decorator.__qualname__ = orig_class.__qualname__
    Decorator.__module__ = orig_class.__module__
    Decorator.__doc__ = orig_class.__doc__
    _decorated.add(Decorator)
    return Decorator


def is_persistent(obj):
    try:
        if obj in _decorated:
            return True
    except TypeError:
        pass
    return type(obj) in _decorated


def import_hook(hook):
    """Register `hook(meta) -> meta`, called for every persistent object being unpickled (reference persistence.py:150-176)."""
    assert callable(hook)
    _import_hooks.append(hook)
    return hook


def _mirror_class(class_name, module_src):
    """The mirror class for a pickled (class name, module source) pair. Names are unique per module but not across modules
    (`TriPlaneGenerator` lives in both training/triplane.py and training/triplane_cond.py), so the candidate whose module
    defines the largest share of the classes declared in the stored source wins."""
    declared = set(re.findall(r'^class\s+(\w+)', module_src or '', flags=re.M))
    root = __name__.rsplit('.', 2)[0]
    best, best_score = None, -1.0
    for mod_name in _MIRROR_MODULES:
        mod = importlib.import_module(f'{root}.{mod_name}')
        cls = getattr(mod, class_name, None)
        if not isinstance(cls, type):
            continue
        own = {k for k, v in vars(mod).items() if isinstance(v, type) and v.__module__ == mod.__name__}
        score = len(declared & own) / max(len(declared), 1)
        if score > best_score:
            best, best_score = cls, score
    return best


def _flatten_state(state, prefix, out):
    """Parameters and buffers of a pickled torch.nn.Module state (its __dict__), children included."""
    for kind in ('_parameters', '_buffers'):
        for k, v in (state.get(kind) or {}).items():
            if v is not None:
                out[prefix + k] = v
    skip = set(state.get('_non_persistent_buffers_set') or ())
    for k in skip:
        out.pop(prefix + k, None)
    for k, child in (state.get('_modules') or {}).items():
        if child is not None:
            for name, t in child.state_dict().items():
                out[f'{prefix}{k}.{name}'] = t


_MODULE_INTERNALS = {'training', '_parameters', '_buffers', '_modules', '_non_persistent_buffers_set', '_init_args',
                     '_init_kwargs', '_orig_module_src', '_orig_class_name', '_is_full_backward_hook'}


def _is_plain(v, depth=0):
    if v is None or isinstance(v, (bool, int, float, str, bytes)):
        return True
    if depth > 6:
        return False
    if isinstance(v, (list, tuple, set, frozenset)):
        return all(_is_plain(e, depth + 1) for e in v)
    if isinstance(v, dict):
        return all(_is_plain(k, depth + 1) and _is_plain(e, depth + 1) for k, e in v.items())
    try:
        import numpy as np
        import torch
        return isinstance(v, (np.ndarray, np.generic, torch.Tensor, torch.dtype, torch.memory_format))
    except ImportError:
        return False


def _restore_plain_attributes(obj, state):
    """The reference restores the whole pickled `__dict__` (persistence.py:197-203), so attributes mutated after
    construction survive a save/load cycle: `G.neural_rendering_resolution` (training_loop.py:558 copies it onto G_ema
    before every snapshot), `rendering_kwargs`, `_last_planes`. The mirror is rebuilt through its constructor, so those
    plain (non-module, non-hook) attributes are copied over afterwards, children included."""
    src = state if isinstance(state, dict) else getattr(state, '__dict__', {})
    for k, v in src.items():
        if k in _MODULE_INTERNALS or k.endswith('_hooks') or k.endswith('_hooks_with_kwargs') or k.endswith('_hooks_always_called'):
            continue
        if k.startswith('_') and k not in ('_last_planes',):
            continue
        cur = obj.__dict__.get(k, None)
        if k in obj.__dict__ and not _is_plain(cur):
            continue                                    # helper objects the mirror builds itself (renderer, ray sampler, ...)
        if _is_plain(v):
            obj.__dict__[k] = copy.deepcopy(v) if isinstance(v, (dict, list, set)) else v
    dst_children = obj.__dict__.get('_modules') or {}
    for k, child in (src.get('_modules') or {}).items():
        if child is not None and dst_children.get(k) is not None:
            _restore_plain_attributes(dst_children[k], child.__dict__)


def _reconstruct_persistent_obj(meta):
    """Constructor named by the reference's pickles (persistence.py:181-204)."""
    from .. import dnnlib
    meta = dnnlib.EasyDict(meta)
    meta.state = dnnlib.EasyDict(meta.state)
    for hook in _import_hooks:
        meta = hook(meta)
        assert meta is not None
    assert meta.version == _version, f'unsupported persistence version {meta.version}'
    assert meta.type == 'class'
    cls = _mirror_class(meta.class_name, meta.get('module_src'))
    if cls is None:
        if not allow_embedded_source:
            raise ModuleNotFoundError(f'no mirror of persistent class {meta.class_name!r}; set '
                                      'torch_utils.persistence.allow_embedded_source = True to execute the source stored in the pickle')
        return _reconstruct_from_source(meta)
    import torch
    if issubclass(cls, torch.nn.Module) and '_init_kwargs' in meta.state:
        obj = cls(*meta.state.get('_init_args', ()), **meta.state['_init_kwargs'])
        tensors = {}
        _flatten_state(meta.state, '', tensors)
        missing, unexpected = obj.load_state_dict(tensors, strict=False)
        missing = [k for k in missing if k in dict(obj.named_parameters())]      # buffers may be re-derived constants
        if missing or unexpected:
            raise RuntimeError(f'{meta.class_name}: checkpoint does not match the mirror class '
                               f'(missing {missing[:4]}, unexpected {list(unexpected)[:4]})')
        _restore_plain_attributes(obj, meta.state)
        obj.train(bool(meta.state.get('training', True)))
        requires = {k: v.requires_grad for k, v in tensors.items() if hasattr(v, 'requires_grad')}
        for name, p in obj.named_parameters():
            if name in requires:
                p.requires_grad_(requires[name])
        return obj
    obj = cls.__new__(cls)
    setstate = getattr(obj, '__setstate__', None)
    if callable(setstate):
        setstate(meta.state)
    else:
        obj.__dict__.update(meta.state)
    return obj


_src_modules = {}


def _reconstruct_from_source(meta):
    """The reference's own procedure: run the stored module source and instantiate its class without calling __init__."""
    import types
    import uuid
    module = _src_modules.get(meta.module_src)
    if module is None:
        module = types.ModuleType('_imported_module_' + uuid.uuid4().hex)
        sys.modules[module.__name__] = module
        _src_modules[meta.module_src] = module
        exec(meta.module_src, module.__dict__)          # noqa: S102 -- opt-in, see allow_embedded_source
    cls = persistent_class(module.__dict__[meta.class_name])
    obj = cls.__new__(cls)
    setstate = getattr(obj, '__setstate__', None)
    if callable(setstate):
        setstate(meta.state)
    else:
        obj.__dict__.update(meta.state)
    return obj
