"""Attribute dictionaries, by-name object construction and `open_url` (reference dnnlib/util.py:41-58, 262-310, 373-480)."""
import importlib
from typing import Any


class EasyDict(dict):
    """dict whose items are also attributes."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __delattr__(self, name: str) -> None:
        del self[name]


_PKG = __name__.rsplit('.', 2)[0]          # 'pix2pix3d_b200'
_MIRRORED_ROOTS = ('training', 'torch_utils', 'dnnlib', 'camera_utils')


def get_obj_by_name(name: str) -> Any:
    """Resolve 'pkg.mod.attr.sub' by trying the longest importable module prefix first. Names under the
    reference's top-level packages ('training.superresolution.X', ...) resolve inside this package."""
    if name.split('.', 1)[0] in _MIRRORED_ROOTS and not name.startswith(_PKG + '.'):
        try:
            return get_obj_by_name(f'{_PKG}.{name}')
        except ImportError:
            pass
    parts = name.split('.')
    last_err = None
    for cut in range(len(parts) - 1, 0, -1):
        mod_name = '.'.join(parts[:cut])
        try:
            obj = importlib.import_module(mod_name)
        except ModuleNotFoundError as err:
            # only swallow "this prefix is not a module"; propagate failures inside a real module
            if err.name is not None and not mod_name.startswith(err.name):
                raise
            last_err = err
            continue
        try:
            for attr in parts[cut:]:
                obj = getattr(obj, attr)
            return obj
        except AttributeError as err:
            last_err = err
    raise ImportError(f'cannot resolve {name!r}') from last_err


def call_func_by_name(*args, func_name: str = None, **kwargs) -> Any:
    assert func_name is not None
    fn = get_obj_by_name(func_name)
    assert callable(fn)
    return fn(*args, **kwargs)


def construct_class_by_name(*args, class_name: str = None, **kwargs) -> Any:
    return call_func_by_name(*args, func_name=class_name, **kwargs)


def is_url(obj: Any, allow_file_urls: bool = False) -> bool:
    """True for strings of the form scheme://host/... (reference dnnlib/util.py:373-395)."""
    import urllib.parse
    if not isinstance(obj, str) or '://' not in obj:
        return False
    if allow_file_urls and obj.startswith('file://'):
        return True
    try:
        parts = urllib.parse.urlparse(obj)
        return bool(parts.scheme) and bool(parts.netloc) and '.' in parts.netloc
    except ValueError:
        return False


def open_url(url: str, cache_dir: str = None, num_attempts: int = 10, verbose: bool = True, return_filename: bool = False,
             cache: bool = True) -> Any:
    """Binary file object (or file name) for a local path, a file:// URL or an http(s) URL -- the call the reference's scripts
    wrap around `legacy.load_network_pkl` (dnnlib/util.py:398-480, `generate_samples.py:93`). Downloads are cached by URL
    hash under `cache_dir` (default ~/.cache/dnnlib/downloads)."""
    import hashlib
    import io
    import os
    import re
    import urllib.parse
    import urllib.request
    assert num_attempts >= 1
    assert not (return_filename and not cache)
    if not re.match('^[a-z]+://', url):
        return url if return_filename else open(url, 'rb')
    if url.startswith('file://'):
        filename = urllib.parse.unquote(urllib.parse.urlparse(url).path)
        if re.match(r'^/[a-zA-Z]:', filename):
            filename = filename[1:]
        return filename if return_filename else open(filename, 'rb')
    assert is_url(url), url
    cache_dir = cache_dir or os.path.join(os.path.expanduser('~'), '.cache', 'dnnlib', 'downloads')
    stem = hashlib.md5(url.encode('utf-8')).hexdigest() + '_' + re.sub(r'[^0-9a-zA-Z-._]', '_', url.rsplit('/', 1)[-1])[:128]
    cached = os.path.join(cache_dir, stem)
    if cache and os.path.isfile(cached):
        return cached if return_filename else open(cached, 'rb')
    err = None
    for attempt in range(num_attempts):
        try:
            if verbose:
                print(f'Downloading {url} ...', flush=True)
            with urllib.request.urlopen(url) as r:
                data = r.read()
            break
        except KeyboardInterrupt:
            raise
        except Exception as e:        # noqa: BLE001 -- retried, re-raised after the last attempt
            err = e
    else:
        raise IOError(f'failed to download {url}') from err
    if cache:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = cached + '.tmp.' + hashlib.md5(os.urandom(8)).hexdigest()
        with open(tmp, 'wb') as f:
            f.write(data)
        os.replace(tmp, cached)
        if return_filename:
            return cached
    return io.BytesIO(data)


# ---------------------------------------------------------------------------------------------
# Host-side conveniences the reference's launch / training scripts call (train.py:63-66, training_loop.py:581,
# metrics/metric_utils.py): console tee, cache directory, time formatting, by-name function lookup. Surface of the
# reference's dnnlib/util.py:58-170, 238-330.
# ---------------------------------------------------------------------------------------------
import os as _os
import sys as _sys
import tempfile as _tempfile


class Logger:
    """Tee of stdout+stderr into an optional file; restores both streams on close (usable as a context manager)."""

    def __init__(self, file_name: str = None, file_mode: str = 'w', should_flush: bool = True):
        self.file = open(file_name, file_mode) if file_name is not None else None
        self.should_flush = should_flush
        self.stdout, self.stderr = _sys.stdout, _sys.stderr
        _sys.stdout = _sys.stderr = self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def write(self, text):
        if isinstance(text, bytes):
            text = text.decode()
        if not text:
            return
        if self.file is not None:
            self.file.write(text)
        self.stdout.write(text)
        if self.should_flush:
            self.flush()

    def flush(self):
        if self.file is not None:
            self.file.flush()
        self.stdout.flush()

    def close(self):
        self.flush()
        if _sys.stdout is self:
            _sys.stdout = self.stdout
        if _sys.stderr is self:
            _sys.stderr = self.stderr
        if self.file is not None:
            self.file.close()
            self.file = None


_cache_dir = None


def set_cache_dir(path: str) -> None:
    global _cache_dir
    _cache_dir = path


def make_cache_dir_path(*paths: str) -> str:
    for base in (_cache_dir, _os.environ.get('DNNLIB_CACHE_DIR')):
        if base is not None:
            return _os.path.join(base, *paths)
    for var in ('HOME', 'USERPROFILE'):
        if var in _os.environ:
            return _os.path.join(_os.environ[var], '.cache', 'dnnlib', *paths)
    return _os.path.join(_tempfile.gettempdir(), '.cache', 'dnnlib', *paths)


def _dhms(seconds):
    s = int(round(float(seconds)))
    return s // 86400, (s // 3600) % 24, (s // 60) % 60, s % 60, s


def format_time(seconds) -> str:
    d, h, m, s, total = _dhms(seconds)
    if total < 60:
        return f'{total}s'
    if total < 3600:
        return f'{total // 60}m {s:02}s'
    if total < 86400:
        return f'{total // 3600}h {m:02}m {s:02}s'
    return f'{d}d {h:02}h {m:02}m'


def format_time_brief(seconds) -> str:
    d, h, m, s, total = _dhms(seconds)
    if total < 60:
        return f'{total}s'
    if total < 3600:
        return f'{total // 60}m {s:02}s'
    if total < 86400:
        return f'{total // 3600}h {m:02}m'
    return f'{d}d {h:02}h'


def ask_yes_no(question: str) -> bool:
    while True:
        ans = input(f'{question} [y/n]').strip().lower()
        if ans in ('y', 'yes', 'true', '1'):
            return True
        if ans in ('n', 'no', 'false', '0'):
            return False


def tuple_product(t) -> Any:
    out = 1
    for v in t:
        out *= v
    return out


def is_pickleable(obj: Any) -> bool:
    import io
    import pickle
    try:
        with io.BytesIO() as stream:
            pickle.dump(obj, stream)
        return True
    except Exception:
        return False


def get_module_from_obj_name(obj_name: str):
    """('pkg.mod', 'attr.sub') split of a dotted name: the longest importable module prefix wins."""
    parts = obj_name.split('.')
    for cut in range(len(parts) - 1, 0, -1):
        mod_name, local = '.'.join(parts[:cut]), '.'.join(parts[cut:])
        try:
            mod = importlib.import_module(mod_name)
            get_obj_from_module(mod, local)
            return mod, local
        except (ImportError, AttributeError):
            continue
    raise ImportError(obj_name)


def get_obj_from_module(module, obj_name: str) -> Any:
    obj = module
    for part in obj_name.split('.') if obj_name else []:
        obj = getattr(obj, part)
    return obj


def get_module_dir_by_obj_name(obj_name: str) -> str:
    import inspect
    module, _ = get_module_from_obj_name(obj_name)
    return _os.path.dirname(inspect.getfile(module))


def is_top_level_function(obj: Any) -> bool:
    return callable(obj) and obj.__name__ in _sys.modules[obj.__module__].__dict__


def get_top_level_function_name(obj: Any) -> str:
    assert is_top_level_function(obj)
    module = obj.__module__
    if module == '__main__':
        module = _os.path.splitext(_os.path.basename(_sys.modules[module].__file__))[0]
    return module + '.' + obj.__name__
