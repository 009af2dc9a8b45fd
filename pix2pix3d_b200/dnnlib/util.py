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
