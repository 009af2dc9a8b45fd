"""The C-ABI library loads and exports every symbol include/p3d.h declares (no compute: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'p3d.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(p3d_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ('p3d_render_fwd', 'p3d_bias_act', 'p3d_upfirdn2d', 'p3d_ray_sampler', 'p3d_run_model'):
        assert must in syms


def test_library_exports_all_declared_symbols():
    from pix2pix3d_b200 import _lib
    if not _lib.available():
        pytest.skip('libp3d.so not built (run python -m pix2pix3d_b200.build)')
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, missing
    assert handle.p3d_abi_version() == _lib.ABI_VERSION


def test_ctypes_table_covers_header():
    from pix2pix3d_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_struct_layouts_match_header_sizes():
    """sizeof of the ctypes mirrors equals what a C compiler computes for include/p3d.h."""
    import subprocess
    import tempfile
    from pix2pix3d_b200 import _lib
    prog = '#include <stdio.h>\n#include "p3d.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(p3d_decoder_t), sizeof(p3d_render_args_t), sizeof(p3d_conv_args_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 't.c')
        open(c, 'w').write(prog)
        exe = os.path.join(td, 't')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        a, b, c = (int(v) for v in subprocess.check_output([exe]).split())
    from pix2pix3d_b200 import tcconv
    assert ctypes.sizeof(_lib.DecoderDesc) == a
    assert ctypes.sizeof(_lib.RenderArgs) == b
    assert ctypes.sizeof(tcconv.ConvArgs) == c


def test_cuda_ops_fail_loudly_without_library(monkeypatch):
    from pix2pix3d_b200 import _lib
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libp3d.so')
    monkeypatch.setattr(_lib, '_lib', None)
    with pytest.raises(RuntimeError, match='missing'):
        _lib.lib()


def test_library_exports_nothing_undeclared():
    """Every `p3d_*` symbol the shared library exports is declared in include/p3d.h (no entry points outside the documented ABI)."""
    import shutil
    import subprocess
    from pix2pix3d_b200 import _lib
    if not _lib.available() or not shutil.which('nm'):
        pytest.skip('needs the built library and nm')
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH], text=True)
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if ln.split() and re.fullmatch(r'p3d_[a-z0-9_]+', ln.split()[-1])})
    assert exported == declared_symbols()
