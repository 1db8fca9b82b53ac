"""NVRTC compilation of generated / user CUDA source, with an in-memory and an
on-disk cubin cache, and module / function handles.

Mirrors cupy/cuda/compiler.py (`_compile_with_cache_cuda` :655-790: option list
incl. `-ftz=true` :667, SHA-1 key over arch/options/NVRTC version/source
:690-708, disk cache cupy/cuda/_compiler_cache.py:78-182) and
cupy/cuda/function.pyx:193-232 (Module.load / get_function).  The compiler and
loader themselves are the C-ABI calls b200_jit_compile / b200_module_load.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import threading

from cupy_b200 import _lib
from cupy_b200._core import _dryrun

_ARCH = 'sm_100a'
_INCLUDE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'csrc', 'include')
_CUDA_INCLUDE = os.environ.get('CUDA_HOME', os.environ.get('CUDA_PATH', '/usr/local/cuda')) + '/include'

_kernel_name_re = re.compile(r'^[a-zA-Z_][a-zA-Z_0-9]*$')


def is_valid_kernel_name(name):
    return _kernel_name_re.match(name) is not None


def default_options():
    return ('--gpu-architecture=' + _ARCH, '--std=c++17', '-ftz=true', '-lineinfo',
            '-I' + _INCLUDE_DIR, '-I' + _CUDA_INCLUDE, '-default-device',
            '-diag-suppress=177', '-diag-suppress=550')


def cache_dir():
    d = os.environ.get('CUPY_B200_CACHE_DIR', os.path.expanduser('~/.cupy_b200/kernel_cache'))
    return d


_header_digest = None


def _headers_digest():
    """Digest of the skeleton headers: a header edit invalidates cached cubins."""
    global _header_digest
    if _header_digest is None:
        h = hashlib.sha1()
        root = os.path.join(_INCLUDE_DIR, 'b200')
        for fn in sorted(os.listdir(root)):
            with open(os.path.join(root, fn), 'rb') as f:
                h.update(fn.encode())
                h.update(f.read())
        _header_digest = h.hexdigest()
    return _header_digest


def compile_to_cubin(source, options=(), name='kernel.cu'):
    """source -> cubin bytes (no GPU needed: NVRTC cross-compiles for sm_100a)."""
    opts = tuple(default_options()) + tuple(options)
    # the checkout's own include directory is named by role, not by path (its CONTENT is in the header digest),
    # so a cache survives moving or re-cloning the tree
    key_opts = tuple('-I<b200/include>' if o == '-I' + _INCLUDE_DIR else o for o in opts)
    key = hashlib.sha1(('\0'.join(key_opts) + '\0' + _headers_digest() + '\0' + source).encode()).hexdigest()
    use_disk = os.environ.get('CUPY_B200_CACHE_IN_MEMORY', '0') != '1'
    path = os.path.join(cache_dir(), key + '.cubin')
    if use_disk and os.path.exists(path):
        with open(path, 'rb') as f:
            return f.read()
    c_opts = (ctypes.c_char_p * len(opts))(*[o.encode() for o in opts])
    image = ctypes.c_void_p()
    size = ctypes.c_size_t()
    st = _lib.lib.b200_jit_compile(source.encode(), name.encode(), len(opts), c_opts,
                                   ctypes.byref(image), ctypes.byref(size))
    if st != 0:
        log = (_lib.lib.b200_jit_last_log() or b'').decode('utf-8', 'replace')
        if st == _lib.E_COMPILE:
            if os.environ.get('CUPY_B200_DUMP_CUDA_SOURCE_ON_ERROR', '0') == '1':
                log += '\n---- source ----\n' + '\n'.join(
                    '%4d  %s' % (i + 1, line) for i, line in enumerate(source.split('\n')))
            raise _lib.CompileException(st, _lib.last_error(), log, source)
        _lib.check(st)
    try:
        cubin = ctypes.string_at(image.value, size.value)
    finally:
        _lib.lib.b200_jit_free_image(image)
    if use_disk:
        try:
            os.makedirs(cache_dir(), exist_ok=True)
            tmp = path + '.%d.tmp' % os.getpid()
            with open(tmp, 'wb') as f:
                f.write(cubin)
            os.replace(tmp, path)
        except OSError:
            pass
    return cubin


class Module:
    """A loaded cubin (lives for the life of the process, like the reference's memoized modules)."""

    def __init__(self, cubin):
        self._cubin = cubin          # keep the image alive
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.b200_module_load(cubin, ctypes.byref(h)))
        self._handle = h

    def get_function(self, name):
        f = ctypes.c_void_p()
        _lib.check(_lib.lib.b200_module_get_function(self._handle, name.encode(), ctypes.byref(f)))
        return Function(self, name, f)


class Function:
    __slots__ = ('module', 'name', 'handle')

    def __init__(self, module, name, handle):
        self.module, self.name, self.handle = module, name, handle


_function_memo = {}
_memo_lock = threading.Lock()


def get_function(source, name, options=()):
    """Compile (cached) + load on the current device + look up `name`."""
    if _dryrun.enabled:
        compile_to_cubin(source, options, name + '.cu')     # must compile for sm_100a
        return Function(None, name, None)
    import torch
    dev = torch.cuda.current_device()
    key = (dev, name, options, source)
    fn = _function_memo.get(key)
    if fn is None:
        cubin = compile_to_cubin(source, options, name + '.cu')
        fn = Module(cubin).get_function(name)
        with _memo_lock:
            fn = _function_memo.setdefault(key, fn)
    return fn
