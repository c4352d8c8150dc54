"""ctypes binding of libnemo_b200.so (the C ABI declared in include/nemo_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI.
There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, '..', 'include', 'nemo_b200.h')
LIB_PATH = os.path.join(HERE, 'libnemo_b200.so')

_CTYPES = {'int': ctypes.c_int, 'float': ctypes.c_float, 'size_t': ctypes.c_size_t}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every function the header declares."""
    txt = open(path).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    txt = re.sub(r'//[^\n]*', '', txt)
    protos = {}
    for m in re.finditer(r'(const\s+char\s*\*|int)\s+(nemo_b200_\w+)\s*\(([^)]*)\)\s*;', txt):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if 'char' in ret else ctypes.c_int
        argtypes, argnames = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                nm = re.findall(r'(\w+)$', a)[0]
                ty = a[:-len(nm)].strip()
                if ty.endswith('*'):
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CTYPES[ty.replace('const ', '')])
                argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


# kernels enqueued per C-ABI call when it is not exactly one
_KERNELS_PER_CALL = {'p2p_allreduce': 3, 'embed_fwd': 2, 'embed_bwd': 2, 'monotonic_bwd': 2, 'readout_bwd': 2, 'last_error': 0, 'version': 0, 'sm_arch': 0}


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()
        self.launches = 0   # kernels of this library launched so far (bench.py reports the delta)
        self.profile = None  # set to {} to collect per-entry-point device time (CUDA events; tools/profile_step.py)
        self._pending, self.tag = [], ''

    def collect_profile(self):
        torch.cuda.synchronize()
        for name, e0, e1 in self._pending:
            t, n = self.profile.get(name, (0.0, 0))
            self.profile[name] = (t + e0.elapsed_time(e1), n + 1)
        self._pending = []
        return self.profile

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    'libnemo_b200.so is not built (%s). Run `python -c "import __graft_entry__ as g; g.build()"` '
                    'or `python nemo_cvpr2023_b200/build.py`. There is no CPU or PyTorch fallback.' % LIB_PATH)
            dll = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes, _) in self.protos.items():
                fn = getattr(dll, name)   # AttributeError if the .so lacks a declared symbol
                fn.restype = restype
                fn.argtypes = argtypes
            self._dll = dll
        return self._dll

    def __getattr__(self, short):
        name = 'nemo_b200_' + short
        dll = self.load()
        fn = getattr(dll, name)
        names = self.protos[name][2]
        has_stream = bool(names) and names[-1] == 'stream'

        nk = _KERNELS_PER_CALL.get(short, 1)

        def call(*args):
            self.launches += nk
            a = [_arg(x) for x in args]
            if has_stream:
                a.append(torch.cuda.current_stream().cuda_stream)
            if self.profile is not None and has_stream:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*a)
                e1.record()
                self._pending.append((short + (':' + self.tag if self.tag else ''), e0, e1))
            else:
                rc = fn(*a)
            if rc != 0:
                raise RuntimeError('%s failed (%d): %s' % (name, rc, dll.nemo_b200_last_error().decode()))
        call.__name__ = name
        self.__dict__[short] = call
        return call


def _arg(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError('nemo_b200 kernels take CUDA tensors; got a %s tensor (no CPU fallback)' % x.device)
        if not x.is_contiguous():
            raise RuntimeError('nemo_b200 kernels take contiguous tensors')
        return x.data_ptr()
    return x


lib = _Lib()


class AdamSeg(ctypes.Structure):
    """nemo_b200_adam_seg (include/nemo_b200.h)."""
    _fields_ = [('p', ctypes.c_void_p), ('g', ctypes.c_void_p), ('m', ctypes.c_void_p), ('v', ctypes.c_void_p),
                ('n', ctypes.c_ulonglong), ('lr', ctypes.c_float), ('beta1', ctypes.c_float), ('beta2', ctypes.c_float),
                ('eps', ctypes.c_float), ('weight_decay', ctypes.c_float), ('decoupled', ctypes.c_int),
                ('step', ctypes.c_void_p), ('flags', ctypes.c_int)]


ADAM_SKIP, ADAM_OWNS_STEP, ADAM_MAX_SEGS = 1, 2, 16


def adam_multi(segs, ticket):
    """segs: list of AdamSeg; one launch of nemo_b200_adam_multi on the current stream."""
    arr = (AdamSeg * len(segs))(*segs)
    lib.adam_multi(ctypes.addressof(arr), len(segs), ticket)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('nemo_cvpr2023_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
