"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/dd_kernels.h declares (no compute is attempted without a GPU)."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope='module')
def lib_path():
  from daydreamer_b200 import build
  return build.build()


def declared_symbols():
  text = (ROOT / 'include' / 'dd_kernels.h').read_text()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(dd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib_path):
  lib = ctypes.CDLL(str(lib_path))
  names = declared_symbols()
  assert len(names) >= 40
  for name in names:
    assert hasattr(lib, name), f'{name} declared in dd_kernels.h but not exported'


def test_binding_covers_header(lib_path):
  from daydreamer_b200 import backend
  lib = backend.load_library(lib_path)
  typed = set(backend.SIGNATURES) | {'dd_version', 'dd_last_error', 'dd_launch_count',
                                     'dd_observe_fwd', 'dd_observe_fwd_grid',
                                     'dd_observe_bwd', 'dd_observe_bwd_grid', 'dd_program_bytes',
                                     'dd_program_build', 'dd_program_deps', 'dd_program_debug'}
  assert typed == set(declared_symbols())
  assert lib.dd_version() >= 100
  assert lib.dd_launch_count() == 0


def test_library_is_sm100a(lib_path):
  import shutil
  import subprocess
  cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
  out = subprocess.run([cuobjdump, '-lelf', str(lib_path)], capture_output=True, text=True).stdout
  assert 'sm_100a' in out, out


def test_product_fails_loudly_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  from daydreamer_b200.backend import CudaBackend
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    CudaBackend()


def test_product_never_imports_oracle_or_hostsim():
  for f in (ROOT / 'daydreamer_b200').rglob('*.py'):
    src = f.read_text()
    assert not re.search(r'^\s*(from|import)\s+(oracle|tests)\b', src, flags=re.M), f
