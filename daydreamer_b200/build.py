"""Builds the sm_100a kernel library in-tree (daydreamer_b200/lib/libdd_kernels.so).

nvcc cross-compiles without a GPU, so this runs in the CPU-only container and
the resulting .so travels to the GPU box with the repo snapshot."""
import hashlib
import os
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / 'csrc'
LIBDIR = ROOT / 'lib'
LIB = LIBDIR / 'libdd_kernels.so'
SOURCES = ['dd_gemm.cu', 'dd_gemm_tc.cu', 'dd_gemm_skinny.cu', 'dd_gemm_thin.cu', 'dd_rowops.cu', 'dd_scan.cu', 'dd_elem.cu', 'dd_conv.cu', 'dd_ctrl.cu', 'dd_rng.cu', 'dd_machine.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
]


def _digest():
  h = hashlib.sha256()
  files = sorted(CSRC.glob('*.cu')) + sorted(CSRC.glob('*.cuh'))
  files.append(ROOT.parent / 'include' / 'dd_kernels.h')
  for f in files:
    h.update(f.name.encode())
    h.update(f.read_bytes())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force=False, verbose=False):
  """Compile if sources changed.  Returns the library path."""
  LIBDIR.mkdir(exist_ok=True)
  stamp = LIBDIR / 'libdd_kernels.sha256'
  digest = _digest()
  if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
    return LIB
  nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
  cmd = [nvcc, *NVCC_FLAGS, *[str(CSRC / s) for s in SOURCES], '-o', str(LIB)]
  if verbose:
    cmd.insert(1, '-Xptxas=-v')
    print(' '.join(cmd))
  subprocess.run(cmd, check=True)
  stamp.write_text(digest)
  return LIB


if __name__ == '__main__':
  import sys
  print(build(force=True, verbose='-v' in sys.argv))
