"""Host side of mate_b200_step_host's compacted device -> host legs (mate_b200/csrc/mate_hostpath.cuh), without a GPU:
tests/native/hostpath_harness.cu builds the tables and streams the two compaction kernels would leave behind on the host,
lets expand_blocks / ExpandPool rebuild (leg 1) or patch (leg 2, MATE_STEP_HOST_ROWS_KEPT) the rows, and compares bytes --
ragged sizes, aligned and unaligned destinations, all-zero / dense / mixed rows.  The device side is covered by
tests/test_cuda_parity.py::test_step_host_*."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_expand_and_patch_on_the_host(tmp_path):
    nvcc = os.environ.get('NVCC') or shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not found')
    exe = str(tmp_path / 'hostpath_harness')
    subprocess.run([nvcc, '-O2', '-std=c++17', '-o', exe, os.path.join(HERE, 'native', 'hostpath_harness.cu'), '-lpthread'],
                   check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stdout + out.stderr
