"""The C-ABI library loads on a box without a GPU and exports every symbol include/slic_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from video_similarity_search_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "slic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slic_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.load().__dict__.get("_name", _lib.library_path()))
    for n in names:
        assert hasattr(lib, n), "libslic_b200.so does not export %s" % n
    # the Python binding covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.slic_abi_version() == 1
    assert isinstance(lib.slic_last_error(), bytes)


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from video_similarity_search_b200.backend import CudaBackend
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CudaBackend()
    lib = _lib.load()
    assert lib.slic_require_device() == -4          # SLIC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.slic_last_error()


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    assert lib.slic_normalize_rows(None, -1, 0, 0, None, None, None, 0, None) == -1
    assert b"normalize_rows" in lib.slic_last_error()
    assert lib.slic_rows_topk(None, 1, 10, 10, 0, 11, None, None, None) == -1


def test_argument_validation_of_the_multi_gpu_and_widening_entries_needs_no_gpu():
    lib = _lib.load()
    assert lib.slic_nn_top1_sym_part(None, None, 20000, 64, 64, 0, 2, None, 0.0, None, None, None) == -1      # null pointers
    assert b"nn_top1_sym_part" in lib.slic_last_error()
    assert lib.slic_nn_top1_sym_part(None, None, 20000, 64, 60, 0, 2, None, 0.0, None, None, None) == -1      # d_pad % 64
    assert lib.slic_sym_row_bests(None, None, 20000, 64, 64, 2, 2, None, None) == -1                          # part >= parts (after nulls)
    assert lib.slic_unpack_neighbor_keys(None, 0, None, None, None, None) == -1
    assert lib.slic_scatter_last_wins(None, None, -1, 10, -1, None, None, None) == -1
    assert lib.slic_center_columns(None, 0, 0, None, None, None) == -1
    # 70 000 x 70 000 contingency cells exceed the dense limit: refused before any device work
    buf = ctypes.create_string_buffer(64)
    assert lib.slic_cluster_metrics(buf, buf, 70000, 70000, 70000, 1, buf, None) == -3                         # SLIC_ERR_UNSUPPORTED
    assert b"contingency" in lib.slic_last_error()


def test_argument_validation_of_the_peer_window_entries_needs_no_gpu():
    """csrc/comm.cu: bad arguments are refused before any device work (status -1, message naming the entry)."""
    lib = _lib.load()
    assert lib.slic_comm_create(None, 2, 1000, None) == -1 and b"comm_create" in lib.slic_last_error()
    devs = (ctypes.c_int32 * 9)(*range(9))
    out = ctypes.c_void_p()
    assert lib.slic_comm_create(devs, 9, 1000, ctypes.addressof(out)) == -1             # more than 8 devices
    assert lib.slic_comm_create(devs, 2, 0, ctypes.addressof(out)) == -1                 # max_rows
    assert lib.slic_comm_window_create(0, None, None) == -1
    assert lib.slic_comm_connect(None, 0, 2, None) == -1 and b"comm_connect" in lib.slic_last_error()
    assert lib.slic_comm_nn_top1(None, None, None, 20000, 64, 64, None, None, None, None) == -1
    assert lib.slic_comm_finch(None, None, 20000, 64, 1, 32, None, None, None, None, None, None) == -1
    assert lib.slic_finch_multi(None, None, 20000, 64, None, 1, 32, None, None, None, None, None) == -1
    assert b"finch_multi" in lib.slic_last_error()
    assert lib.slic_comm_last_timeline(None, None) == -1
    assert lib.slic_copy_to_device(None, None, 10, None) == -1 and b"copy_to_device" in lib.slic_last_error()
    assert lib.slic_comm_destroy(None) == 0                                              # destroying nothing is fine
    assert lib.slic_set_upload_overlap(0) == 0 and lib.slic_set_upload_overlap(-1) == 0


def test_bench_without_a_gpu_fails_loudly_and_names_the_cpu_arm():
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
