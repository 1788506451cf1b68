"""The C-ABI library loads and exports every symbol include/bc_engine.h declares (no GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bc_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(bc):
    lib = bc.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libbc_engine.so does not export {n}"
    from blume_capel_b200 import engine
    assert sorted(engine.SYMBOLS) == names, "python binding and header disagree"


def test_abi_version(bc):
    assert bc.load_library().bc_abi_version() == 2


def test_no_torch_types_or_cpp_in_header():
    src = open(os.path.join(ROOT, "include", "bc_engine.h")).read()
    assert "torch" not in src and "std::" not in src and "at::" not in src
    assert 'extern "C"' in src


def test_argument_errors_need_no_gpu(bc):
    lib = bc.load_library()
    h = C.c_void_p()
    assert lib.bc_create(C.byref(h), 33, 0, 1, 0, 0, 0) == -1  # odd torus -> BC_ERR_ARG
    assert b"even" in lib.bc_last_error(None)
    assert lib.bc_create(C.byref(h), 32, 5, 1, 0, 0, 0) == -1
    assert lib.bc_create(C.byref(h), 32, 0, 0, 0, 0, 0) == -1
    assert lib.bc_create(C.byref(h), 32, 0, 1, 9, 0, 0) == -5  # unsupported storage
    assert lib.bc_create(None, 32, 0, 1, 0, 0, 0) == -1


def test_new_entry_points_check_their_arguments(bc):
    """Round-2 entry points reject NULL handles / buffers before touching a device."""
    lib = bc.load_library()
    buf = C.create_string_buffer(256)
    n = C.c_int64()
    assert lib.bc_gap_histogram(None, 0, 0, 0, buf) == -1
    assert lib.bc_cluster_lists(None, 0, 0, 0, buf, buf, buf, C.byref(n), C.byref(n)) == -1
    assert lib.bc_upload_packed(None, 0, buf, 0) == -1 and lib.bc_download_packed(None, 0, buf, 0) == -1
    assert lib.bc_packed_size(None, 1) == -1
    assert lib.bc_slab_cluster_export(None, buf) == -1 and lib.bc_slab_cluster_connect(None, buf) == -1
    assert lib.bc_pack_spins(None, 4, 1, buf) == -1 and lib.bc_unpack_spins(buf, 0, 1, buf) == -1


def test_fails_loudly_without_a_gpu(bc):
    """No CPU fallback: on a box without CUDA, bc_create must fail with BC_ERR_CUDA."""
    lib = bc.load_library()
    h = C.c_void_p()
    rc = lib.bc_create(C.byref(h), 32, 0, 1, 0, 0, 0)
    if rc == 0:
        lib.bc_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == -2
    assert b"no CPU fallback" in lib.bc_last_error(None)
    with pytest.raises(bc.BCError):
        bc.Engine(32)


def test_product_does_not_touch_the_oracle():
    """The shipped package and library must not import, link or mention oracle/."""
    pkg = os.path.join(ROOT, "blume_capel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "bc_oracle" not in text and "oracle." not in text and "from oracle" not in text, f


def test_host_campaign_driver_arguments_and_no_fallback(tmp_path, bc):
    """bc_runner (the runner.sh campaign in one process): argument errors need no GPU, and without a GPU it fails
    loudly instead of computing anything on the CPU."""
    import subprocess
    exe = os.path.join(ROOT, "blume_capel_b200", "host", "bc_runner")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: bc_runner root T D J" in r.stderr
    r = subprocess.run([exe, "d", "0.6", "1.9", "1.0", "--L", "64,128", "--sweeps-per-sample", "4", "--lockstep"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "lockstep" in r.stderr
    r = subprocess.run([exe, "d", "0.6", "1.9", "1.0", "--L", "8,12", "--runs", "1,2,3"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 2 and "--runs" in r.stderr
    lib = bc.load_library()
    h = C.c_void_p()
    if lib.bc_create(C.byref(h), 32, 0, 1, 0, 0, 0) == 0:
        lib.bc_destroy(h)
        return  # a CUDA device is present: the GPU tests cover the rest
    r = subprocess.run([exe, "d", "0.6", "1.9", "1.0", "--L", "8", "--runs", "2", "--nsamples", "1", "--mkdir"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "d" / "spin" / "8" / "0" / "0.txt")


def test_header_is_plain_c_and_links(tmp_path):
    """include/bc_engine.h compiles as C99 (what a cgo / JNI / ctypes-free binding would include) and every entry
    point it declares resolves against libbc_engine.so at link time."""
    import subprocess
    names = declared_symbols()
    src = tmp_path / "abi.c"
    body = "\n".join(f"    p[{i}] = (void*)&{n};" for i, n in enumerate(names))
    src.write_text('#include "bc_engine.h"\n#include <stdio.h>\nint main(void) {\n'
                   f"    void* p[{len(names)}];\n{body}\n"
                   f"    printf(\"%d %p\\n\", bc_abi_version(), p[{len(names) - 1}]);\n    return 0;\n}}\n")
    lib_dir = os.path.join(ROOT, "blume_capel_b200")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"),
                    "-o", str(exe), str(src), "-L", lib_dir, "-lbc_engine", f"-Wl,-rpath,{lib_dir}"], check=True,
                   capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0] == "2"


def test_wire_format_host_conversion_needs_no_gpu(bc):
    """bc_pack_spins / bc_unpack_spins: four sites per byte, code spin + 1 in bits 2i..2i+1, rows padded to whole bytes."""
    import numpy as np
    rng = np.random.default_rng(3)
    for L in (1, 3, 4, 6, 64, 101):
        s = rng.integers(-1, 2, size=(5, L), dtype=np.int8)
        p = bc.pack_spins(s)
        assert p.size == 5 * ((L + 3) // 4)
        assert np.array_equal(bc.unpack_spins(p, L), s)
    assert bc.pack_spins(np.array([[-1, 0, 1, 1, 0]], np.int8)).tolist() == [0b10100100, 0b01010101]
