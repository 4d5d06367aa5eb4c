"""CPU tests: the C-ABI library loads and exports every symbol include/pfcu.h declares; no compute calls."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pfcu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfcu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import pfcu

    lib = pfcu.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(pfcu.EXPORTS) == names
    assert lib.pfcu_abi_version() == 1


def test_host_streamer_library_loads_and_only_depends_on_the_c_abi():
    """lib/libpfhost.so (host/frame_streamer.cpp, the application-side C++ submit loop): loads next to libpfcu.so, exports
    what its header declares, and every symbol it leaves undefined is one include/pfcu.h declares."""
    import subprocess

    import pfcu

    assert hasattr(pfcu.host_lib(), "pfhost_stream_frames")
    hdr = open(os.path.join(ROOT, "pathfinder-cpp_b200", "host", "frame_streamer.h")).read()
    assert re.findall(r"\b(pfhost_[a-z_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)) == ["pfhost_stream_frames"]
    nm = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(ROOT, "pathfinder-cpp_b200", "lib", "libpfhost.so")],
                        capture_output=True, text=True, check=True).stdout
    used = sorted(set(re.findall(r"\b(pfcu_[a-z0-9_]+)", nm)))
    assert used and set(used) <= set(declared_symbols()), used


def test_no_cpu_fallback_without_device():
    """Without a GPU the product path must fail loudly (PFCU_ERR_CUDA), never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import pfcu

    with pytest.raises(pfcu.PfcuError, match="no CUDA device"):
        pfcu.Renderer(0)


def test_product_does_not_touch_oracle():
    """Nothing under pathfinder-cpp_b200/ or include/ may import, link or reference oracle/."""
    for base in ("pathfinder-cpp_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in dirpath or os.sep + "lib" in dirpath:
                continue
            for f in files:
                if f.endswith((".cu", ".h", ".cpp", ".py", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    for word in ("pforacle", "pfref", "libpforacle", "pfshader", "glsl_shim"):
                        assert word not in text, (f, word)
