"""Architecture plugins: `extern "C" fl::Module* createModule(nFeature, nLabel)` in a shared object, loaded the way
Train.cpp does (`ModulePlugin(FLAGS_arch).arch(numFeatures, numClasses)`, recipes/slimIPL/src/Train.cpp:390-395)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "wav2letter_b200")
SRC = os.path.join(ROOT, "tests", "plugin", "tds_plugin.cpp")
ARCH = """V -1 NFEAT 1 0
C2 1 4 5 1 2 1 -1 -1
R
LN 3
TDS 4 5 80 0.0
V 0 320 1 0
RO 1 0 3 2
L 320 NLABEL
"""


def build_plugin(tmp_path):
    out = str(tmp_path / "tds_plugin.so")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
           "-L", LIBDIR, "-l:libw2l_b200.so", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def test_plugin_builds_and_exports_create_module(tmp_path):
    """CPU: the plugin compiles against fl_compat.h, links to the library's exported C++ surface and exports createModule"""
    import wav2letter_b200  # noqa: F401  (the library must be built)

    so = build_plugin(tmp_path)
    lib = ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
    assert hasattr(lib, "createModule")


@pytest.mark.gpu
def test_plugin_network_matches_arch_file(tmp_path):
    import torch
    from wav2letter_b200.trainer import Trainer

    so = build_plugin(tmp_path)
    a = Trainer(ARCH, 80, 12, "ctc", "none", lr=0.0)
    b = Trainer(so, 80, 12, "ctc", "none", lr=0.0)  # a path ending in .so goes through ModulePlugin
    assert "SmallTds plugin" in b.describe()
    b.set_flat(a.get_flat(0, 0), 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    feat = torch.randn((2, 1, 80, 48), device="cuda", generator=g)
    tgt = torch.randint(0, 11, (2, 4), device="cuda", generator=g, dtype=torch.int32)
    la, lb = a.step(feat, tgt, True), b.step(feat, tgt, True)
    torch.cuda.synchronize()
    assert torch.equal(la, lb)
    assert torch.equal(a.get_flat(0, 1), b.get_flat(0, 1))  # identical kernels, identical gradients
