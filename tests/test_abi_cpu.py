"""CPU: the C-ABI shared library loads and exports exactly the symbols include/vaegan_b200.h declares, the ctypes
prototypes cover all of them, and the host-side module mirror matches the oracle's state_dict contract.
No compute entry point is called here (there is no GPU)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vaegan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vaegan_b200.h but not exported by the library"


def test_ctypes_prototypes_cover_header():
    import vaegan_b200
    from importlib import import_module
    protos = import_module("vaegan_b200._lib").PROTOTYPES
    assert sorted(protos) == _declared_symbols()


def test_no_gpu_means_loud_failure(lib):
    import vaegan_b200 as vb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.vg_version() == 100
    assert lib.vg_device_check() != 0          # no device -> error code, not a silent fallback
    assert lib.vg_last_error()
    enc = vb.Encoder([3, 64, 64], 128)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(2, 3, 64, 64))


@pytest.mark.parametrize("hw,nz", [(64, 128), (128, 256), (256, 100)])
def test_state_dict_contract(hw, nz):
    import vaegan_b200 as vb
    from oracle import vaegan_oracle as vo
    torch.manual_seed(0)
    mine = (vb.Encoder([3, hw, hw], nz), vb.Generator(nz=nz, hw=hw), vb.Discriminator(hw=hw))
    ref = vo.build_nets(vo.NetConfig(hw=hw, nz=nz))
    for a, b in zip(mine, ref):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            assert sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype, k
        a.load_state_dict(sb)      # round trip
    # encoder ctor side effect (main_vae.py:43-45): BN buffers after the train-mode dry run
    e = mine[0]
    assert int(e.cnn[0].bn.num_batches_tracked) == 1
    assert torch.allclose(e.cnn[0].bn.running_var, torch.full_like(e.cnn[0].bn.running_var, 0.9))


def test_weights_init_matches_reference_semantics():
    import vaegan_b200 as vb
    from oracle import vaegan_oracle as vo
    torch.manual_seed(7)
    g1 = vb.Generator(nz=16, hw=8)
    g1.apply(vb.weights_init)
    torch.manual_seed(7)
    g2 = vo.make_generator(nz=16, hw=8)
    g2.apply(vo.weights_init)
    for (k, a), (_, b) in zip(g1.state_dict().items(), g2.state_dict().items()):
        assert torch.equal(a, b), k


def test_discriminator_rejects_small_images_like_reference():
    """the native (hw=256) Discriminator cannot take 64x64 input."""
    import vaegan_b200 as vb
    d = vb.Discriminator()
    spec_chain = d._layers()
    h = w = 64
    with pytest.raises(RuntimeError, match="Kernel size can't be greater than actual input size"):
        for layer in spec_chain:
            h, w = layer.spec.out_hw(h, w)


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/vaegan_b200.h compiles as C (no C++ / torch types at the boundary) and the POD structs have the size
    and field offsets the ctypes mirror assumes."""
    import ctypes
    import shutil
    import subprocess
    from importlib import import_module
    import vaegan_b200  # noqa: F401
    _lib = import_module("vaegan_b200._lib")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vaegan_b200.h"\n'
                   'int main(void) {\n'
                   '  printf("%zu %zu %zu\\n", sizeof(VgConvGeom), sizeof(VgPackItem), sizeof(VgEpilogue));\n'
                   '  printf("%zu %zu %zu %zu\\n", offsetof(VgConvGeom, big_c_valid), offsetof(VgPackItem, small_c),\n'
                   '         offsetof(VgEpilogue, sums), offsetof(VgEpilogue, stats));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes, offs = [int(v) for v in out[:3]], [int(v) for v in out[3:]]
    assert sizes == [ctypes.sizeof(_lib.VgConvGeom), ctypes.sizeof(_lib.VgPackItem), ctypes.sizeof(_lib.VgEpilogue)]
    assert offs == [_lib.VgConvGeom.big_c_valid.offset, _lib.VgPackItem.small_c.offset, _lib.VgEpilogue.sums.offset,
                    _lib.VgEpilogue.stats.offset]
