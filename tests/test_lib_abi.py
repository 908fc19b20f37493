"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/ecad_b200.h declares.
No compute calls here (no GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from ecad_b200.build import build_library
    from ecad_b200 import _lib

    build_library()
    return _lib.load()


def _declared_symbols():
    text = (ROOT / "include" / "ecad_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecadk_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from ecad_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared  # the ctypes table and the header agree


def test_abi_version_and_error_channel(lib):
    assert lib.ecadk_abi_version() == 2  # 2: EcadkBlocksArgs.self_bias (padded token counts)
    assert isinstance(lib.ecadk_last_error(), bytes)


def test_struct_layouts_match_header(lib):
    """sizeof of the ctypes mirrors == what a C compiler computes for the header's structs."""
    import subprocess
    import tempfile

    from ecad_b200 import _lib

    src = r'''
#include <stdio.h>
#include "ecad_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(EcadkReuse), sizeof(EcadkResidualLnArgs),
         sizeof(EcadkModelDesc), sizeof(EcadkBlockWeights), sizeof(EcadkBlocksArgs), sizeof(EcadkFluxDesc),
         sizeof(EcadkFluxDoubleWeights), sizeof(EcadkFluxSingleWeights), sizeof(EcadkFluxArgs),
         sizeof(EcadkProfileRecord));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "t.c"
        c.write_text(src)
        exe = Path(d) / "t"
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], check=True)
        sizes = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    mine = [ctypes.sizeof(t) for t in (_lib.EcadkReuse, _lib.EcadkResidualLnArgs, _lib.EcadkModelDesc,
                                        _lib.EcadkBlockWeights, _lib.EcadkBlocksArgs, _lib.EcadkFluxDesc,
                                        _lib.EcadkFluxDoubleWeights, _lib.EcadkFluxSingleWeights, _lib.EcadkFluxArgs,
                                        _lib.EcadkProfileRecord)]
    assert mine == sizes


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of computing on the CPU, and never imports oracle/."""
    import torch

    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler

    if torch.cuda.is_available():
        pytest.skip("CPU-box check")
    with pytest.raises(RuntimeError, match="no CPU path"):
        B200PixArtTransformer2D({}, dit_scheduler=SequentialDiTScheduler())
    for f in (ROOT / "ecad_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("oracle/", "").replace("the oracle", ""), f
