"""Structure of the built library's SASS (no GPU needed: `cuobjdump -sass` on the in-tree .so).

* Blackwell-native: the tensor path is tcgen05 (`UTCHMMA`) fed by TMA (`UTMALDG`), never `HMMA` (mma.sync / wmma).
* tcgen05 / TMA instructions are issued from warp-uniform code: issued inside an `if (lane == 0)` branch, ptxas wraps
  every operand of a `UTCHMMA` / `UTMALDG` into an `ELECT / R2UR.BROADCAST / BRA.U.ANY` loop (~125 clocks of the issuing
  thread per MMA - DESIGN.md section 4, session 4).  The GEMM kernels must hold none of these loops, and the MMA warps of
  the streaming attention kernels must issue their `UTCHMMA`s back to back.
"""
import collections
import re
import shutil
import subprocess

import pytest

from ecad_b200 import build


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    lib = build.build_library()
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    funcs: dict[str, list[str]] = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur.append(m.group(1))
    names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True, check=True).stdout.split("\n")
    return {n: ops for n, ops in zip(names, funcs.values())}


def test_tensor_path_is_tcgen05_only(sass):
    total = collections.Counter(op.split(".")[0] for ops in sass.values() for op in ops)
    assert total["HMMA"] == 0, "legacy mma.sync / wmma instructions in the library"
    assert total["UTCHMMA"] > 1000 and total["UTMALDG"] > 500 and total["LDTM"] > 100


def test_gemm_kernels_have_no_uniform_register_waterfall_loops(sass):
    gemms = {n: ops for n, ops in sass.items() if re.search(r"ecadk::gemm2?_(bf16|splitk)_kernel<", n)}
    assert len(gemms) >= 20
    # the persistent kernels: none at all.  Split-K: its bulk-store epilogue still sits in a lane-0 branch (one loop per
    # store, 7 in the gated-residual instance; measured neutral) - before the conversion these kernels held 47-57 each
    limit = lambda n: 8 if "splitk" in n else 0  # noqa: E731
    bad = {n: ops.count("BRA.U.ANY") for n, ops in gemms.items() if ops.count("BRA.U.ANY") > limit(n)}
    assert not bad, f"ELECT / R2UR.BROADCAST / BRA.U.ANY loops (tcgen05 / TMA issued from divergent code): {bad}"


@pytest.mark.parametrize("pattern,run", [(r"attn_flash_kernel<128, false>", 4), (r"attn_flash_kernel<72, false>", 4),
                                         (r"attn_flash2_kernel<false>", 5)])
def test_streaming_attention_issues_its_mmas_back_to_back(sass, pattern, run):
    """Q K^T of a key block is `run` consecutive UTCHMMAs (at most a UMOV / NOP between two of them)."""
    ops = next(o for n, o in sass.items() if re.search(pattern, n))
    best = cur = 0
    gap = 0
    for op in ops:
        if op.startswith("UTCHMMA"):
            cur += 1
            gap = 0
            best = max(best, cur)
        elif cur and op.split(".")[0] in ("UMOV", "NOP", "UIADD3", "ULOP3") and gap < 2:
            gap += 1
        else:
            cur = 0
    assert best >= run, f"longest UTCHMMA run {best} < {run}: the MMA warp is issuing through R2UR loops again"
