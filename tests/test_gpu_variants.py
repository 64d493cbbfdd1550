"""Every selectable kernel path stays parity-green: single-CTA GEMMs (LRN_GEMM_V1), per-layer pair GEMMs
(LRN_NO_CHAIN), direct-store epilogues (LRN_NO_STAGED), the chain kernel without / with the tensor-memory
conv5 (LRN_CHAIN5), row-major instead of tiled operand rows (LRN_NO_TILED), the decoder context side as K/V GEMMs + SDPA (LRN_CTX_ATTN=0) or stock (LRN_FAST_DECODER=0).
The selection is read once per process, hence subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{}, {"LRN_GEMM_V1": "1"}, {"LRN_NO_CHAIN": "1"}, {"LRN_NO_CHAIN": "1", "LRN_NO_STAGED": "1"},
                                 {"LRN_CHAIN5": "0"}, {"LRN_NO_TILED": "1"}, {"LRN_CTX_ATTN": "0"}, {"LRN_FAST_DECODER": "0"}],
                         ids=["default", "v1", "no_chain", "no_chain_direct", "chain4", "row_major_operands", "kv_sdpa_decoder",
                              "stock_decoder"])
def test_kernel_variant_matches_reference(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_check.py")], env=e, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "variant ok" in r.stdout
