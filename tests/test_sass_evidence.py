"""What the shipped library is made of, read from its SASS on CPU (`cuobjdump -sass` of libpfn_b200.so): the hot
kernels must contain the Blackwell-native instructions the design claims (B200_PROFILING.md's mnemonics) -- UTCHMMA
(tcgen05.mma kind::tf32), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor loads), UTCBAR (tcgen05.commit), UBLKCP
(cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier) -- and no legacy warp-level MMA (HMMA).  The committed tables
profiles/r2_sass_*.txt are the same counts per object file (scripts/sass_table.py)."""
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UBLKCP", "LDGSTS", "SYNCS", "HMMA"]


@pytest.fixture(scope="module")
def kernels():
    """{demangled-ish kernel name: {mnemonic: count}} for every sm_100a function of the library."""
    from poweflownet_b200.build import build
    lib = build()
    try:
        out = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, timeout=900)
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert out.returncode == 0, out.stderr[-1000:]
    assert "sm_100a" in out.stdout  # compiled for the arch-specific target, not a generic sm_100 / PTX-only build
    table, cur = {}, None
    for line in out.stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = table.setdefault(m.group(1), {k: 0 for k in MNEMONICS})
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for k in MNEMONICS:
                if op.startswith(k):
                    cur[k] += 1
    assert table
    return table


def _sum(kernels, name_part, mnemonic):
    hits = {k: v for k, v in kernels.items() if name_part in k}
    assert hits, f"no kernel named *{name_part}* in the library"
    return sum(v[mnemonic] for v in hits.values())


def test_graph_resident_kernels_run_on_tcgen05_with_tmem_and_tma(kernels):
    for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS"):
        assert _sum(kernels, "k_mpn_fused_fwd", m) > 0, m


def test_layerwise_gemm_and_weight_gradients_run_on_tcgen05(kernels):
    for name in ("k_gemm_tc", "k_wgrad_group", "k_wgrad_tc"):
        for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"):
            assert _sum(kernels, name, m) > 0, (name, m)
    # operands written into tensor memory by the converter warps (A of k_gemm_tc, dY^T of k_wgrad_group)
    assert _sum(kernels, "k_gemm_tc", "STTM") > 0
    assert _sum(kernels, "k_wgrad_group", "STTM") > 0


def test_edge_aggregation_forward_stages_through_bulk_copies_and_mbarriers(kernels):
    for m in ("UBLKCP", "LDGSTS", "SYNCS"):
        assert _sum(kernels, "k_ea_fwd_tma", m) > 0, m


def test_no_legacy_warp_level_mma_anywhere(kernels):
    assert sum(v["HMMA"] for v in kernels.values()) == 0
