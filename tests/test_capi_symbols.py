"""The C-ABI library loads without a GPU and exports every symbol include/lunaris_b200.h declares; the ctypes table in
lunaris_orion_b200/_capi.py agrees with the header (names and argument counts)."""
import os
import re

from lunaris_orion_b200 import _capi

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lunaris_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long)\s+(lun_\w+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_header_declares_the_path():
    d = _declared()
    for name in ("lun_conv_taps_bf16", "lun_wgrad_taps_bf16", "lun_bn_finalize", "lun_affine_fwd_bf16",
                 "lun_attn_ref_rows_bf16", "lun_gn_mish_fwd_bf16", "lun_reparam_fwd"):
        assert name in d


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported by liblunaris_b200.so"


def test_ctypes_table_matches_header():
    d = _declared()
    assert set(d) == set(_capi.SIGNATURES), set(d) ^ set(_capi.SIGNATURES)
    for name, n in d.items():
        assert len(_capi.SIGNATURES[name]) == n, name


def test_version_and_no_compute_without_gpu():
    assert _capi.lib().lun_abi_version() == 1
    assert _capi.lib().lun_launch_count() >= 0
