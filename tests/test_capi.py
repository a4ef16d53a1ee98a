"""CPU tier: the C-ABI library loads, exports every symbol include/mtn_b200.h declares, and the
ctypes structure layouts match the C ones (checked by compiling the header with gcc).
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "mtn_b200.h")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mtn_b200 import _lib
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "libmtn_b200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding and header disagree"
    assert L.mtn_abi_version() == _lib.ABI_VERSION == 7
    assert L.mtn_mask_words(1) == 4 and L.mtn_mask_words(128) == 4 and L.mtn_mask_words(129) == 8
    assert L.mtn_attn_site_workspace_bytes(32, 256, 512, 512) > 32 * 256 * 512 * 2 * 5
    assert L.mtn_ffn_workspace_bytes(100, 512, 2048) >= 100 * (512 + 2048) * 2


def test_header_is_plain_c_and_struct_layouts_match(tmp_path):
    from mtn_b200 import _lib
    structs = {"MtnLinearArgs": _lib.LinearArgs, "MtnAttnCoreArgs": _lib.AttnCoreArgs,
               "MtnAttnSiteArgs": _lib.AttnSiteArgs, "MtnFfnArgs": _lib.FfnArgs,
               "MtnGemmArgs": _lib.GemmArgs, "MtnLinearDgradArgs": _lib.LinearDgradArgs,
               "MtnLinearWgradArgs": _lib.LinearWgradArgs, "MtnLayerNormBwdArgs": _lib.LayerNormBwdArgs,
               "MtnEmbedBwdArgs": _lib.EmbedBwdArgs, "MtnAttnCoreBwdArgs": _lib.AttnCoreBwdArgs,
               "MtnDecodeSite": _lib.DecodeSite, "MtnDecodeClusterArgs": _lib.DecodeClusterArgs,
               "MtnAttnSiteFusedArgs": _lib.AttnSiteFusedArgs}
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "mtn_b200.h"', 'int main(void){']
    for cname, cls in structs.items():
        prog.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            prog.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    prog.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(c), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so error paths are testable on the CPU box."""
    from mtn_b200 import _lib
    L = _lib.lib()
    a = _lib.LinearArgs()
    assert L.mtn_linear_fwd(ctypes.byref(a), None) == -5          # MTN_E_ARG: NULL operands
    assert b"NULL" in L.mtn_last_error()
    a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = 16, 16, 8, 64, 12, 12, 12
    assert L.mtn_linear_fwd(ctypes.byref(a), None) == -1          # MTN_E_SHAPE: K % 8
    assert b"multiple of 8" in L.mtn_last_error()
    c = _lib.AttnCoreArgs()
    c.q = c.k = c.v = c.out = 16
    c.B, c.h, c.Lq, c.Lk, c.d_k = 1, 1, 4, 4, 48
    assert L.mtn_attn_core_fwd(ctypes.byref(c), None) == -1
    assert b"d_k=48" in L.mtn_last_error()
    assert L.mtn_layernorm_fwd(None, None, None, 1e-6, 1, 4, None, None, None) == -5
    # the cluster decoding step (ABI v7): shape support and argument validation happen on the host
    assert L.mtn_decode_cluster_supported(64, 512, 8, 2048) == 1 and L.mtn_decode_cluster_supported(128, 512, 8, 2048) == 1
    assert L.mtn_decode_cluster_supported(129, 512, 8, 2048) == 0 and L.mtn_decode_cluster_supported(64, 1024, 16, 4096) == 0
    assert L.mtn_decode_cluster_max_sites() >= 42                  # N = 6 layers x (self + 5 cross + feed-forward)
    d = _lib.DecodeClusterArgs()
    assert L.mtn_decode_cluster_fwd(ctypes.byref(d), None) == -5  # NULL site list / rows / norm
    assert b"NULL" in L.mtn_last_error()
    site = (_lib.DecodeSite * 1)()
    d.sites, d.n_sites, d.x_in, d.out, d.norm_a, d.norm_b = site, 1, 16, 16, 16, 16
    d.B, d.d, d.h, d.d_ff = 64, 256, 4, 1024
    assert L.mtn_decode_cluster_fwd(ctypes.byref(d), None) == -1  # MTN_E_SHAPE
    assert b"d = 512" in L.mtn_last_error()
    d.d, d.h, d.d_ff, d.rows_per_dialogue = 512, 8, 2048, 5
    assert L.mtn_decode_cluster_fwd(ctypes.byref(d), None) == -1 and b"rows_per_dialogue" in L.mtn_last_error()
    d.rows_per_dialogue = 1
    assert L.mtn_decode_cluster_fwd(ctypes.byref(d), None) == -5 and b"site 0" in L.mtn_last_error()   # site with NULL pointers


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mtn_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MtnError, match="not found"):
        _lib.lib()
