"""CPU: the C-ABI library loads, exports every symbol the headers declare, and fails loudly without a GPU."""
import os
import re

import pytest

import helix_b200 as hb
from helix_b200 import _lib, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("helix_b200.h", "helix_b200_kernels.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(hbk?_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    l = hb.load_library()
    decl = declared_symbols()
    assert decl, "no declarations parsed"
    missing = [n for n in sorted(decl) if not hasattr(l, n)]
    assert not missing, missing
    assert decl == set(_lib.SIGNATURES), (decl ^ set(_lib.SIGNATURES))
    assert l.hb_abi_version() == 2


def test_struct_sizes_match_header_layout():
    import ctypes as C
    assert C.sizeof(_lib.EngineCfg) == 64 and C.sizeof(_lib.ModelDescC) == 100
    assert C.sizeof(_lib.SamplingC) == 64


def test_ctypes_layouts_match_the_c_header_field_by_field(tmp_path):
    """Compile the public header with gcc and compare sizeof / offsetof of every struct field with the ctypes mirror:
    a cgo / ctypes binding is only as good as its struct layouts."""
    import ctypes as C
    import subprocess
    structs = {"hb_engine_cfg": _lib.EngineCfg, "hb_model_desc": _lib.ModelDescC, "hb_sampling": _lib.SamplingC,
               "hb_stats": _lib.StatsC}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "helix_b200.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), (cname, got[cname], C.sizeof(ct))
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_memory_estimate_matches_survey_numbers():
    est = hb.engine.memory_estimate(configs.llama3_8b(), hb.EngineConfig(max_seqs=32, max_ctx=2048))
    assert abs(est["weights"] - 16.06e9) < 0.02e9            # 16.06 GB bf16
    assert est["kv"] == 32 * 2048 * 131072                   # 131072 B/token
    l1 = hb.engine.memory_estimate(configs.llama32_1b(), hb.EngineConfig(max_seqs=1, max_ctx=64))
    assert abs(l1["weights"] - 2.47e9) < 0.01e9 and l1["kv"] == 64 * 32768
    with pytest.raises(hb.HBError):
        bad = configs.llama3_8b()
        bad.head_dim = 96
        hb.engine.memory_estimate(bad, hb.EngineConfig())


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hb.HBError) as ei:
        hb.Engine(hb.EngineConfig())
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "helix_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
