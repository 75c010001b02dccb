"""The C-ABI library loads without a GPU and exports every symbol include/vtb200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vtb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "vtb_gemm_bf16" in syms and "vtb_attention_bwd" in syms and len(syms) >= 18


def test_library_exports_every_declared_symbol():
    from vtb200 import lib

    handle = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"libvtb200.so lacks {missing}"


def test_python_binding_lists_the_same_symbols():
    from vtb200 import lib

    assert sorted(lib.SYMBOLS) == declared_symbols()
    lib.load()  # sets prototypes for all of them
    assert lib.load().vtb_version() >= 100


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the parameter structs: field counts follow the header (guards silent ABI drift)."""
    from vtb200 import lib

    text = open(os.path.join(ROOT, "include", "vtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)

    def fields(struct_name):
        body = re.search(r"typedef struct \{([^{}]*)\} " + struct_name + ";", text, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = [p.strip() for p in decl.split(",")]
            names.append(re.findall(r"([A-Za-z_0-9]+)$", parts[0])[0])
            for extra in parts[1:]:
                names.append(re.findall(r"([A-Za-z_0-9]+)$", extra)[0])
        return names

    assert fields("vtb_gemm_params") == [f[0] for f in lib.GemmParams._fields_]
    assert fields("vtb_attn_params") == [f[0] for f in lib.AttnParams._fields_]


def test_ops_refuse_to_run_without_cuda():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vtb200 import lib

    with pytest.raises(RuntimeError, match="CUDA device required"):
        lib.get()


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/vtb200.h compiles as strict C99 (no C++-isms, no torch / CUDA types) and a plain C
    program links against libvtb200.so and calls through it (no GPU needed for vtb_version / vtb_last_error)."""
    import subprocess

    from vtb200 import lib

    src = tmp_path / "abi_probe.c"
    src.write_text(
        '#include "vtb200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  vtb_gemm_params g; vtb_attn_params a; (void)g; (void)a;\n"
        "  /* a bad call must fail through the error channel, not crash: NULL source pointers */\n"
        "  float m[3] = {0, 0, 0}, s[3] = {1, 1, 1};\n"
        "  int rc = vtb_input_batch(0, 1, 0, 1, 8, 8, m, s, 0, 0);\n"
        '  printf("%d %d %s\\n", vtb_version(), rc, vtb_last_error());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi_probe"
    lib_dir = os.path.dirname(lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", lib_dir, "-l:libvtb200.so", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(None, 2)
    assert int(out[0]) >= 100 and int(out[1]) != 0 and "null pointer" in out[2]
