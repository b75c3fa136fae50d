"""CPU: the C-ABI library loads and exports exactly what include/mvdetr_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "mvdetr_b200.h")


def declared_functions():
    text = open(HEADER).read()
    return re.findall(r"^MVD_API\s+[\w\s\*]+?\b(mvd_\w+)\s*\(", text, flags=re.M)


def test_header_declares_the_path():
    names = declared_functions()
    for required in ("mvd_msda_fwd_f32", "mvd_msda_fwd_f64", "mvd_msda_bwd_f32", "mvd_msda_bwd_f64",
                     "mvd_warp_fwd_f32", "mvd_warp_bwd_f32", "mvd_msda_fused_fwd_f32", "mvd_msda_fwd_viewgrid_f32"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from mvdetr_b200 import _C
    lib = ctypes.CDLL(_C.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _C.SIGNATURES, f"{name} has no ctypes signature in mvdetr_b200/_C.py"
    assert set(_C.SIGNATURES) == set(declared_functions())


def test_header_compiles_as_c():
    src = '#include "mvdetr_b200.h"\nint main(void){return MVD_OK;}\n'
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), "-x", "c", "-", "-o",
                        "/dev/null"], input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_version_and_error_strings():
    from mvdetr_b200 import _C
    assert _C.lib.mvd_version() >= 100
    assert _C.error_string(0) == "ok"
    for code in (-1, -2, -3, -4, -5):
        assert "mvdetr_b200" in _C.error_string(code)
    assert _C.error_string(1)  # a cudaError_t decodes to the runtime's text


def test_null_and_shape_errors_do_not_touch_the_gpu():
    from mvdetr_b200 import _C
    assert _C.lib.mvd_msda_fwd_f32(None, None, None, None, None, 1, 1, 1, 4, 1, 1, 1, None, None) == -1
    one = ctypes.c_void_p(16)
    assert _C.lib.mvd_msda_fwd_f32(one, one, one, one, one, 0, 1, 1, 4, 1, 1, 1, one, None) == -2
    assert _C.lib.mvd_warp_fwd_f32(None, None, 1, 1, 1, 1, 1, 1, None, 0, None) == -1
    assert _C.lib.mvd_warp_fwd_f32(one, one, 1, 0, 1, 1, 1, 1, one, 0, None) == -2


def test_missing_library_fails_loudly():
    code = ("import os, os.path as p\n_real = p.exists\n"
            "p.exists = lambda x: False if str(x).endswith('libmvdetr_b200.so') else _real(x)\n"
            "try:\n    import mvdetr_b200\nexcept ImportError as e:\n    print('IMPORTERROR', e)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, cwd=REPO, text=True)
    assert "IMPORTERROR" in r.stdout and "no CPU or PyTorch fallback" in r.stdout, r.stdout + r.stderr


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "mvdetr_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "torch_port" not in text, f


def test_every_binding_call_passes_the_declared_number_of_arguments():
    """Static check (no GPU): each `_C.lib.mvd_*(...)` call in the package and the tests has the arity of
    _C.SIGNATURES (a mismatch only shows up as a TypeError when the call is reached on the GPU box)."""
    import ast
    import glob
    from mvdetr_b200 import _C
    files = glob.glob(os.path.join(REPO, "mvdetr_b200", "*.py")) + glob.glob(os.path.join(REPO, "tests", "*.py")) + \
        [os.path.join(REPO, "bench.py")]
    seen = 0
    for f in files:
        for node in ast.walk(ast.parse(open(f).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in _C.SIGNATURES \
                    and isinstance(node.func.value, ast.Attribute) and node.func.value.attr == "lib":
                want = len(_C.SIGNATURES[node.func.attr])
                starred = [a for a in node.args if isinstance(a, ast.Starred)]
                if starred:  # ops.py passes `*aux` = (attn_out, loc_out)
                    assert len(node.args) - len(starred) + 2 * len(starred) == want, (f, node.lineno, node.func.attr)
                else:
                    assert len(node.args) == want, (f, node.lineno, node.func.attr)
                seen += 1
    assert seen >= 15


def test_hot_kernels_are_sm100a_native_in_sass():
    """SASS of the built library (no GPU needed): the hot-path kernels stage through TMA (UTMALDG) and mbarriers (SYNCS),
    the GEMM issues tcgen05 MMAs with the A operand in tensor memory (UTCHMMA tmem[..], gdesc[..]) fed by tcgen05.st
    (STTM), and none of them carries the uniform-register waterfall (BRA.U.ANY) that a divergent issue loop gets."""
    import shutil
    from mvdetr_b200 import _C
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", _C.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    per_kernel, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = []
        elif name is not None:
            per_kernel[name].append(line)

    def kernels(tag):
        ks = {k: "\n".join(v) for k, v in per_kernel.items() if tag in k}
        assert ks, f"no kernel named *{tag}* in the library"
        return ks

    for tag in ("linear_split_ts_kernel", "warp_tma_cl_kernel", "msda_vg_kernel"):
        for k, text in kernels(tag).items():
            assert "UTMALDG" in text and "SYNCS" in text, k
            assert "BRA.U.ANY" not in text, f"{k}: issue loop fell back to a uniform-register waterfall"
    for k, text in kernels("linear_split_ts_kernel").items():
        assert re.search(r"UTCHMMA tmem\[\w+\], gdesc\[\w+\], tmem\[", text), f"{k}: MMA does not take A from tensor memory"
        assert "STTM" in text and "LDTM" in text and "UTCBAR" in text, k


def test_every_export_is_mapped_to_a_reference_interface_in_the_docs():
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    missing = [n for n in declared_functions() if n not in doc]
    assert not missing, f"INTEGRATION.md does not mention {missing}"
    header = open(HEADER).read()
    assert header.count("ref:") + header.count("ref ") >= 10   # entry points cite the reference interface they replace
