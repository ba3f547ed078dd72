"""The header is the product's boundary: it must compile as plain C and link against the shared library."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_c_and_links(tmp_path):
    from zkemail_rs_b200.engine import LIB_PATH, load_library
    load_library()
    exe = tmp_path / "c_abi_smoke"
    libdir = os.path.dirname(LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "smoke.c"), "-o", str(exe),
                           "-L", libdir, "-l:libzkemail_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok")
