"""The drop-in boundary used from compiled code: include/zgml_cuda.h must be valid C99 and C++17, and a plain-C
program linked against libzgml_cuda.so must reproduce the reference's own known-answer qmatmul program
(src/backend/reference.zig:710-761) — the same calls the Zig translator (zig/cuda.zig) makes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
SRC = os.path.join(ROOT, "tests", "cpp", "cabi_golden.c")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")


def test_header_is_valid_c99_and_cpp17():
    hdr = os.path.join(INC, "zgml_cuda.h")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)


def _build_c_program():
    from zgml_b200.build import build_cuda
    lib = build_cuda()
    os.makedirs(OUT_DIR, exist_ok=True)
    exe = os.path.join(OUT_DIR, "cabi_golden")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INC, SRC, "-o", exe, "-L", os.path.dirname(lib), "-lzgml_cuda",
                    f"-Wl,-rpath,{os.path.dirname(lib)}", "-lm"], check=True)
    return exe


def test_c_program_links_against_the_library():
    assert os.path.exists(_build_c_program())


@pytest.mark.gpu
def test_reference_known_answer_program_from_plain_c():
    exe = _build_c_program()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
