"""The Rust side of the boundary (ffi/) cannot be compiled here (no Rust toolchain); what can be checked is that the
generated `-sys` crate declares exactly the symbols of include/testudo_b200.h, with the argument counts of the header,
and that the wrapper / KAT sources only call functions that exist."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    return open(os.path.join(ROOT, *parts)).read()


def test_sys_crate_is_generated_from_the_current_header():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "scripts", "gen_ffi_sys.py"), "--check"]) == 0, \
        "ffi/testudo-b200-sys/src/lib.rs is stale: run scripts/gen_ffi_sys.py"


def test_sys_crate_matches_header_symbol_by_symbol():
    header = re.sub(r"/\*.*?\*/", " ", _read("include", "testudo_b200.h"), flags=re.S)
    rs = _read("ffi", "testudo-b200-sys", "src", "lib.rs")
    protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(tb200_\w+)\s*\(([^;{]*)\)\s*;", header, flags=re.S)}
    decls = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (tb200_\w+)\(([^)]*)\)", rs)}
    assert set(protos) == set(decls) and len(protos) > 80
    for name, args in protos.items():
        a = " ".join(args.split())
        n_c = 0 if a in ("", "void") else a.count(",") + 1
        n_rs = 0 if not decls[name].strip() else decls[name].count(":")
        assert n_c == n_rs, name
    # the ctypes table of the Python binding covers the same set
    from testudo_b200 import _lib
    assert set(_lib.SIGNATURES) == set(protos)


def test_wrapper_and_kat_only_call_declared_functions():
    rs = _read("ffi", "testudo-b200-sys", "src", "lib.rs")
    decls = set(re.findall(r"pub fn (tb200_\w+)\(", rs))
    used = set(re.findall(r"sys::(tb200_\w+)\(", _read("ffi", "testudo-b200", "src", "lib.rs")))
    assert used and used <= decls, used - decls
