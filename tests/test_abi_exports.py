"""The drop-in boundary: every entry point include/fw_b200.h declares is exported by the product library (prefix fw_)
and by the oracle (prefix fwo_), and the ctypes table the tests bind through names exactly the same set — so nothing can
be declared but missing, or bound but undeclared. CPU only: loading the product library needs no GPU."""
import re
import subprocess
from pathlib import Path

from firewheel_b200 import _capi

ROOT = Path(__file__).resolve().parent.parent


def declared():
    text = (ROOT / "include" / "fw_b200.h").read_text()
    names = set(re.findall(r"FW_FN\((\w+)\)", text))
    names.discard("name")  # the macro's own parameter
    return names


def exported(path, prefix):
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1][len(prefix):] for line in out.splitlines() if len(line.split()) == 3 and line.split()[1] == "T" and line.split()[-1].startswith(prefix)}


def test_header_ctypes_table_and_both_libraries_agree(oracle, product):
    decl = declared()
    assert len(decl) >= 75
    assert set(_capi.SIGNATURES) == decl, (set(_capi.SIGNATURES) ^ decl)
    assert decl <= exported(product.path, "fw_"), decl - exported(product.path, "fw_")
    assert decl <= exported(oracle.path, "fwo_"), decl - exported(oracle.path, "fwo_")
    # the product must not reach into the oracle: no fwo_ symbol, no dependency on its library
    needed = subprocess.run(["readelf", "-d", product.path], capture_output=True, text=True).stdout
    assert "fw_oracle" not in needed and not exported(product.path, "fwo_")


def test_c_host_example_builds_against_the_header_and_links(product, tmp_path):
    """examples/host.c is the binding a maintainer would write: C99, include/fw_b200.h only, linked against the product library.
    Without a GPU it builds the graph, prints the compiled schedule and stops — the product has no CPU fallback."""
    exe = tmp_path / "host"
    lib_dir = Path(product.path).parent
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "examples" / "host.c"), "-o", str(exe),
                    "-L", str(lib_dir), "-lfirewheel_b200", "-lm", f"-Wl,-rpath,{lib_dir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "schedule: 4 nodes" in r.stdout and "beep_test" in r.stdout and "volume" in r.stdout
