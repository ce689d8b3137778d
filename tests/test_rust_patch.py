"""CPU: the reference-side binding ships as an applicable patch: integration/wrenc_b200.patch adds
src/b200.rs (extern "C" block + safe wrapper incl. submit_pinned / decisions), build.rs, a `b200` cargo feature, and the
`--b200` / `--b200-check` picture loop of main.rs.  It must apply cleanly to the reference tree, and every C-ABI function it
binds must be declared in include/wrenc_b200.h with the same arity.  (It cannot be compiled here: no rustc / cargo.)"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "wrenc_b200.patch")
REF = "/root/reference"


def test_patch_binds_only_declared_functions_with_matching_arity():
    patch = open(PATCH).read()
    header = open(os.path.join(ROOT, "include", "wrenc_b200.h")).read()
    bound = re.findall(r"^\+\s+fn (wrenc_b200_\w+)\(([^;]*?)\)(?:\s*->\s*[^;]+)?;", patch, re.S | re.M)
    assert len(bound) >= 10
    for name, params in bound:
        m = re.search(r"\b%s\(([^;]*?)\);" % name, header, re.S)
        assert m, f"{name} is not declared in include/wrenc_b200.h"
        n_rust = len([p for p in params.replace("\n+", " ").split(",") if p.strip()])
        c_params = m.group(1).strip()
        n_c = 0 if c_params in ("", "void") else len(c_params.split(","))
        assert n_rust == n_c, f"{name}: {n_rust} parameters in the Rust binding, {n_c} in the header"
    # the config struct mirrors the header field by field
    rust_fields = re.findall(r"^\+\s+pub (\w+): (?:i32|\*const c_char),", patch, re.M)[:10]
    c_fields = re.findall(r"^\s+(?:int32_t|const char \*)\s*([\w, ]+);", header[header.index("typedef struct {"):header.index("} wrenc_b200_config;")], re.M)
    c_fields = [f.strip() for grp in c_fields for f in grp.split(",")]
    assert rust_fields == c_fields, (rust_fields, c_fields)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="/root/reference not present")
def test_patch_applies_to_the_reference_tree(tmp_path):
    work = tmp_path / "wrenc"
    work.mkdir()
    shutil.copytree(os.path.join(REF, "src"), work / "src")
    shutil.copy(os.path.join(REF, "Cargo.toml"), work / "Cargo.toml")
    r = subprocess.run(["git", "apply", "--check", "--verbose", PATCH], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["git", "apply", PATCH], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    main = (work / "src" / "main.rs").read_text()
    assert "mod b200;" in main and "encode_with_slice_data" in main and "submit_pinned" in main and "b200_check" in main
    assert (work / "src" / "b200.rs").exists() and (work / "build.rs").exists()
    assert "b200 = []" in (work / "Cargo.toml").read_text()
    # balanced delimiters in the new / edited Rust sources (the closest thing to a syntax check without rustc)
    for f in ("src/b200.rs", "src/main.rs", "src/slice_encoder.rs", "build.rs"):
        txt = re.sub(r'"(?:\\.|[^"\\])*"', '""', (work / f).read_text())
        txt = re.sub(r"//[^\n]*", "", txt)
        txt = re.sub(r"'(?:\\.|[^'\\])'", "' '", txt)
        for a, b in ("()", "[]", "{}"):
            assert txt.count(a) == txt.count(b), f"{f}: unbalanced {a}{b}"
