"""CPU: the C-ABI library loads, exports every symbol include/rabitq_b200.h declares, and refuses to compute
without a CUDA device (no CPU fallback).  No compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rb():
    import rabitq_b200
    from rabitq_b200 import build

    build.build()
    rabitq_b200.lib()
    return rabitq_b200


def test_header_symbols_all_exported(rb):
    hdr = open(os.path.join(ROOT, "include", "rabitq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rabitq_[a-z_0-9]+)\s*\(", hdr))
    declared.discard("rabitq_index")
    assert declared == set(rb.ABI_SYMBOLS)
    L = rb.lib()
    for s in declared:
        assert hasattr(L, s), s


def test_no_torch_types_in_header():
    hdr = open(os.path.join(ROOT, "include", "rabitq_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # comments may mention torch; declarations may not
    assert "torch" not in code.lower() and "std::" not in code and "at::" not in code and "#include <cuda" not in code


def test_io_errors_like_reference(rb, tmp_path):
    with pytest.raises(rb.RabitqError) as e:
        rb.RaBitQ.load_from_dir(str(tmp_path / "missing"))
    assert e.value.code == 1  # RABITQ_EIO


def test_product_never_touches_oracle():
    """The product tree must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "rabitq_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("against the oracle", "").replace("the oracle)", "").lower() or f == "__init__.py" and "import oracle" not in src, f


def test_no_gpu_means_loud_failure(rb, case_d128):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = case_d128["arrays"]
    with pytest.raises(rb.RabitqError) as e:
        rb.RaBitQ.from_arrays(a["dim"], a["base"], a["orthogonal"], a["centroids"], a["offsets"], a["map_ids"], a["codes"], a["factors"])
    assert e.value.code == 3  # RABITQ_ECUDA
    assert "no CPU fallback" in str(e.value)


def test_service_and_cli_twins_fail_loudly_without_an_index():
    """The host-side twins are built in-tree; without a loadable index (and, here, without a device) they exit non-zero
    with the reference's panic text instead of serving anything."""
    import subprocess

    from rabitq_b200 import build as bld

    bld.build()
    r = subprocess.run([bld.SERVICE, "-d", "/nonexistent/dir", "-p", "0"], capture_output=True, text=True, timeout=30)
    assert r.returncode != 0 and "orthogonal.fvecs" in r.stderr
    r = subprocess.run([bld.SERVICE], capture_output=True, text=True, timeout=30)
    assert r.returncode == 2 and "--dir is required" in r.stderr


def test_every_option_is_documented_in_the_header():
    """`rabitq_set_option` names (rabitq_capi.cu) <-> the option list in include/rabitq_b200.h: a knob nobody can find is not an API."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "rabitq_b200", "csrc", "rabitq_capi.cu")).read()
    body = src[src.index("int rabitq_set_option("):]
    body = body[:body.index("return RABITQ_OK;")]
    names = set(re.findall(r'n == "([a-z0-9_]+)"', body))
    assert len(names) >= 15
    header = open(os.path.join(root, "include", "rabitq_b200.h")).read()
    missing = sorted(n for n in names if f'"{n}"' not in header)
    assert not missing, f"options without documentation in the header: {missing}"
