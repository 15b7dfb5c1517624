"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes, glob, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(uco_b200_\w+)\s*\(", src))
    return names


def test_library_built_and_exports_every_declared_symbol():
    import ucoslam_b200
    assert os.path.exists(ucoslam_b200.LIB_PATH), "run python ucoslam-cv3_b200/build.py"
    lib = ctypes.CDLL(ucoslam_b200.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 9
    for n in declared:
        assert hasattr(lib, n), "missing export " + n
    # the Python binding covers the same set
    assert set(ucoslam_b200.SIGNATURES) == declared


def test_version_and_no_cpu_fallback():
    import torch, ucoslam_b200
    lib = ucoslam_b200.load()
    assert lib.uco_b200_version() >= 100
    if not torch.cuda.is_available():
        # product path must fail loudly without a device
        with pytest.raises(ucoslam_b200.UcoError):
            ucoslam_b200.Context(0)


def test_sm100a_sass_uses_tma_and_popc():
    """The k-NN kernel is built for sm_100a and stages train tiles with TMA bulk copies (UBLKCP in SASS)."""
    import subprocess, shutil, ucoslam_b200
    if not shutil.which("cuobjdump"):
        pytest.skip("no cuobjdump")
    sass = subprocess.run(["cuobjdump", "-sass", ucoslam_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass and "POPC" in sass
