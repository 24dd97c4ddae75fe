"""Host image ingest under damage: every golden fixture truncated and with bytes overwritten, decoded by the host's
decoders built with AddressSanitizer + UndefinedBehaviorSanitizer (tests/fuzz/decode_fuzz.cpp).  A damaged file may
decode or be refused; it must not crash, read out of bounds or hit undefined behaviour.  (stb_image, which the
reference uses, is the same kind of code for the same job: vendor/stb_image.h.)  20 variants per file here (~40 s);
`decode_fuzz 300 tests/golden/images/*` is the long run (21 300 variants, clean)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "heightmap-ray-marcher_b200" / "host"
IMAGES = ROOT / "tests" / "golden" / "images"


@pytest.fixture(scope="module")
def fuzz_binary(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = tmp_path_factory.mktemp("fuzz") / "decode_fuzz"
    srcs = [str(ROOT / "tests" / "fuzz" / "decode_fuzz.cpp")] + sorted(str(p) for p in HOST.glob("image_*.cpp"))
    cmd = ["g++", "-std=c++11", "-O1", "-g", "-fwrapv", "-ffp-contract=off", "-fsanitize=address,undefined",
           "-fno-sanitize-recover=undefined", "-I", str(HOST), *srcs, "-lz", "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr:
        pytest.skip("sanitizer runtime not available")
    assert res.returncode == 0, res.stderr[-2000:]
    return out


def test_damaged_images_never_crash_the_decoders(fuzz_binary):
    files = sorted(str(p) for p in IMAGES.iterdir())
    assert len(files) >= 71
    res = subprocess.run([str(fuzz_binary), "20", *files], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    decoded, refused = (int(res.stdout.split()[0]), int(res.stdout.split()[2]))
    assert decoded + refused == 20 * len(files) and refused > 0 and decoded > 0
