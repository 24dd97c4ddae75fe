"""Build libhmrm.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python heightmap-ray-marcher_b200/build.py [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libhmrm.so"
HOST = PKG / "host"
HMAP = PKG / "hmap"          # the reference's binary name; headless host over libhmrm.so

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # exact mode: no FMA contraction anywhere (north star)
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall",
    "-shared",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def deps() -> list[Path]:
    return sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "hmrm.h", Path(__file__)])


def build_hmap(force: bool = False) -> Path:
    """g++ host (config grammar, image I/O, CLI) linked against libhmrm.so."""
    srcs = sorted(HOST.glob("*.cpp"))
    newest = max(p.stat().st_mtime for p in list(HOST.glob("*")) + [PKG.parent / "include" / "hmrm.h", LIB])
    if not force and HMAP.exists() and HMAP.stat().st_mtime >= newest:
        return HMAP
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-ffp-contract=off", "-fwrapv", "-o", str(HMAP), *[str(s) for s in srcs],
           "-L", str(PKG), "-lhmrm", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True, cwd=str(PKG))
    return HMAP


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in deps())


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("HMRM_NVCC_EXTRA", "").split()      # experiments only (e.g. -DHMRM_LIN_THREADS=128)
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-o", str(LIB), *[str(s) for s in sources()]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=str(PKG))
    return LIB


def build_variant(name: str, defines: list[str]) -> Path:
    """Another build of the library with extra -D flags, for kernel A/B experiments (tools/): variants/libhmrm_<name>.so,
    selected with HMRM_LIBRARY.  Never loaded by the product or the tests."""
    out = PKG / "variants" / f"libhmrm_{name}.so"
    out.parent.mkdir(exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *defines, "-o", str(out), *[str(s) for s in sources()]]
    subprocess.run(cmd, check=True, cwd=str(PKG))
    return out


def build_hmap_fake_sdl(shim_dir: Path, fake_sdl_obj: Path, out: Path) -> Path:
    """Interactive `hmap` (-DHMAP_WITH_SDL) linked against a scripted fake SDL: test builds only (the shim and
    the object live with the test oracle; SDL2 itself is not installable in this image)."""
    srcs = sorted(HOST.glob("*.cpp"))
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-ffp-contract=off", "-DHMAP_WITH_SDL", "-I", str(shim_dir),
           "-o", str(out), *[str(s) for s in srcs], str(fake_sdl_obj), "-L", str(PKG), "-lhmrm", "-lz", "-pthread",
           f"-Wl,-rpath,{PKG}"]
    subprocess.run(cmd, check=True, cwd=str(PKG))
    return out


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose=True)
    build_hmap(force="--force" in sys.argv)
    print(LIB)
    print(HMAP)
