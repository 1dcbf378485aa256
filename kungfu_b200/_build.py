"""In-tree build of the native pieces of kungfu_b200.

Counterpart of the reference's CMake + setup.py build (CMakeLists.txt:9-15,53-73, setup_pytorch.py:21-53),
without Go/TensorFlow: everything is C++17 / CUDA C++ for sm_100a.

Targets
  _kfcuda   : CUDA kernel library + PeerMesh + pybind11 bindings (nvcc, `-gencode arch=compute_100a,code=sm_100a`)
  _kfrt     : CPU runtime (transport, session, peer, store, config server, monitor...) + bindings (g++)
  kungfu-run: the launcher binary (g++)

All artefacts are written next to the sources (`kungfu_b200/*.so`, `kungfu_b200/bin/*`) so that they
travel with the repository snapshot; a JIT cache outside the tree would not.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
BIN = PKG / "bin"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# NB: $CXX in this image points at a wrapper that links libstdc++ statically; a second libstdc++ inside a process that
# also hosts libtorch crashes in iostream/locale code. Always use the system compiler (dynamic libstdc++) unless
# KUNGFU_CXX overrides it.
CXX = os.environ.get("KUNGFU_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++"))
CUDA_HOME = os.environ.get("CUDA_HOME") or str(Path(NVCC).resolve().parent.parent)

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-ccbin", CXX, "-lineinfo", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler",
                           "-fvisibility=hidden", "--expt-relaxed-constexpr", "-diag-suppress", "177,550,2361"]
CXX_FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", "-pthread", "-fvisibility=hidden"]
LINK_FLAGS: list[str] = []
# Sanitizer presets for the C++ runtime (the reference configures none): KUNGFU_SANITIZE=thread | address
_san = os.environ.get("KUNGFU_SANITIZE", "")
if _san in ("thread", "address"):
    _f = ["-fsanitize=" + ("thread" if _san == "thread" else "address,undefined"), "-fno-omit-frame-pointer"]
    CXX_FLAGS += _f
    LINK_FLAGS += _f
    OBJ = PKG / ("build-" + _san)
    BIN = PKG / ("bin-" + _san)


def _pybind_includes() -> list[str]:
    import pybind11

    return ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]


def _ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build command failed:\n  " + " ".join(cmd) + "\n" + r.stdout)


def _stale(out: Path, deps: list[Path], cmd_sig: str) -> bool:
    sig_file = out.with_suffix(out.suffix + ".sig")
    if not out.exists() or not sig_file.exists():
        return True
    if sig_file.read_text() != cmd_sig:
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def _headers(*dirs: Path) -> list[Path]:
    hs: list[Path] = []
    for d in dirs:
        hs += list(d.glob("*.h")) + list(d.glob("*.cuh")) + list(d.glob("*.hpp"))
    return hs


def _compile(src: Path, obj: Path, cmd: list[str], deps: list[Path]) -> bool:
    sig = hashlib.sha1(" ".join(cmd).encode()).hexdigest()
    if not _stale(obj, [src] + deps, sig):
        return False
    obj.parent.mkdir(parents=True, exist_ok=True)
    _run(cmd)
    obj.with_suffix(obj.suffix + ".sig").write_text(sig)
    return True


def build_cuda(verbose: bool = False, force: bool = False) -> Path:
    """Compile every .cu for sm_100a and link `_kfcuda`."""
    out = PKG / ("_kfcuda" + _ext_suffix())
    kdir, mdir, bdir, edir = CSRC / "kernels", CSRC / "mesh", CSRC / "bindings", CSRC / "engine"
    hdrs = _headers(kdir, mdir, edir)
    jobs = []
    objs = []
    for cu in sorted(kdir.glob("*.cu")):
        obj = OBJ / "cuda" / (cu.stem + ".o")
        extra = []
        cmd = [NVCC] + NVCC_FLAGS + extra + ["-I" + str(kdir), "-c", str(cu), "-o", str(obj)]
        jobs.append((cu, obj, cmd))
        objs.append(obj)
    host_inc = ["-I" + str(kdir), "-I" + str(mdir), "-I" + CUDA_HOME + "/include"] + _pybind_includes()
    for cpp in [mdir / "peer_mesh.cpp", edir / "grad_engine.cpp", bdir / "cuda_ext.cpp"]:
        obj = OBJ / "cuda" / (cpp.stem + ".o")
        cmd = [CXX] + CXX_FLAGS + host_inc + ["-c", str(cpp), "-o", str(obj)]
        jobs.append((cpp, obj, cmd))
        objs.append(obj)
    if force:
        for _, obj, _ in jobs:
            obj.unlink(missing_ok=True)
    changed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        futs = [ex.submit(_compile, s, o, c, hdrs) for s, o, c in jobs]
        for f in futs:
            changed |= bool(f.result())
    if changed or not out.exists():
        _run([NVCC, "-ccbin", CXX] + ARCH_FLAGS + ["-shared", "-o", str(out)] + [str(o) for o in objs] + ["-lpthread"])
        if verbose:
            print(f"[kungfu_b200] linked {out.name}")
    return out


def build_runtime(verbose: bool = False, force: bool = False) -> Path | None:
    """Compile the CPU runtime + launcher (pure C++, no CUDA)."""
    rdir, ldir = CSRC / "runtime", CSRC / "launcher"
    srcs = sorted(rdir.glob("*.cpp"))
    if not srcs:
        return None
    out = (PKG if not _san else OBJ) / ("_kfrt" + _ext_suffix())
    OBJ.mkdir(parents=True, exist_ok=True)
    hdrs = _headers(rdir, ldir)
    inc = ["-I" + str(rdir)] + _pybind_includes()
    jobs, objs = [], []
    for cpp in srcs:
        obj = OBJ / "rt" / (cpp.stem + ".o")
        jobs.append((cpp, obj, [CXX] + CXX_FLAGS + inc + ["-c", str(cpp), "-o", str(obj)]))
        objs.append(obj)
    if force:
        for _, obj, _ in jobs:
            obj.unlink(missing_ok=True)
    changed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        futs = [ex.submit(_compile, s, o, c, hdrs) for s, o, c in jobs]
        for f in futs:
            changed |= bool(f.result())
    if changed or not out.exists():
        _run([CXX, "-shared", "-o", str(out)] + [str(o) for o in objs] + LINK_FLAGS + ["-pthread", "-ldl"])
        if verbose:
            print(f"[kungfu_b200] linked {out.name}")
    # launcher + tools: every launcher/*.cpp with a main() links against the runtime objects (minus bindings)
    BIN.mkdir(exist_ok=True)
    lib_objs = [str(o) for o in objs if o.stem != "bindings"]
    # libkungfu-comm.so: the runtime with its C ABI (csrc/runtime/capi.h) and NO dependency on Python — what a C / C++ / foreign-language
    # program links against (reference: the cgo c-archive libkungfu-comm + libkungfu.so, srcs/cmake/kungfu.cmake)
    libdir = PKG / ("lib" if not _san else "lib-" + _san)
    libdir.mkdir(exist_ok=True)
    comm = libdir / "libkungfu-comm.so"
    if changed or not comm.exists():
        _run([CXX, "-shared", "-o", str(comm)] + lib_objs + LINK_FLAGS + ["-pthread", "-ldl"])
        if verbose:
            print(f"[kungfu_b200] linked lib/{comm.name}")
    tdir = CSRC / "tests"
    for cpp in sorted(ldir.glob("*.cpp")) + sorted(tdir.glob("*.cpp")):
        exe = BIN / cpp.stem.replace("_", "-")
        obj = OBJ / "launcher" / (cpp.stem + ".o")
        c = _compile(cpp, obj, [CXX] + CXX_FLAGS + ["-I" + str(rdir), "-I" + str(ldir), "-c", str(cpp), "-o", str(obj)], hdrs)
        if c or changed or not exe.exists():
            _run([CXX, "-o", str(exe), str(obj)] + lib_objs + LINK_FLAGS + ["-pthread", "-ldl"])
            if verbose:
                print(f"[kungfu_b200] linked bin/{exe.name}")
    dist = BIN / "kungfu-distribute"  # the same binary under its second name (argv[0] selects the mode)
    if (BIN / "kungfu-rrun").exists() and not dist.exists():
        try:
            os.symlink("kungfu-rrun", dist)
        except OSError:
            shutil.copy2(BIN / "kungfu-rrun", dist)
    return out


def build_all(verbose: bool = False, force: bool = False) -> None:
    build_runtime(verbose, force)
    build_cuda(verbose, force)


if __name__ == "__main__":
    build_all(verbose=True, force="--force" in sys.argv)
