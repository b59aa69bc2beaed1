"""Build recipe for ``oracle/_ref``: the UNMODIFIED reference CUDA extensions.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
path; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs use it.

The three reference extensions (``gridencoder/src``, ``shencoder/src``,
``freqencoder/src`` under ``/root/reference``) are compiled *from where they lie*
(no source is copied into this repository) into ``oracle/_ref/_ref_<name>.so`` for
``sm_100``.  Differences from the reference's own ``setup.py``
(``gridencoder/setup.py:7-13``):

* ``-std=c++17`` instead of ``-std=c++14`` (torch >= 2.1 headers ``#error`` on 14);
* an explicit ``-gencode arch=compute_100,code=sm_100`` (the reference passes no arch);
* ``TORCH_EXTENSION_NAME=_ref_<name>`` so the module cannot be confused with the
  product's own ``_gridencoder`` shim;
* freqencoder keeps its ``-use_fast_math`` (``freqencoder/backend.py:9``).

The resulting modules expose the reference's 8 pybind functions and need a GPU to
*run* (every entry point starts with ``CHECK_CUDA``), so they are exercised by the
``-m gpu`` tests on the B200 box, where ``/root/reference`` does not exist: the
``.so`` files travel there with the repo snapshot (git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("SANERF_REFERENCE", "/root/reference")

EXTS = {
    "gridencoder": (["gridencoder.cu", "bindings.cpp"], []),
    "shencoder": (["shencoder.cu", "bindings.cpp"], []),
    "freqencoder": (["freqencoder.cu", "bindings.cpp"], ["-use_fast_math"]),
}


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths() + [sysconfig.get_paths()["include"]]
    lib = ce.library_paths()
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, lib, abi


def build(force: bool = False, verbose: bool = True) -> dict:
    """Compile every reference extension; returns {name: path or None}."""
    os.makedirs(OUT, exist_ok=True)
    results = {}
    if not os.path.isdir(REFERENCE):
        # GPU box / CI without the reference checkout: use whatever is prebuilt.
        for name in EXTS:
            p = os.path.join(OUT, f"_ref_{name}.so")
            results[name] = p if os.path.exists(p) else None
        return results

    inc, lib, abi = _torch_flags()
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    for name, (srcs, extra) in EXTS.items():
        out = os.path.join(OUT, f"_ref_{name}.so")
        src_paths = [os.path.join(REFERENCE, name, "src", s) for s in srcs]
        newest = max(os.path.getmtime(s) for s in src_paths)
        if not force and os.path.exists(out) and os.path.getmtime(out) > newest:
            results[name] = out
            continue
        cmd = [nvcc, "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
               "-gencode", "arch=compute_100,code=sm_100",
               "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
               "-U__CUDA_NO_HALF2_OPERATORS__",
               f"-DTORCH_EXTENSION_NAME=_ref_{name}",
               "-DTORCH_API_INCLUDE_EXTENSION_H",
               f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
               "-w"] + extra
        for i in inc:
            cmd += ["-isystem", i]
        cmd += ["-I", os.path.join(REFERENCE, name, "src")]
        # bindings.cpp must be compiled as host C++ by nvcc as well.
        cmd += ["-x", "cu"] + src_paths
        for l in lib:
            cmd += ["-L", l, "-Xlinker", f"-rpath={l}"]
        cmd += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
                "-o", out]
        if verbose:
            print("[oracle/_ref] building", name, file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            if verbose:
                print(r.stdout[-4000:], r.stderr[-4000:], file=sys.stderr)
            results[name] = None
        else:
            results[name] = out
    return results


def load(name: str):
    """Import ``oracle/_ref/_ref_<name>.so`` (requires torch; needs a GPU to *call*)."""
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    path = os.path.join(OUT, f"_ref_{name}.so")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    res = build(force="--force" in sys.argv)
    for k, v in res.items():
        print(k, "->", v)
    sys.exit(0 if all(res.values()) else 1)
