#!/usr/bin/env python
"""Build the UNMODIFIED reference `ransac_voting` CUDA extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is imported by the product
package (`clean_pvnet_b200/`); only tests/, __graft_entry__.smoke() and
bench.py (reference arm / cpu_baseline leg) may touch it.

What this does (all outputs land in oracle/_ref/, which is git-ignored but
travels to the GPU box with gpurun):

  1. copies /root/reference/lib/csrc/ransac_voting/{src/*,ransac_voting_gpu.py}
     to a scratch dir under /tmp (the reference tree is read-only and its
     sources must never enter this repository's history);
  2. applies the torch-2.x compatibility patches listed in SURVEY.md §8c --
     no algorithmic change:
       * src/ransac_voting.cpp:5   `extern THCState* state;`  (unused; THC is gone)
       * ransac_voting_gpu.py:2    import path -> plain `import ransac_voting`
       * ransac_voting_gpu.py:36,142  masked_select needs a bool mask on torch>=1.2
       * torch.solve (a stub that only raises in torch 2.x) -> torch.linalg.solve shim, otherwise
         b_inv's bare `except` silently returns the identity (ransac_voting_gpu.py:105-108) and the
         reference returns ATb -- garbage of magnitude 1e4 px -- without any error
  3. compiles src/ransac_voting_kernel.cu UNMODIFIED with torch's
     BuildExtension for sm_100 (the reference setup.py passes no arch flags;
     TORCH_CUDA_ARCH_LIST=10.0 is what a user on a B200 would get).

Usage:  python oracle/build_ref.py            (no-op if /root/reference is absent)
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("PVNET_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "lib", "csrc", "ransac_voting")


def _patch_cpp(text):
    out = []
    for line in text.splitlines(keepends=True):
        if re.match(r"\s*extern\s+THCState\s*\*\s*state\s*;", line):
            out.append("// [oracle/build_ref.py] removed unused THCState extern (THC no longer exists)\n")
        else:
            out.append(line)
    return "".join(out)


def _patch_py(text):
    text = text.replace(
        "import lib.csrc.ransac_voting.ransac_voting as ransac_voting",
        "import ransac_voting  # [oracle/build_ref.py] import path alias\n"
        "def _pvb_solve_works():  # [oracle/build_ref.py] torch.solve exists in torch 2.x but only raises\n"
        "    try:\n"
        "        torch.solve(torch.eye(2), torch.eye(2))\n"
        "        return True\n"
        "    except Exception:\n"
        "        return False\n"
        "if not _pvb_solve_works():\n"
        "    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)")
    # masked_select requires a bool mask on modern torch
    text = text.replace(
        "masked_select(torch.unsqueeze(torch.unsqueeze(cur_mask, 2), 3))",
        "masked_select(torch.unsqueeze(torch.unsqueeze(cur_mask.bool(), 2), 3))")
    return text


def build(force=False):
    if not os.path.isdir(SRC):
        print(f"[build_ref] {SRC} not present; nothing to do (prebuilt oracle/_ref is used as-is)")
        return False
    os.makedirs(OUT, exist_ok=True)
    so = [f for f in os.listdir(OUT) if f.startswith("ransac_voting") and f.endswith(".so")]
    if so and os.path.exists(os.path.join(OUT, "ransac_voting_gpu.py")) and not force:
        print(f"[build_ref] up to date: {so[0]}")
        return True
    tmp = tempfile.mkdtemp(prefix="pvnet_ref_build_")
    try:
        os.makedirs(os.path.join(tmp, "src"))
        for f in ("cuda_common.h", "ransac_voting_kernel.cu"):
            shutil.copy(os.path.join(SRC, "src", f), os.path.join(tmp, "src", f))
        with open(os.path.join(SRC, "src", "ransac_voting.cpp")) as fh:
            cpp = _patch_cpp(fh.read())
        with open(os.path.join(tmp, "src", "ransac_voting.cpp"), "w") as fh:
            fh.write(cpp)
        shutil.copy(os.path.join(SRC, "setup.py"), os.path.join(tmp, "setup.py"))
        env = dict(os.environ)
        env["TORCH_CUDA_ARCH_LIST"] = "10.0"
        env.setdefault("MAX_JOBS", "4")
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, env=env)
        built = [f for f in os.listdir(tmp) if f.startswith("ransac_voting") and f.endswith(".so")]
        assert built, "reference extension did not build"
        for f in built:
            shutil.copy(os.path.join(tmp, f), os.path.join(OUT, f))
        with open(os.path.join(SRC, "ransac_voting_gpu.py")) as fh:
            py = _patch_py(fh.read())
        with open(os.path.join(OUT, "ransac_voting_gpu.py"), "w") as fh:
            fh.write(py)
        print(f"[build_ref] built {built} -> {OUT}")
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    build(force="--force" in sys.argv)
