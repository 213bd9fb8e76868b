"""Recipe for the live-reference CPU arm: installs the UNMODIFIED reference into the git-ignored baseline/_ref/ (it travels to
the GPU box with the snapshot; /root/reference does not exist there).

  python baseline/install_ref.py            (build() in __graft_entry__.py runs it whenever /root/reference is present)

Step 1 is the sanctioned offline install (pip --no-index --target baseline/_ref, from a copy under /tmp because the build writes
into the source tree).  The reference's setup.py lists packages=["codae"] only, so pip installs codae/__init__.py and none of
the subpackages; step 2 completes that install with the three subpackages of the hot path (codae/model, codae/tool,
codae/dataset: pure Python, BSD-2, LICENSE.md copied along) and its three YAML configs.  Nothing under baseline/_ref is tracked
by git, nothing in the product imports it: only bench.py's reference arm / cpu_baseline leg does (baseline/live_reference.py).
"""
import os
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def install(verbose=True):
    if not os.path.isdir(REF):
        return False
    shutil.rmtree(DST, ignore_errors=True)
    tmp = tempfile.mkdtemp(prefix="codae_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src)
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                            "--find-links", "/opt/wheelhouse", "--target", DST, src], capture_output=True, text=True)
        if r.returncode != 0:
            if verbose:
                print("pip install of the reference failed:", r.stderr[-500:])
            os.makedirs(os.path.join(DST, "codae"), exist_ok=True)
            shutil.copy(os.path.join(REF, "codae", "__init__.py"), os.path.join(DST, "codae", "__init__.py"))
        for sub in ("model", "tool", "dataset"):          # omitted by the reference's setup.py (packages=["codae"])
            d = os.path.join(DST, "codae", sub)
            shutil.rmtree(d, ignore_errors=True)
            shutil.copytree(os.path.join(REF, "codae", sub), d, ignore=shutil.ignore_patterns("__pycache__"))
        shutil.copytree(os.path.join(REF, "config"), os.path.join(DST, "config"), dirs_exist_ok=True)
        shutil.copy(os.path.join(REF, "LICENSE.md"), os.path.join(DST, "LICENSE.md"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print("reference installed into", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
