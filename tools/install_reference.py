"""Install the UNMODIFIED reference where the GPU box can import it: baseline/_ref/ (git-ignored, travels with gpurun).

The contract's recipe `pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` fails here
("Neither 'setup.py' nor 'pyproject.toml' found": the reference is a plain directory of Python packages, not a
distribution), so this does what a pure-Python install does -- it copies the package directories the hot path needs
(Transformer/, utils/) and the one dataset file the default script reads.  Nothing from baseline/_ref is imported by
the product (nlpscratch_b200/); it is used by
  * bench.py --impl reference   (times the reference's own Transformer on the host cores), and
  * tests/test_gpu_reference_driver.py (runs the reference's own utils/train.py::train_transformer, unmodified,
    against nlpscratch_b200.Transformer).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
WANT = ("Transformer", "utils", os.path.join("dataset", "tagged_train_mini.txt"))


def install(src="/root/reference", dest=DEST, quiet=False):
    if not os.path.isdir(src):
        return os.path.isdir(os.path.join(dest, "utils"))      # GPU box: use what travelled with the snapshot
    for rel in WANT:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    if not quiet:
        print("reference installed to", dest)
    return True


def reference_path():
    """Directory to put on sys.path to import the reference: the mounted tree when present, else the installed copy."""
    if os.path.isdir("/root/reference/utils"):
        return "/root/reference"
    if os.path.isdir(os.path.join(DEST, "utils")):
        return DEST
    return None


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
