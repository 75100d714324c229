"""Recipe for the reference arm of bench.py: puts the UNMODIFIED reference modules on the hot path into baseline/_ref/.

    python baseline/make_ref.py

The reference (aclyde11/molecular-VAE) is 20 flat scripts with no setup.py / pyproject, so
`pip install --target baseline/_ref /root/reference` has nothing to build (recorded in DESIGN.md); its importable
modules are copied byte for byte instead.  baseline/_ref/ is git-ignored -- reference sources never
enter this repository's history -- but not gpurun-ignored, so the files travel to the GPU box the same way the built
.so does.  __graft_entry__.build() runs this whenever /root/reference is present.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MVAE_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
# models on the path + what they import + train.py (only its loss_function, lines 31-38, is ever executed: AST-extracted)
FILES = ["models.py", "models2d.py", "mosesvae.py", "mosesfile.py", "vocab.py", "featurizer.py", "config.py", "train.py"]


def make_ref(verbose=True):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        lines.append(f"{hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print(f"baseline/_ref: {len(FILES)} reference modules copied from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
