"""Stage the reference's own Sinkhorn code for the CPU arm of bench.py (``cpu_baseline.kind = "reference"``).

TEST / BENCH INFRASTRUCTURE ONLY -- nothing in the product package imports ``oracle/``.

The reference is pure Python: there is nothing to compile.  The one file of it that holds the Sinkhorn
arithmetic of the hot path and imports with numpy + scipy alone is
``/root/reference/perturbot/perturbot/match/utils.py`` (``sinkhorn_scaling`` :6-115, the in-tree NumPy copy of
POT's Sinkhorn-Knopp that ``MRI_PET_OT_nojax.py:143`` runs; ``init_matrix_np`` :125-184).  ``/root/reference`` does
not exist on the GPU box, so ``__graft_entry__.build()`` runs this recipe in the build container: the file is
staged UNMODIFIED into ``oracle/_ref/`` (git-ignored, so no reference source enters the history; not
gpurun-ignored, so it travels to the box like the built ``.so``) together with a manifest holding its sha256.
``load()`` imports the staged file by path and verifies the checksum, so what bench.py times is byte for byte
the reference's code.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
REF_FILE = "perturbot/perturbot/match/utils.py"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_FILE = os.path.join(OUT_DIR, "perturbot_match_utils.py")
MANIFEST = os.path.join(OUT_DIR, "MANIFEST.json")


def _sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(verbose: bool = False) -> str | None:
    """Stage the reference file if the reference tree is present (build container); keep what is already
    staged otherwise (GPU box).  Returns the staged path or None."""
    src = os.path.join(REF_ROOT, REF_FILE)
    if os.path.exists(src):
        os.makedirs(OUT_DIR, exist_ok=True)
        shutil.copyfile(src, OUT_FILE)
        with open(MANIFEST, "w") as fh:
            json.dump({"source": f"{REF_ROOT}/{REF_FILE}", "staged_as": os.path.basename(OUT_FILE),
                       "sha256": _sha256(OUT_FILE), "functions": ["sinkhorn_scaling", "init_matrix_np"],
                       "modified": False}, fh, indent=1)
        if verbose:
            print(f"staged {src} -> {OUT_FILE}")
    return OUT_FILE if os.path.exists(OUT_FILE) else None


def load():
    """Import the staged reference module; None when it was never staged (then bench.py falls back to the
    oracle port and says ``kind = "port"``)."""
    if not (os.path.exists(OUT_FILE) and os.path.exists(MANIFEST)):
        return None
    with open(MANIFEST) as fh:
        man = json.load(fh)
    if man.get("sha256") != _sha256(OUT_FILE):
        raise RuntimeError("oracle/_ref/perturbot_match_utils.py does not match its manifest")
    spec = importlib.util.spec_from_file_location("ref_perturbot_match_utils", OUT_FILE)
    mod = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():  # the reference's docstrings hold unescaped backslashes
        warnings.simplefilter("ignore", SyntaxWarning)
        spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
