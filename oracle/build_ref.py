"""Recipe for oracle/_ref/: the UNMODIFIED reference compiled to byte code (TEST / BASELINE INFRASTRUCTURE ONLY).

    python -m oracle.build_ref          # needs /root/reference (build container); __graft_entry__.build() calls it

The reference's hot path is pure Python (PointNet/models/*.py, PointNet/attacks/torchattacks/**), so "compiling it
from the sources where they lie" is ``py_compile``: every file the path imports is compiled from /root/reference
straight into ``oracle/_ref/PointNet/...`` as source-less byte code (no reference source text enters the repository;
``oracle/_ref/`` is git-ignored and travels to the GPU box with the gpurun snapshot exactly like the built ``.so``).
The files carry the suffix ``.pycode`` instead of ``.pyc`` (snapshots drop ``*.pyc``); a small meta-path finder maps the
reference's module names onto them.
``bench.py --impl reference`` and the ``cpu_baseline`` leg import the reference from there and time ITS OWN classes
(``torchattacks.tar_NB_attack`` over ``pointnet2_sem_seg.get_model``), kind "reference".
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/PointNet"
REF_OUT = os.path.join(HERE, "_ref", "PointNet")
SUFFIX = ".pycode"
FILES = [
    "models/pointnet_util.py", "models/pointnet2_sem_seg.py", "models/pointnet2_sem_seg_msg.py",
    "attacks/torchattacks/__init__.py", "attacks/torchattacks/attack.py",
    "attacks/torchattacks/attacks/__init__.py", "attacks/torchattacks/attacks/nontarget.py",
    "attacks/torchattacks/attacks/target.py",
]


def build() -> bool:
    """Returns True when oracle/_ref is usable (freshly compiled, or already present when the sources are absent)."""
    if not os.path.isdir(REF_SRC):
        return available()
    import warnings
    for rel in FILES:
        dst = os.path.join(REF_OUT, rel[:-3] + SUFFIX)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # the reference's `is "literal"` comparisons (attack.py:37,40)
            py_compile.compile(os.path.join(REF_SRC, rel), cfile=dst, dfile="reference/PointNet/" + rel, doraise=True, quiet=2)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_OUT, rel[:-3] + SUFFIX)) for rel in FILES)


class _RefFinder:
    """Meta-path finder for the reference's own module names (as its scripts import them with models/ and attacks/ on
    sys.path, NB_nontarget_test_semseg.py:18-20,33): byte code from oracle/_ref, nothing else."""

    def __init__(self):
        m, a = os.path.join(REF_OUT, "models"), os.path.join(REF_OUT, "attacks", "torchattacks")
        self.files = {
            "models.pointnet_util": os.path.join(m, "pointnet_util" + SUFFIX),
            "pointnet2_sem_seg": os.path.join(m, "pointnet2_sem_seg" + SUFFIX),
            "pointnet2_sem_seg_msg": os.path.join(m, "pointnet2_sem_seg_msg" + SUFFIX),
            "torchattacks": os.path.join(a, "__init__" + SUFFIX),
            "torchattacks.attack": os.path.join(a, "attack" + SUFFIX),
            "torchattacks.attacks": os.path.join(a, "attacks", "__init__" + SUFFIX),
            "torchattacks.attacks.nontarget": os.path.join(a, "attacks", "nontarget" + SUFFIX),
            "torchattacks.attacks.target": os.path.join(a, "attacks", "target" + SUFFIX),
        }
        self.packages = {"models": m, "torchattacks": a, "torchattacks.attacks": os.path.join(a, "attacks")}

    def find_spec(self, fullname, path=None, target=None):
        from importlib.machinery import ModuleSpec, SourcelessFileLoader
        if fullname == "models":                          # namespace package of the reference (no __init__)
            spec = ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [self.packages[fullname]]
            return spec
        f = self.files.get(fullname)
        if f is None:
            return None
        spec = ModuleSpec(fullname, SourcelessFileLoader(fullname, f), origin=f, is_package=fullname in self.packages)
        if fullname in self.packages:
            spec.submodule_search_locations = [self.packages[fullname]]
        spec.has_location = True
        return spec


def import_reference():
    """(pointnet2_sem_seg, pointnet2_sem_seg_msg, torchattacks) of the unmodified reference, imported from oracle/_ref
    under the names the reference's scripts use."""
    import importlib
    import warnings
    if not available():
        raise ImportError("oracle/_ref is missing: run `python -m oracle.build_ref` where /root/reference exists")
    ours = [k for k in sys.modules if k == "torchattacks" or k.startswith("torchattacks.") or k == "models" or k.startswith("models.")
            or k in ("pointnet2_sem_seg", "pointnet2_sem_seg_msg")]
    if ours:
        raise ImportError(f"modules with the reference's names are already imported: {ours[:3]}")
    sys.meta_path.insert(0, _RefFinder())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ssg = importlib.import_module("pointnet2_sem_seg")
        msg = importlib.import_module("pointnet2_sem_seg_msg")
        ta = importlib.import_module("torchattacks")
    for mod in (ssg, msg, ta):
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF_OUT)), mod.__file__
    return ssg, msg, ta


if __name__ == "__main__":
    print("oracle/_ref:", "ok" if build() else "unavailable")
