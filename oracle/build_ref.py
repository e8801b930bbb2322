"""TEST / BENCH INFRASTRUCTURE ONLY -- populates ``oracle/_ref/`` with the reference's own hot-path files.

The reference (alondj/Mesh_R-CNN_Computer_Vision_project) is pure Python: there is nothing to compile.  For the
reference arm of ``bench.py`` (``--impl reference``) to time the *reference's own* CPU implementation on the GPU box
-- where ``/root/reference`` does not exist -- the six files of the hot path are copied, byte for byte, from where they
lie under ``/root/reference`` into ``oracle/_ref/`` (git-ignored, so never part of the history; not gpurun-ignored, so
it travels with the working tree like a built ``.so``).  ``__graft_entry__.build()`` runs this in the dev container.

    python -m oracle.build_ref              # dev container only; a no-op when /root/reference is absent

Files: meshRCNN/layers.py, meshRCNN/loss_functions.py, meshRCNN/utils.py,
utils/mesh_sampling.py, utils/process.py, utils/rotation.py; plus meshRCNN/shapenet_model.py for the drop-in test.  They are imported behind the shims of
``oracle/ref_import.py`` (stub packages, symeig -> eigh, stable argsort); nothing in them is edited.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MESHRCNN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("meshRCNN/layers.py", "meshRCNN/loss_functions.py", "meshRCNN/utils.py", "utils/mesh_sampling.py",
         "utils/process.py", "utils/rotation.py",
         # the reference's own model file: tests/test_dropin_gpu.py executes it UNMODIFIED with `.layers` / `.loss_functions`
         # bound to this repo's modules (the drop-in claim of INTEGRATION.md, checked on the GPU box)
         "meshRCNN/shapenet_model.py")


def build_ref(verbose: bool = False) -> bool:
    """Copies the hot-path files; returns True when ``oracle/_ref`` is complete afterwards."""
    if os.path.isdir(os.path.join(SRC, "meshRCNN")):
        for rel in FILES:
            src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                if verbose:
                    print("oracle/_ref/%s" % rel)
    return complete()


def complete() -> bool:
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    print("oracle/_ref %s" % ("complete" if ok else "NOT available (no reference tree here)"))
    sys.exit(0)
