# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference package under ``oracle/_ref/``.

    python oracle/build_ref.py            # build container only: needs /root/reference

fidelity/stoke is pure Python, so "building" it is a verbatim copy of its package directory
(``/root/reference/stoke`` -> ``oracle/_ref/stoke``).  ``oracle/_ref/`` is git-ignored (reference sources never enter this
repository's history) but it is NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot, where
``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time the reference's own CPU path
(``cpu_baseline.kind == "reference"``) through ``oracle/ref_shim.py``.  When the copy is absent those legs fall back to
``oracle/stoke_port.py`` (``kind == "port"``), the restatement that the tests pin bit-for-bit against this same package.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("STOKE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def _make_writable(root: str):
    for d, _, files in os.walk(root):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)


def build(verbose: bool = True) -> bool:
    src_pkg = os.path.join(SRC, "stoke")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"build_ref: {src_pkg} not present (GPU box?) -- keeping whatever is under {DST}")
        return os.path.isdir(os.path.join(DST, "stoke"))
    dst_pkg = os.path.join(DST, "stoke")
    if os.path.isdir(dst_pkg):
        _make_writable(dst_pkg)
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    _make_writable(dst_pkg)   # the source tree is mounted read-only; the copy must stay removable
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copy of /root/reference/stoke made by oracle/build_ref.py; git-ignored, test infrastructure only.\n")
    if verbose:
        n = sum(len(files) for _, _, files in os.walk(dst_pkg))
        print(f"build_ref: copied {n} files to {dst_pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
