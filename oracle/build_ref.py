"""Recipe for oracle/_ref/ (TEST INFRASTRUCTURE; outputs are git-ignored and never committed).

The reference's hot path lives in un-vendored packages, so there is nothing of it to compile into a CPU baseline.  What the
reference DOES hold of the path is its Python glue -- scripts/train.py, util/rlkit_utils.py, util/rlkit_custom.py, util/arguments.py: the
callers of the drop-in boundary.  This recipe byte-compiles those four files FROM WHERE THEY LIE under /root/reference into
the archive oracle/_ref/refpy.bin (a zip of `.pyc` members only -- a build product, like the `.so` a C reference would give; no source is
copied), so that the GPU box, where /root/reference does not exist, can run the reference's own entry point unmodified against this package
(tests/test_gpu_reference_dropin.py, PYTHONPATH = robosuite_benchmark_b200/compat : repo : oracle/_ref/refpy.bin -- zipimport loads the
sourceless modules; the archive is not named .zip/.pyc because the box snapshot skips byte-code files).

  python oracle/build_ref.py        (also run by __graft_entry__.build() whenever /root/reference is present)
"""
import os
import py_compile
import sys
import tempfile
import zipfile

REF = os.environ.get("RSB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "refpy.bin")
FILES = ["scripts/train.py", "util/rlkit_utils.py", "util/rlkit_custom.py", "util/arguments.py"]


def build(verbose=False):
    if not os.path.isdir(REF):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(OUT + ".tmp", "w", zipfile.ZIP_STORED) as z:
        for d in sorted({os.path.dirname(f) for f in FILES}):
            z.writestr(zipfile.ZipInfo(d + "/"), "")                   # directory entries: `util` and `scripts` are namespace packages
        for rel in FILES:
            dst = os.path.join(tmp, os.path.basename(rel) + "c")
            py_compile.compile(os.path.join(REF, rel), cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True)
            z.write(dst, rel + "c")
            if verbose:
                print("compiled", rel, "->", os.path.relpath(OUT, HERE) + ":" + rel + "c")
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    out = build(verbose=True)
    print(out or f"{REF} not present: nothing built")
    sys.exit(0)
