"""TEST / BENCH INFRASTRUCTURE ONLY — recipe that makes the *real* reference hot path travel to the GPU box.

The reference is pure Python, so "compiling the reference from its own sources" means byte-compiling: the six
modules of ZeningLin/PEneo that make up the hot path (and nothing else of it) are compiled with ``py_compile``
from where they lie under ``/root/reference`` into ``oracle/_ref/<package>/<module>.refbc`` (a regular ``.pyc`` image under a
neutral extension: the gpurun snapshot drops ``*.pyc`` files).  ``oracle/_ref/`` is
git-ignored (no reference source or binary ever enters the history) but not gpurun-ignored, so the byte code
travels with the repository snapshot like our own built ``.so``; ``ref_shim.load_reference()`` imports it through
a sourceless loader on a box that has no ``/root/reference``.  No reference source text is copied.

    python oracle/build_ref.py            # called by __graft_entry__.build() when /root/reference exists

Consumers: ``bench.py --impl reference`` and the ``cpu_baseline`` leg (they time the reference's own
``PEneoDecoder.forward`` + ``sample_decode_peneo`` on the host cores, ``kind: "reference"``), and the tests that
pin the oracle.  Nothing under ``peneo_b200/`` may import it.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("PENEO_REFERENCE_ROOT", "/root/reference")

# (package, module): the hot path and what it imports
MODULES = (
    ("model", "configuration_peneo"),
    ("model", "peneo_decoder"),
    ("model", "custom_loss"),
    ("pipeline", "decode"),
    ("pipeline", "evaluation"),
    ("data", "data_utils"),
)


def build(quiet: bool = False) -> bool:
    """Byte-compile the reference modules into oracle/_ref.  Returns False when the reference tree is absent."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "peneo_decoder.py")):
        return False
    for pkg, mod in MODULES:
        src = os.path.join(REFERENCE_ROOT, pkg, mod + ".py")
        dst = os.path.join(OUT, pkg, mod + ".refbc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path recorded in tracebacks points at the reference tree, not at a file of this repository
        py_compile.compile(src, cfile=dst, dfile=f"<reference>/{pkg}/{mod}.py", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    if not quiet:
        print(f"compiled {len(MODULES)} reference modules into {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
