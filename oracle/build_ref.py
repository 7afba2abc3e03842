"""Stage the UNMODIFIED reference as a byte-compiled module under oracle/_ref/.

    python oracle/build_ref.py     # needs /root/reference/RBDReference.py (build container only)

The reference is one pure-Python file, so "building" it means byte-compiling it: the source
stays where it lies under /root/reference, only the marshalled code object
`RBDReference.codeobj` lands in oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it travels
to the GPU box like our own built .so files; a plain .pyc would be dropped by the snapshot).  `load_reference()` imports that sourceless module; `bench.py`
uses it for `cpu_baseline.kind == "reference"` and tests use it as a second checker
when present.  Nothing in the product package touches it.  TEST INFRASTRUCTURE.
"""
import marshal
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/RBDReference.py"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_PYC = os.path.join(OUT_DIR, "RBDReference.codeobj")


def build(verbose=True):
    """Byte-compile the reference into oracle/_ref/.  Returns True if the .pyc is (now) there."""
    if not os.path.exists(REF_SRC):
        if verbose:
            print("oracle/build_ref: %s not present (GPU box?) - using prebuilt %s: %s"
                  % (REF_SRC, OUT_PYC, os.path.exists(OUT_PYC)))
        return os.path.exists(OUT_PYC)
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(REF_SRC, "rb") as fh:
        code = compile(fh.read(), "RBDReference.py", "exec", dont_inherit=True, optimize=0)
    with open(OUT_PYC, "wb") as fh:
        fh.write(("%d.%d\n" % sys.version_info[:2]).encode())
        marshal.dump(code, fh)
    with open(os.path.join(OUT_DIR, "PROVENANCE.txt"), "w") as fh:
        fh.write("compile()+marshal of %s with python %s\n" % (REF_SRC, sys.version.split()[0]))
    if verbose:
        print("oracle/build_ref: wrote", OUT_PYC)
    return True


def load_reference():
    """Return the reference's RBDReference class from oracle/_ref, or None if not staged."""
    if not os.path.exists(OUT_PYC):
        return None
    try:
        with open(OUT_PYC, "rb") as fh:
            if fh.readline().decode().strip() != "%d.%d" % sys.version_info[:2]:
                return None
            code = marshal.load(fh)
        mod = types.ModuleType("_rbd_reference_staged")
        exec(code, mod.__dict__)
        return mod.RBDReference
    except Exception:  # stale magic number etc. -> treat as unavailable
        return None


if __name__ == "__main__":
    ok = build()
    cls = load_reference()
    print("staged reference importable:", cls is not None)
    sys.exit(0 if ok else 1)
