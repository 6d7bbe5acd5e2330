"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference modules for oracle validation.

The reference (/root/reference, read-only, Python 2 source) cannot be imported under
CPython 3 as-is.  This loader reads the three hot-path files *where they lie*, applies the
mechanical py3 patch described in SURVEY.md section 8c in memory (nothing is copied into
the repo), and exec()s the result into fresh module objects:

  * ``print "..."`` statements -> ``print("...")``
        extension_correction.py:25,314,321,399 ; kmers_for_component.py:26
  * ``len(kmers.keys()[0])`` -> ``len(next(iter(kmers)))``   extension_correction.py:220

It only works in the build container (the GPU box has no /root/reference); callers must
check :func:`available` first.  Used by ``tests/golden/make_golden.py`` (fixture generation)
and ``tests/test_oracle_vs_reference.py`` (pins ``oracle/shannon_oracle.py`` to the reference).
Nothing in the product (``shannon_b200/``) may import this file.
"""
import os
import re
import sys
import types

REFERENCE_DIR = os.environ.get("SHANNON_REFERENCE_DIR", "/root/reference")

_PRINT_STMT = re.compile(r'^(\s*)print (".*)$', re.M)
_PRINT_STMT2 = re.compile(r"^(\s*)print ('.*)$", re.M)       # mbgraph.py:149


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "extension_correction.py"))


def _patched_source(name):
    with open(os.path.join(REFERENCE_DIR, name + ".py")) as f:
        src = f.read()
    src = src.expandtabs(8)      # Python 2 reads a tab as "to the next multiple of 8" (faster_reps.py, mbgraph.py)
    src = _PRINT_STMT.sub(lambda m: "%sprint(%s)" % (m.group(1), m.group(2)), src)
    src = _PRINT_STMT2.sub(lambda m: "%sprint(%s)" % (m.group(1), m.group(2)), src)
    src = src.replace("len(kmers.keys()[0])", "len(next(iter(kmers)))")
    src = src.replace("from sets import Set\n", "")          # faster_reps.py:6 (unused there)
    return src


def load(name):
    """Return a fresh module object for reference module ``name`` (fresh globals every
    call: the reference keeps ``rmer_to_contig`` / ``cmer_to_contig`` as module globals,
    extension_correction.py:12-14, so one module object is good for exactly one run)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    mod = types.ModuleType("_shannon_ref_" + name)
    mod.__file__ = os.path.join(REFERENCE_DIR, name + ".py")
    if name == "multibridging":
        # it does `from mbgraph import *`
        sys.modules["mbgraph"] = load("mbgraph")
    if name == "kmers_for_component":
        # it does `from weight_updated_graph import weight_updated_graph`
        wug = load("weight_updated_graph")
        sys.modules["weight_updated_graph"] = wug
    code = compile(_patched_source(name), mod.__file__, "exec")
    exec(code, mod.__dict__)
    return mod
