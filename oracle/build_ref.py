#!/usr/bin/env python
"""Build the UNMODIFIED reference (isuthermography/heatsim2) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under heatsim2_b200/ imports this or its
output; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may use oracle/_ref.

What it does (the reference's own setup.py cannot run here: it imports the
removed ``numpy.distutils`` (setup.py:10) and its checked-in ``*.c`` files are
Cython-0.29 output that does not compile against CPython 3.12):

  1. cython -3 on the three ``.pyx`` files *where they lie* under the
     reference tree (read-only); generated C goes to oracle/_ref/build/.
  2. gcc -O2 on the generated C (+ the reference's own
     heatsim2/alternatingdirection_c.c for the ADI container) -> extension
     modules in oracle/_ref/heatsim2/.
  3. "installs" the package's pure-python modules next to them, exactly like
     ``pip install --target`` would (oracle/_ref/ is git-ignored: no reference
     source ever enters the repository history).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
Exit code 0 and prints the target dir on success; exit 3 when the reference
tree is absent (GPU box: the prebuilt oracle/_ref travels with the snapshot).
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")

PYX = ["tridiag", "crank_nicolson", "alternatingdirection_c_pyx"]
EXTRA_C = {"alternatingdirection_c_pyx": ["alternatingdirection_c.c"]}


def have_ref():
    """True when a previously built reference is importable from oracle/_ref."""
    pkg = os.path.join(REF_OUT, "heatsim2")
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(pkg, m + suffix)) for m in PYX) and \
        os.path.exists(os.path.join(pkg, "__init__.py"))


def build(reference="/root/reference", force=False, verbose=True):
    src = os.path.join(reference, "heatsim2")
    if have_ref() and not force:
        return REF_OUT
    if not os.path.isdir(src):
        return None
    import numpy as np
    from Cython.Compiler.Main import compile as cy_compile, CompilationOptions, default_options

    pkg = os.path.join(REF_OUT, "heatsim2")
    bld = os.path.join(REF_OUT, "build")
    os.makedirs(pkg, exist_ok=True)
    os.makedirs(bld, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pyinc = sysconfig.get_paths()["include"]
    for mod in PYX:
        cfile = os.path.join(bld, mod + ".c")
        opts = CompilationOptions(default_options, output_file=cfile,
                                  language_level=3,
                                  include_path=[src])
        res = cy_compile([os.path.join(src, mod + ".pyx")], opts,
                         full_module_name="heatsim2." + mod)
        if res.num_errors:
            raise RuntimeError("cython failed on %s" % mod)
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w",
               "-DNPY_NO_DEPRECATED_API=0",
               "-I", pyinc, "-I", np.get_include(), "-I", src, cfile]
        cmd += [os.path.join(src, c) for c in EXTRA_C.get(mod, [])]
        cmd += ["-o", os.path.join(pkg, mod + suffix)]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    for fn in sorted(os.listdir(src)):
        if fn.endswith(".py"):
            shutil.copyfile(os.path.join(src, fn), os.path.join(pkg, fn))
    return REF_OUT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    out = build(a.reference, a.force)
    if out is None:
        print("reference tree not found and no prebuilt oracle/_ref", file=sys.stderr)
        sys.exit(3)
    print(out)


if __name__ == "__main__":
    main()
