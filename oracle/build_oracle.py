"""oracle/build_oracle.py -- TEST INFRASTRUCTURE ONLY.

Builds the two CPU checkers used by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py:

* oracle/liboracle.so        -- our own C restatement (oracle/crick_oracle.c)
* oracle/_ref/libcrick_ref.so -- the UNMODIFIED reference C sources, compiled from
  where they lie under /root/reference through oracle/ref_harness.c.  Only
  possible where /root/reference exists (this container); the built .so is
  git-ignored but travels to the GPU box with the gpurun snapshot.

Compile flags for the reference follow its own build (setup.py:12-23, observed
compile line: gcc -fno-strict-overflow -DNDEBUG -g -O2 -fPIC, no -march => no FMA).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CRICK_REFERENCE_DIR", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")

REF_CFLAGS = ["-fno-strict-overflow", "-DNDEBUG", "-g", "-O2", "-fPIC", "-w"]


def _run(cmd):
    subprocess.run(cmd, check=True)


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.stat(target).st_mtime
    return all(os.path.exists(s) and os.stat(s).st_mtime <= t for s in sources)


def build_restatement(force=False):
    src = os.path.join(HERE, "crick_oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if not force and _newer(out, [src]):
        return out
    # -ffp-contract=off: the reference build has no FMA contraction
    _run(["gcc", "-O2", "-fno-strict-overflow", "-ffp-contract=off", "-fPIC", "-shared",
          "-Wall", "-o", out, src, "-lm", "-lpthread"])
    return out


def build_reference(force=False):
    """Returns the path of libcrick_ref.so, or None when it can neither be
    built (no /root/reference) nor found prebuilt."""
    out = os.path.join(REF_OUT, "libcrick_ref.so")
    harness = os.path.join(HERE, "ref_harness.c")
    ref_src = os.path.join(REF, "crick")
    if not os.path.isdir(ref_src):
        return out if os.path.exists(out) else None
    deps = [harness, os.path.join(ref_src, "tdigest_stubs.c"),
            os.path.join(ref_src, "stats_stubs.c"),
            os.path.join(ref_src, "space_saving_stubs.c.in")]
    if not force and _newer(out, deps):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    # Tempita expansion exactly as the reference's setup.py:26-51 does it; the
    # generated file stays in oracle/_ref/ (git-ignored).
    from Cython import Tempita as tempita
    with open(os.path.join(ref_src, "space_saving_stubs.c.in")) as f:
        expanded = tempita.sub(f.read())
    with open(os.path.join(REF_OUT, "space_saving_stubs.c"), "w") as f:
        f.write(expanded)
    import numpy as np
    inc = ["-I" + REF_OUT, "-I" + ref_src, "-I" + os.path.join(ref_src, "klib"),
           "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()]
    _run(["gcc"] + REF_CFLAGS + inc + ["-shared", "-fvisibility=hidden", "-o", out, harness,
                                       "-lm", "-lpthread"])
    return out


def stage_reference_tests():
    """The reference's own pytest files (crick/tests/*.py) are the parity contract of the drop-in
    package: tests/test_reference_suite.py runs them UNMODIFIED against the
    `crick` shim.  /root/reference does not exist on the GPU box, so they are staged -- like the
    built reference .so -- under oracle/_ref/ (git-ignored: nothing of the reference enters the
    repository's history; the directory travels with the gpurun snapshot).  Returns the staged
    directory, or None when there is neither a reference tree nor a staged copy."""
    import shutil
    src = os.path.join(REF, "crick", "tests")
    dst = os.path.join(REF_OUT, "ref_tests")
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else None
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if name.startswith("test_") and name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    return dst


def main():
    force = "--force" in sys.argv
    print("restatement:", build_restatement(force))
    print("reference:  ", build_reference(force))
    print("ref tests:  ", stage_reference_tests())


if __name__ == "__main__":
    main()
