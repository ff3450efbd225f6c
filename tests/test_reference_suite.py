"""The reference's own test-suite, UNMODIFIED, against the drop-in package.

crick/tests/test_tdigest.py:72-444, test_stats.py:20-179, test_space_saving.py:23-378 and
test_version.py are what "drop-in" means operationally: here they
are copied to a scratch directory and run by pytest in a subprocess whose `import crick` resolves
to this repository's shim (crick/ -> crick_b200 -> libcrick_cuda.so).  The files come from
/root/reference where it exists (the build container) or from the copy oracle/build_oracle.py
staged under oracle/_ref/ref_tests/ (the GPU box).
"""
import os
import shutil
import subprocess
import sys
import xml.etree.ElementTree as ET

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Node ids that are allowed to fail, with the reason.  Empty: the object dtype -- the one hole the
# round-1 verdict listed -- is a host-side Stream-Summary now (crick_b200/_spsv_object.py).
EXPECTED_FAILURES = {}


def _reference_tests_dir():
    for d in (os.path.join(os.environ.get("CRICK_REFERENCE_DIR", "/root/reference"), "crick", "tests"),
              os.path.join(ROOT, "oracle", "_ref", "ref_tests")):
        if os.path.isfile(os.path.join(d, "test_tdigest.py")):
            return d
    return None


@pytest.mark.gpu
def test_reference_test_suite_passes_unmodified(tmp_path):
    src = _reference_tests_dir()
    assert src is not None, ("no copy of the reference's tests: run oracle/build_oracle.py where "
                             "/root/reference exists (it stages them under oracle/_ref/ref_tests)")
    dst = tmp_path / "ref_tests"
    dst.mkdir()
    names = sorted(n for n in os.listdir(src) if n.startswith("test_") and n.endswith(".py"))
    assert {"test_tdigest.py", "test_stats.py", "test_space_saving.py"} <= set(names)
    for n in names:
        shutil.copyfile(os.path.join(src, n), dst / n)
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    xml = tmp_path / "report.xml"
    # rootdir = the scratch directory, no conftest / ini of this repository in play
    proc = subprocess.run(
        [sys.executable, "-m", "pytest", str(dst), "-q", "-p", "no:cacheprovider", "-o", "addopts=",
         "--rootdir", str(dst), "-c", os.devnull, "--junitxml", str(xml)],
        cwd=str(dst), env=env, capture_output=True, text=True, timeout=1500)
    tail = (proc.stdout[-6000:] + "\n" + proc.stderr[-2000:])
    assert xml.exists(), "pytest did not run:\n" + tail
    # the shim, not some other crick, must have been what the tests imported
    probe = subprocess.run([sys.executable, "-c", "import crick, crick_b200; print(crick.__file__); "
                            "assert crick.TDigest is crick_b200.TDigest"], cwd=str(dst), env=env,
                           capture_output=True, text=True)
    assert probe.returncode == 0 and os.path.join(ROOT, "crick") in probe.stdout, probe.stdout + probe.stderr
    failed, passed, skipped = [], 0, 0
    for case in ET.parse(str(xml)).getroot().iter("testcase"):
        node = "%s::%s" % (case.get("classname", "").split(".")[-1] + ".py", case.get("name"))
        if case.find("failure") is not None or case.find("error") is not None:
            failed.append(node)
        elif case.find("skipped") is not None:
            skipped += 1
        else:
            passed += 1
    unexpected = [f for f in failed if f not in EXPECTED_FAILURES]
    missing = [f for f in EXPECTED_FAILURES if f not in failed]
    print("reference test-suite against the shim: %d passed, %d failed, %d skipped" % (passed, len(failed), skipped))
    assert not unexpected, "reference tests failing against the drop-in:\n%s\n%s" % ("\n".join(unexpected), tail)
    assert not missing, "expected failures that now pass (remove them from the list): %r" % missing
    assert passed >= 95, "only %d reference tests ran\n%s" % (passed, tail)
