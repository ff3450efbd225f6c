"""tests/golden/make_golden_object.py -- regenerates tests/golden/golden_object.json.

BUILD container only (like make_golden.py): imports the built reference and records what its
``SpaceSaving(dtype=object)`` does on seeded streams of strings / ints / tuples, for the host-side
Stream-Summary of crick_b200/_spsv_object.py.

    python tests/golden/make_golden_object.py /tmp/refbuild
"""
import json
import os
import sys

import numpy as np

ref_dir = sys.argv[1] if len(sys.argv) > 1 else "/tmp/refbuild"
sys.path.insert(0, ref_dir)
import crick  # noqa: E402  (the reference)

assert os.path.dirname(crick.__file__).startswith(ref_dir)
HERE = os.path.dirname(os.path.abspath(__file__))


def items_of(kind, idx):
    if kind == "str":
        return ["k%d" % i for i in idx]
    if kind == "int":
        return [int(i) * 7 - 50 for i in idx]
    return [("t", int(i) % 5, int(i)) for i in idx]     # tuples (JSON: lists)


def obj_array(items):
    a = np.empty(len(items), dtype=object)
    for i, it in enumerate(items):
        a[i] = it
    return a


def dump(s):
    # snapshot through JSON text at once: the reference's merge can leave a key in its hash map
    # that it never INCREF'd (space_saving_stubs.c.in:334/345) and DECREFs it
    # when the summary is freed, so item objects may die while this script still holds them
    rows = [[list(i) if isinstance(i, tuple) else i, int(c), int(e)] for i, c, e in s.counters().tolist()]
    return json.loads(json.dumps(rows))


cases = []
KEEP = []   # nothing the reference touched is freed before the process ends (see dump())
rng = np.random.default_rng(11)
for kind in ("str", "int", "tuple"):
    for cap, ndistinct, n in ((10, 50, 400), (8, 6, 100), (5, 30, 64), (20, 200, 3000)):
        idx = np.minimum(rng.zipf(1.3, n), ndistinct) - 1
        cnt = rng.integers(1, 6, n)
        items = items_of(kind, idx)
        half = n // 2
        s = crick.SpaceSaving(cap, dtype=object)
        s.update(obj_array(items))
        a = crick.SpaceSaving(cap, dtype=object)
        a.update(obj_array(items[:half]), cnt[:half])
        b = crick.SpaceSaving(cap, dtype=object)
        b.update(obj_array(items[half:]), cnt[half:])
        rec = dict(kind=kind, capacity=cap, idx=idx.tolist(), cnt=cnt.tolist(), plain=dump(s),
                   a=dump(a), b=dump(b))
        m = crick.SpaceSaving(cap, dtype=object)
        m.merge(a, b)
        rec["merged"] = dump(m)
        a.merge(b)
        rec["a_merge_b"] = dump(a)
        cases.append(rec)
        KEEP.append((items, s, a, b, m))

with open(os.path.join(HERE, "golden_object.json"), "w") as f:
    json.dump(dict(cases=cases), f)
print(len(cases), "cases", flush=True)
os._exit(0)
