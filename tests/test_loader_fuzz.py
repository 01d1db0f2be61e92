"""Loader robustness (out-of-bounds reads on malformed glTF): seeded mutation fuzz of the glTF and text-scene loaders
and of the BVH builders behind them, on HOST-ONLY scenes (no GPU).  Every mutated file must either load or be rejected with an
RT_ERR_* code -- the child process must not crash or hang.  Runs in a subprocess so that a segfault is a test failure, not a dead pytest."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

FUZZ = r'''
import os, random, re, shutil, sys, tempfile
sys.path.insert(0, %(root)r)
import rtb200 as rt
random.seed(%(seed)d)
S = os.path.join(%(root)r, "scenes")
d = tempfile.mkdtemp()
ok = rejected = 0
def attempt(load, path):
    global ok, rejected
    try:
        sc = load(path); sc.info(); sc.bvh(); sc.close(); ok += 1
    except rt.RtError:
        rejected += 1
# 1. glTF: byte-level damage (mostly invalid JSON) and number-level damage (valid JSON, hostile counts / offsets / indices / coordinates)
for name in ("practice7_1", "practice7_4"):
    text = open(os.path.join(S, name + ".gltf")).read()
    shutil.copy(os.path.join(S, name + ".bin"), d)
    nums = [m.span() for m in re.finditer(r"(?<=[:\[,\s])-?\d+(\.\d+)?(e-?\d+)?", text)]
    for it in range(%(n)d):
        if it %% 3 == 0:
            b = bytearray(text.encode())
            for _ in range(random.randint(1, 6)):
                k = random.randrange(len(b))
                if random.random() < 0.5: b[k] = random.choice(b'0123456789-e.,:[]{}" ')
                else: del b[k:k + random.randint(1, 8)]
            data = bytes(b)
        else:
            t = text
            for (a, e) in sorted(random.sample(nums, random.randint(1, 4)), reverse=True):
                t = t[:a] + random.choice(["0", "-1", "1", "3", "255", "65536", "4294967295", "-2147483649", "1e30", "-1e30", "0.5", "123456789012", str(random.randint(-10, 10 ** 9))]) + t[e:]
            data = t.encode()
        p = os.path.join(d, "f.gltf"); open(p, "wb").write(data)
        attempt(lambda q: rt.Scene.from_gltf(q, 8, 8, 1, device=-1), p)
# 2. text scenes: line-level damage
for name in ("practice3_5", "practice3_4"):
    lines = open(os.path.join(S, name + ".txt")).read().split("\n")
    for it in range(%(n)d):
        L = list(lines)
        for _ in range(random.randint(1, 5)):
            k = random.randrange(len(L)); r = random.random()
            if r < 0.3: L[k] = re.sub(r"-?\d+(\.\d+)?", lambda m: random.choice(["0", "-1", "1e30", "nan", "inf", "999999999999", "abc", ""]), L[k], count=1)
            elif r < 0.5: del L[k]
            elif r < 0.7: L.insert(k, random.choice(L))
            elif r < 0.85: L[k] = L[k][:random.randrange(len(L[k]) + 1)]
            else: L[k] = L[k] + " " + random.choice(["1", "x", "-3 4", "NEW_PRIMITIVE", "1e400"])
        p = os.path.join(d, "f.txt"); open(p, "w").write("\n".join(L))
        attempt(lambda q: rt.Scene.from_text(q, 8, 8, 1, device=-1), p)
print("FUZZ ok %%d rejected %%d" %% (ok, rejected))
'''


@pytest.mark.parametrize("seed", [1, 2])
def test_mutated_scene_files_never_crash_the_loaders(rt, seed):
    r = subprocess.run([sys.executable, "-c", FUZZ % {"root": ROOT, "seed": seed, "n": 60}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("FUZZ ")][-1].split()
    ok, rejected = int(line[2]), int(line[4])
    assert ok + rejected == 4 * 60 and ok > 20 and rejected > 20, line     # both outcomes occur: the fuzz reaches past the parser
