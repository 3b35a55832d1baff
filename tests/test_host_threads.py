"""The host half runs on several threads for large worlds (rtc_parallel.h: the flattener's subtree walks, the commit's
passes over the primitive arrays).  The result must not depend on how many: same arrays index for index, same commit
plan digest, same error for a bad input.  Each thread count is a fresh process (the pool is sized once, from
RTC_HOST_THREADS).  No GPU needed."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, hashlib, json, sys
sys.path.insert(0, %(root)r)
import ray_tracer_challenge_b200 as rt
from ray_tracer_challenge_b200 import RtcPrim, RtcNode, scenes
from ray_tracer_challenge_b200.api import Material, PointLight

api = rt.new_session()
out = {}

def digest_of(world, cam):
    counts = (C.c_int * 6)()
    api.check(api.lib.sg_flatten(api.ctx, world.handle, counts, None, None, None, None))
    prims = (RtcPrim * max(counts[0], 1))(); nodes = (RtcNode * max(counts[1], 1))()
    refs = (C.c_int32 * max(counts[2], 1))(); shapes = (C.c_int * max(counts[0], 1))()
    api.check(api.lib.sg_flatten(api.ctx, world.handle, counts, prims, nodes, refs, shapes))
    h = hashlib.sha256()
    for a in (prims, nodes, refs, shapes):
        h.update(bytes(a))
    info = api.inspect(cam, world)
    return {"arrays": h.hexdigest(), "counts": list(counts), "plan": {k: info[k] for k in sorted(info) if not k.endswith("_ms")}}

# 1. a divided sphere field with cylinders, cones, cubes and CSG (the c5 layout, smaller)
cam, world = scenes.stress(api, width=64, height=48, n_spheres=20000, n_each=16, n_csg=8)
out["stress"] = digest_of(world, cam)
# 2. a triangle mesh in divided groups (the c4 layout, smaller)
cam, world = scenes.dragon_element(api, width=64, height=48, n_u=96, n_v=48)
out["mesh"] = digest_of(world, cam)
# 3. a flat world: 12 000 top-level spheres, 300 distinct materials in an order that no thread sees from the start
objects = []
for i in range(12000):
    k = (i * 7919) %% 300
    m = Material(color=(k / 300.0, 0.5, 1.0 - k / 300.0), diffuse=0.7, reflective=0.1 if k %% 3 == 0 else 0.0)
    objects.append(api.Sphere.build(api.translation((i %% 100) * 3.0, (i // 100) * 3.0, 0.0), m))
world = api.World(objects, PointLight((-10, 10, -10), (1, 1, 1)))
cam = api.Camera(64, 48, 1.0, api.view_transform((150, 180, -400), (150, 180, 0), (0, 1, 0)))
out["flat"] = digest_of(world, cam)
prims, *_ = api.flatten(world)
out["flat_materials"] = [p.material for p in prims[:4000:37]]
print(json.dumps(out))
"""


def run_child(threads: int) -> dict:
    env = dict(os.environ, RTC_HOST_THREADS=str(threads))
    res = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def by_threads():
    return {t: run_child(t) for t in (1, 3, 8)}


@pytest.mark.parametrize("scene", ["stress", "mesh", "flat"])
def test_flattened_arrays_and_commit_plan_do_not_depend_on_the_thread_count(by_threads, scene):
    one = by_threads[1][scene]
    assert one["counts"][0] >= 9000  # large enough for the threaded path
    for t in (3, 8):
        assert by_threads[t][scene]["counts"] == one["counts"]
        assert by_threads[t][scene]["arrays"] == one["arrays"], f"{scene}: arrays differ between 1 and {t} threads"
        assert by_threads[t][scene]["plan"] == one["plan"], f"{scene}: commit plan differs between 1 and {t} threads"


def test_materials_are_numbered_by_first_use_whatever_the_split(by_threads):
    """300 materials met in a scrambled order by 12 000 top-level spheres: the threads' private tables are merged in
    depth-first order, so primitive i carries the index a sequential walk would have given it."""
    one = by_threads[1]["flat_materials"]
    assert len(set(one)) > 50
    first_use = {}
    expected = []
    for i in range(0, 4000):
        k = (i * 7919) % 300
        first_use.setdefault(k, len(first_use))
        if i % 37 == 0:
            expected.append(first_use[k])
    assert one == expected
    assert by_threads[3]["flat_materials"] == one and by_threads[8]["flat_materials"] == one


BAD_CHILD = r"""
import ctypes as C, sys
sys.path.insert(0, %(root)r)
import ray_tracer_challenge_b200 as rt
lib = C.CDLL(rt.LIB_DEVICE)
lib.rtc_last_error.restype = C.c_char_p
lib.rtc_set_camera.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
ident = (C.c_float * 16)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)
scene = C.c_void_p()
assert lib.rtc_scene_create(C.byref(scene)) == 0
assert lib.rtc_set_camera(scene, 64, 32, 1.0, 0.5, 2.0 / 64, ident) == 0
assert lib.rtc_set_point_light(scene, (C.c_float * 3)(-10, 10, -10), (C.c_float * 3)(1, 1, 1)) == 0
class RtcMaterial(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("v", C.c_float * 7), ("pattern", C.c_int32)]
mat = RtcMaterial()
mat.color[:], mat.v[:], mat.pattern = [1, 1, 1], [0.1, 0.9, 0.9, 200.0, 0.0, 0.0, 1.0], -1
assert lib.rtc_set_materials(scene, 1, C.byref(mat)) == 0
n = 100000
prims = (rt.RtcPrim * n)()
for i in range(n):
    p = prims[i]
    p.type, p.material, p.casts_shadow, p.parent = 0, 0, 1, -1
    p.inv[0] = p.inv[5] = p.inv[10] = p.inv[15] = 1.0
    p.inv[3] = -3.0 * i
    p.bbox_min[:] = [3.0 * i - 1, -1, -1]
    p.bbox_max[:] = [3.0 * i + 1, 1, 1]
prims[90001].material = 5
prims[70003].type = 42
prims[99999].parent = 7
lib.rtc_set_primitives.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(rt.RtcPrim)]
assert lib.rtc_set_primitives(scene, n, prims) == 0
info = rt.RtcCommitInfo()
rc = lib.rtc_scene_inspect(scene, C.byref(info))
print(rc, lib.rtc_last_error().decode())
prims[70003].type = 0
assert lib.rtc_set_primitives(scene, n, prims) == 0
rc = lib.rtc_scene_inspect(scene, C.byref(info))
print(rc, lib.rtc_last_error().decode())
"""


@pytest.mark.parametrize("threads", [1, 8])
def test_the_lowest_bad_primitive_is_the_one_reported(threads):
    env = dict(os.environ, RTC_HOST_THREADS=str(threads))
    res = subprocess.run([sys.executable, "-c", BAD_CHILD % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    first, second = res.stdout.strip().splitlines()[-2:]
    assert first.startswith("-1 ") and "primitive 70003: bad type" in first
    assert second.startswith("-1 ") and "primitive 90001: bad material index" in second


def test_worker_pool_survives_a_fork():
    """A forked child has none of the parent's worker threads: the pool must notice and start its own instead of
    waiting for workers that do not exist."""
    code = r"""
import os, sys
sys.path.insert(0, %(root)r)
import ray_tracer_challenge_b200 as rt
from ray_tracer_challenge_b200 import scenes
api = rt.new_session()
cam, world = scenes.stress(api, width=32, height=24, n_spheres=12000, n_each=4, n_csg=2)
before = api.inspect(cam, world)["digest"]          # starts the pool in the parent
pid = os.fork()
if pid == 0:
    ok = api.inspect(cam, world)["digest"] == before  # would hang without the pid check
    os._exit(0 if ok else 3)
_, status = os.waitpid(pid, 0)
print("child", os.WEXITSTATUS(status))
""" % {"root": ROOT}
    env = dict(os.environ, RTC_HOST_THREADS="4")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.strip().splitlines()[-1] == "child 0"


def test_flattener_stays_inside_its_arrays_under_address_sanitizer(tmp_path):
    """The threads write their subtrees' records at offsets computed from the subtree sizes, into arrays that are sized
    once and not initialised: the same driver built with -fsanitize=address,undefined must run clean."""
    import shutil

    import ray_tracer_challenge_b200 as rt

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path / "flatten_asan"
    pkg = os.path.dirname(rt.LIB_DEVICE)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fPIC", "-pthread",
           "-o", str(exe), os.path.join(ROOT, "tests", "tsan", "flatten_tsan.cpp"), os.path.join(pkg, "csrc", "host", "rtc_host.cpp"),
           "-L" + pkg, "-lrtc_b200", "-Wl,-rpath," + pkg]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "asan" in (build.stderr + build.stdout).lower():
        pytest.skip("AddressSanitizer runtime not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-2000:]
    env = dict(os.environ, RTC_HOST_THREADS="6", ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:exitcode=67")
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr
    assert "flattened: 26000 prims" in run.stdout


def test_flattener_is_race_free_under_thread_sanitizer(tmp_path):
    """The threaded flattener (subtree walks writing disjoint ranges, per-thread material tables, the worker pool's
    hand-off) compiled with -fsanitize=thread and run on a 26 000-primitive world: TSan must stay silent."""
    import shutil

    import ray_tracer_challenge_b200 as rt

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path / "flatten_tsan"
    pkg = os.path.dirname(rt.LIB_DEVICE)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-fPIC", "-pthread", "-o", str(exe),
           os.path.join(ROOT, "tests", "tsan", "flatten_tsan.cpp"), os.path.join(pkg, "csrc", "host", "rtc_host.cpp"),
           "-L" + pkg, "-lrtc_b200", "-Wl,-rpath," + pkg]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "tsan" in (build.stderr + build.stdout).lower():
        pytest.skip("ThreadSanitizer runtime not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-2000:]
    env = dict(os.environ, RTC_HOST_THREADS="6", TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "WARNING: ThreadSanitizer" not in run.stderr
    assert "flattened: 26000 prims" in run.stdout
