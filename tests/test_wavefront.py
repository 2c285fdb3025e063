"""The wavefront ray queue (integration/src/render/wavefront_b200.cc) on the CPU: tests/native/fiber_harness.cc links it
against a stub of the libb200rt entry points it calls and checks the fiber switch, the batching and the grouping of mixed
scenes / shadow depths on several OS threads.  (The GPU end of the queue is covered by tests/test_render.py.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if os.uname().machine != "x86_64":
        pytest.skip("the fiber switch is x86-64 only")
    out = str(tmp_path_factory.mktemp("fiber") / "fiber_harness")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "integration", "include"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "native", "fiber_harness.cc"), os.path.join(ROOT, "integration", "src", "render", "wavefront_b200.cc"), "-o", out, "-lpthread"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    return out


@pytest.mark.parametrize("threads,fibers,jobs,depth,groups", [(1, 1, 50, 3, 1), (1, 8, 100, 3, 2), (4, 256, 20000, 40, 2), (8, 1024, 60000, 10, 3), (2, 64, 5000, 5, 1)])
def test_fiber_queue(harness, threads, fibers, jobs, depth, groups):
    p = subprocess.run([harness, str(threads), str(fibers), str(jobs), str(depth), str(groups)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.startswith("ok "), p.stdout[-2000:]
    rays, batches, calls = map(int, p.stdout.split()[1:4])
    assert rays > jobs and batches > 0 and calls >= batches


def test_photon_mutex_draws_each_tuple_once(tmp_path):
    """b200::PhotonMutex, the lock around SppmIntegrator's shared Halton sequences (integrator_sppm.cc:395-400): as a std::mutex
    and in spin mode (tuples drawn in batches per OS thread while a photon pass runs on fibers) every tuple comes from one step of
    the four sequences and is handed out at most once."""
    out = str(tmp_path / "photon_mutex_test")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "integration", "include"),
           os.path.join(ROOT, "tests", "native", "photon_mutex_test.cc"), "-o", out, "-lpthread"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    p = subprocess.run([out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.startswith("ok "), p.stdout
