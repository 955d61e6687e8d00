"""The helper threads behind speechPlayer_synthesizeBatch's host-side work (nvspeechplayer_b200/csrc/host_pool.h), built on the
CPU by tests/hostsim: parallelFor visits every index exactly once for any (n, grain), thousands of times in a row, with and
without helpers."""
import ctypes
import os
import subprocess
import sys

from tests.hostsim import sim


def test_parallel_for_visits_every_index_once():
    L = sim.lib()
    L.hostsim_host_pool.restype = ctypes.c_int
    L.hostsim_host_pool.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint)]
    helpers = ctypes.c_uint(0)
    assert L.hostsim_host_pool(5000, 300, ctypes.byref(helpers)) == 0
    assert helpers.value <= 64


def test_no_helpers_is_the_serial_loop():
    code = ("import ctypes; from tests.hostsim import sim; L = sim.lib(); h = ctypes.c_uint(9); "
            "rc = L.hostsim_host_pool(200, 50, ctypes.byref(h)); assert rc == 0 and h.value == 0, (rc, h.value)")
    env = dict(os.environ, NVSP_HOST_THREADS="0")
    subprocess.check_call([sys.executable, "-c", code], env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_forked_child_runs_serially():
    """A fork()ed child inherits the pool object without its threads: parallelFor must fall back to the plain loop, not wait for
    helpers that do not exist."""
    code = ("import ctypes, os; from tests.hostsim import sim; L = sim.lib(); h = ctypes.c_uint(0); "
            "assert L.hostsim_host_pool(50, 300, ctypes.byref(h)) == 0; "
            "pid = os.fork(); "
            "rc = L.hostsim_host_pool(50, 300, ctypes.byref(h)); "
            "os._exit(0 if rc == 0 else 3) if pid == 0 else None; "
            "_, st = os.waitpid(pid, 0); assert os.WIFEXITED(st) and os.WEXITSTATUS(st) == 0, st")
    subprocess.run([sys.executable, "-c", code], check=True, timeout=60,
                   cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
