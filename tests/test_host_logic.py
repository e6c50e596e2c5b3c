"""Host logic of the drop-in without a GPU: edge_fuse_b200/csrc/cachemap_api.c (write-behind ring and
flusher, the combining queue of single-page gets with per-request completion, range calls,
counters) compiled against a CPU stand-in of the engine (tests/c/mock_engine.c) and hammered from
many threads by tests/c/host_stress.c — once plain, once under ThreadSanitizer.  Test
infrastructure only: nothing of the product links the stand-in."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "edge_fuse_b200", "csrc", "cachemap_api.c"),
       os.path.join(ROOT, "tests", "c", "mock_engine.c"),
       os.path.join(ROOT, "tests", "c", "host_stress.c")]


def _build(tmp_path, name, extra):
    exe = str(tmp_path / name)
    r = subprocess.run(["gcc", "-std=gnu11", "-O1", "-g", "-pthread", *extra, *SRC, "-o", exe], capture_output=True, text=True)
    return exe if r.returncode == 0 else None, r.stderr


def _run(exe, threads, ops, pshift, limit, *mode, **extra_env):
    d = tempfile.mkdtemp()
    try:
        env = dict(os.environ, CMB200_PERSIST="0", TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0", **extra_env)
        return subprocess.run([exe, d, str(threads), str(ops), str(pshift), str(limit), *mode], capture_output=True, text=True,
                              timeout=limit + 30, env=env)
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("threads,pshift", [(1, 12), (16, 12), (48, 12), (8, 16)])
def test_host_layer_under_many_callers(tmp_path, threads, pshift):
    """Read-your-writes through the write-behind ring, whole pages only, no lost request (every call
    returns), counters that add up — with 1 to 48 caller threads, 4 KiB and 64 KiB pages."""
    exe, err = _build(tmp_path, "host_stress", [])
    assert exe, err
    out = _run(exe, threads, 1500 if pshift == 12 else 400, pshift, 150)
    assert out.returncode == 0 and "host_stress ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("slots,pshift,threads", [("64", 12, 24), ("64", 16, 12), ("0", 12, 12), ("256", 17, 8)])
def test_host_layer_ring_shapes_and_large_pages(tmp_path, slots, pshift, threads):
    """A write-behind ring of 64 slots (wrap-around and back-pressure all the time), no ring at all
    (synchronous puts), and 128 KiB pages, which the fused single-page get does not serve: the queue
    then answers its batches through the synchronous two-kernel call."""
    exe, err = _build(tmp_path, "host_stress", [])
    assert exe, err
    out = _run(exe, threads, 1200 if pshift < 16 else 300, pshift, 150, CMB200_WB_SLOTS=slots)
    assert out.returncode == 0 and "host_stress ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("threads", [1, 16])
def test_host_layer_keeps_capacity_under_eviction(tmp_path, threads):
    """Ten times more keys than the capacity: the flusher evicts before every batch
    (cachemap.c:17-45 as filemap_evict), pages that come back are whole and their key's, and the store
    ends at its capacity."""
    exe, err = _build(tmp_path, "host_stress", [])
    assert exe, err
    out = _run(exe, threads, 4000, 12, 150, "evict")
    assert out.returncode == 0 and "host_stress ok" in out.stdout, out.stdout + out.stderr


def test_host_layer_has_no_data_race(tmp_path):
    """The same run under ThreadSanitizer: the queue's lock-free parts (watching for a launch, for a
    free slot, for one's own answer; the last one out ending the launch) must be race-free."""
    exe, err = _build(tmp_path, "host_stress_tsan", ["-fsanitize=thread"])
    if not exe:
        pytest.skip("gcc cannot link -fsanitize=thread here: " + err[-200:])
    out = _run(exe, 12, 600, 12, 400)
    assert out.returncode == 0 and "host_stress ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
