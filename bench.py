#!/usr/bin/env python
"""bench.py — GiB/s through the cachemap put path (fingerprint -> LZ4 encode -> key-table insert)
on 64 KiB chunks, BASELINE.json's metric.

A step = one pass of the hot path over one batch = the whole 1 GiB synthetic stream of config 1
(16 384 x 64 KiB chunks, 0 % duplicates, classes R/T/Z/M round-robin) put into the
cache under fresh addresses (genid = step).  Per GPU the work is fixed (weak scaling): with N
ranks the global stream is N GiB and chunk k belongs to rank k mod N; after each step the ranks
all-gather their new key records over NCCL and import them into their index replica.

  value      device-timed (CUDA events on the engine's stream), pages already resident in HBM
  e2e        same metric through the C-ABI with page-locked HOST pages: H2D of every page and D2H
             of the per-chunk stored lengths inside the timed region.  Headline = the write-behind
             call cmb200_put_batch_async with two steps in flight; the strictly synchronous
             cmb200_put_batch figure is reported beside it (e2e.synchronous_call)
  roofline   the encode kernel alone: algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline  the reference's own CPU path (oracle/_ref, else the oracle port) on a bounded sample

`--impl reference` times the reference CPU implementation instead (same metric/config).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CHUNK = 65536
PSHIFT = 16
ACCEL = 12
SEED = 42
METRIC = "GiB/s hash+LZ4+dedup on 64 KiB chunks"
GIB = float(1 << 30)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def stream_for_rank(rank: int, world: int, n: int):
    """chunk ids of this rank's shard of the global stream (round-robin) + their addresses."""
    import edge_fuse_b200 as E
    cids = (np.arange(n, dtype=np.uint64) * np.uint64(world) + np.uint64(rank))
    off, nh = E.gen_addr(SEED, cids, PSHIFT)
    return cids, off, nh


# -------------------------------------------------------------------------------------------------
# reference / CPU arm
# -------------------------------------------------------------------------------------------------

def cpu_reference_run(pages: np.ndarray, off: np.ndarray, nh: np.ndarray, threads: int, reps: int = 1):
    """Times the reference's CPU path on `pages` ([n, 65536] host array).  Returns a dict with the
    full-path put/get rate (cachemap_put / cachemap_get on a tmpfs store) and the codec-only rate
    (LZ4_compress_fast / LZ4_decompress_fast, no LMDB), wall clock, `threads` pthreads."""
    import ctypes as C
    import tempfile
    from oracle import ef_oracle as O
    L = O.lib()
    R = O.ref()
    n = len(pages)
    out4 = (C.c_double * 4)()
    if R is not None:
        kind = "reference"
        enc = C.cast(R.LZ4_compress_fast, C.c_void_p)
        dec = C.cast(R.LZ4_decompress_fast, C.c_void_p)
    else:
        kind = "port"
        enc = C.cast(L.ef_port_compress_fast, C.c_void_p)
        dec = C.cast(L.ef_port_decompress_fast, C.c_void_p)
    best_enc, best_dec, comp_bytes = 1e30, 1e30, 0
    for _ in range(max(1, reps)):
        L.ef_cpu_bench_codec(enc, dec, pages.ctypes.data, n, CHUNK, ACCEL, threads, out4)
        assert out4[2] == 0, "CPU codec round trip mismatch"
        best_enc, best_dec, comp_bytes = min(best_enc, out4[0]), min(best_dec, out4[1]), out4[3]
    res = {"kind": kind, "cores": threads, "codec_encode_gibs": n * CHUNK / GIB / best_enc,
           "codec_decode_gibs": n * CHUNK / GIB / best_dec, "ratio": comp_bytes / (n * CHUNK)}
    if R is not None:
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        out3 = (C.c_double * 3)()
        offs = np.ascontiguousarray(off, dtype=np.uint64)
        nhs = np.ascontiguousarray(nh, dtype=np.uint64)
        best_put, best_get = 1e30, 1e30
        for _ in range(max(1, reps)):
            with tempfile.TemporaryDirectory(dir=base) as d:
                cm = R.cachemap_create(d.encode(), max(1024, 2 * n), ACCEL, PSHIFT)
                assert cm, "reference cachemap_create failed"
                L.ef_cpu_bench_store(C.cast(R.cachemap_put, C.c_void_p), C.cast(R.cachemap_get, C.c_void_p),
                                     C.c_void_p(cm), pages.ctypes.data, n, CHUNK, offs.ctypes.data,
                                     nhs.ctypes.data, threads, 1, out3)
                assert out3[2] == 0, "reference get returned different bytes"
                best_put, best_get = min(best_put, out3[0]), min(best_get, out3[1])
                # no cachemap_free(): it can hang in the reference; the LMDB files go
                # away with the temporary directory
        res["put_gibs"] = n * CHUNK / GIB / best_put
        res["get_gibs"] = n * CHUNK / GIB / best_get
    return res


def host_sample(n: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    import edge_fuse_b200 as E
    cids, off, nh = stream_for_rank(0, 1, n)
    pages = np.stack([E.gen_chunk_host(SEED, int(c), CHUNK) for c in cids])
    return pages, off, nh


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.cpu_sample_chunks
    pages, off, nh = host_sample(n)
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_reference_run(pages, off, nh, threads)
        if it >= args.warmup:
            times.append((time.perf_counter() - t0, res))
    # metric of a step = the put path of the reference: cachemap_put (fingerprint-less: the
    # reference has no content hash) when the reference compiled, else the codec-only port
    key = "put_gibs" if "put_gibs" in res else "codec_encode_gibs"
    vals = [r[key] for _, r in times]
    value = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": n * CHUNK / GIB / value * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.gpus, args.chunks),
        "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": threads, "kind": res["kind"],
                         "sample": f"first {n} chunks ({n * CHUNK >> 20} MiB) of the stream per step; "
                                   f"{'cachemap_put on a tmpfs LMDB store' if key == 'put_gibs' else 'LZ4 encode only'}, "
                                   f"{threads} threads",
                         "codec_encode_gibs": res["codec_encode_gibs"], "codec_decode_gibs": res["codec_decode_gibs"],
                         "get_gibs": res.get("get_gibs")},
        "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(gpus: int, chunks: int) -> dict:
    return {"workload": f"config 1: {chunks * CHUNK / GIB:g} GiB synthetic stream per GPU per step, 64 KiB fixed "
                        "chunks, 0% duplicates, classes R/T/Z/M round-robin, EF128 fingerprint + LZ4(accel 12) "
                        "encode + key-table insert",
            "chunk_bytes": CHUNK, "chunks_per_gpu_per_step": chunks, "pshift": PSHIFT, "accel": ACCEL,
            "sharding": f"chunk k -> rank k mod {gpus}" if gpus > 1 else "single GPU",
            "l2": "per-step input (1 GiB) is larger than the 126 MB L2; no explicit flush"}


# -------------------------------------------------------------------------------------------------
# CUDA arm
# -------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(torch, local: int):
    """Runs this rank on the CPUs next to its GPU (NVML's ideal affinity) so that the page-locked
    staging buffers it allocates are first-touched on that NUMA node; with 8 ranks on a two-socket
    host the H2D rate otherwise depends on where the allocator happened to put them.  Returns the
    previous affinity (restored before the CPU baseline, which uses every host thread)."""
    try:
        before = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return before
    except Exception:
        return None


def run_ours(args):
    import torch
    import edge_fuse_b200 as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert E.device_count() > local, f"no CUDA device for rank {rank}: {E.last_error()}"
    torch.cuda.set_device(local)
    affinity_before = bind_to_gpu_numa_node(torch, local)

    n = args.chunks
    total_steps = args.warmup + args.steps
    # 3 passes (device-resident, e2e synchronous, e2e write-behind) of total_steps fresh-address puts, ~0.51 stored bytes per input byte
    arena = int(3 * total_steps * n * CHUNK * 0.56) + (1 << 30)
    eng = E.Engine(pshift=PSHIFT, accel=ACCEL, capacity=6 * total_steps * n, arena_bytes=arena,
                   max_batch=args.max_batch, flags=E.FINGERPRINT, device=local)
    cids, off, nh = stream_for_rank(rank, world, n)
    d_pages = eng.dev_alloc(n * CHUNK)
    eng.gen_chunks_dev(SEED, cids, d_pages)
    h_ptr = E.lib().cmb200_host_alloc(n * CHUNK)
    assert h_ptr, "page-locked host buffer"
    h_pages = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (n * CHUNK)).from_address(h_ptr))
    eng.d2h(h_pages, d_pages)
    page_no = off >> np.uint64(PSHIFT)

    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    sync_all = (lambda: (dist.barrier(), torch.cuda.synchronize())) if dist else torch.cuda.synchronize

    from edge_fuse_b200 import sharding
    step_counter = [0]

    def begin_step():
        """Global stream positions of this step's chunks: chunk i of rank r is position
        base + r + world*i (round-robin sharding), which is also its last-writer-wins sequence."""
        base = 1 + step_counter[0] * world * n
        step_counter[0] += 1
        eng.set_stream_order(base + rank, world)
        return sharding.shard_positions(rank, world, n, base)

    def exchange(u, l, pos, lens):
        """multi-GPU: ONE all-gather (NCCL over NVLink) of this step's key records, then the
        other ranks' records go into the local index replica."""
        if not dist:
            return
        rec = torch.from_numpy(sharding.pack_records(u, l, pos, rank, lens)).cuda(non_blocking=True)
        gathered = sharding.all_gather_records(rec)
        sharding.import_gathered(eng, gathered, rank)

    def addr_for(step: int, lane: int):
        # fresh addresses every step: genid = step (low 20 bits kept, cachemap.c:163)
        l = page_no | (np.uint64(3 * step + lane) << np.uint64(44))
        return nh, l

    ts = np.full(n, 1, dtype=np.uint64)
    sampler = ClockSampler(local)

    # The step as the product runs it: cmb200_put_step (asynchronous; the per-chunk exchange
    # records are packed on the device), then for N > 1 ONE all-gather of those records over NCCL
    # on the engine's stream and cmb200_import_records_dev into the index replica: no host round
    # trip between the encode of one step and the next.
    dev = torch.device("cuda", local)
    rec = [torch.empty((n, 4), dtype=torch.int64, device=dev) for _ in range(2)]
    gath = [torch.empty((world * n, 4), dtype=torch.int64, device=dev) for _ in range(2)] if dist else None
    torch.cuda.synchronize()

    def submit_step(it: int, lane: int, pages, on_dev: bool, lens=None) -> int:
        u, l = addr_for(it, lane)
        begin_step()
        k = it & 1
        with torch.cuda.stream(stream):
            tk = eng.put_step(u, l, pages, ts=ts, on_dev=on_dev, rank=rank, records_dev=rec[k].data_ptr(), lens=lens)
            if dist:
                dist.all_gather_into_tensor(gath[k], rec[k])
                eng.import_records_dev(world * n, gath[k].data_ptr(), rank)
        return tk

    def lens_of(k: int) -> np.ndarray:
        tail = rec[k][:, 3].cpu().numpy()
        return ((tail & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000

    # ---- device-resident: value ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    st0 = None
    for it in range(total_steps):
        if it == args.warmup:
            sync_all()
            sampler.start()
            st0 = eng.stats()
            t_wall0 = time.perf_counter()
        if it >= args.warmup:
            ev[it - args.warmup][0].record(stream)
        submit_step(it, 0, d_pages, True)
        if it >= args.warmup:
            ev[it - args.warmup][1].record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    st1 = eng.stats()
    lens = lens_of((total_steps - 1) & 1)
    dev_ms = [a.elapsed_time(b) for a, b in ev]
    if dist:
        # multi-GPU step time includes the exchange: wall time between the barriers, max over ranks
        tt = torch.tensor([t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item()) / args.steps * 1e3
    else:
        step_ms = float(np.sum(dev_ms)) / args.steps
    value = world * n * CHUNK / GIB / (step_ms * 1e-3)

    # ---- e2e: host pages through the C ABI ----
    # (1) synchronous calls, one step at a time, each bracketed by a barrier + synchronize
    e2e_t = []
    for it in range(total_steps):
        u, l = addr_for(it, 1)
        sync_all()
        t0 = time.perf_counter()
        pos = begin_step()
        lens_e = eng.put(u, l, h_ptr, ts=ts, on_dev=False)
        exchange(u, l, pos, lens_e)
        sync_all()
        if it >= args.warmup:
            e2e_t.append(time.perf_counter() - t0)
    e2e_sync_s = float(np.mean(e2e_t))

    # (2) the write-behind call (cmb200_put_batch_async): step k+1 is submitted before step k's
    # result is read, so its host-to-device copy overlaps the tail of step k's encode.  Every step
    # still copies its own inputs from page-locked host memory and reads its own result (the
    # stored lengths) back inside the timed region; the region ends after the last result is in.
    lens_pin = []
    for _ in range(2):
        ptr = E.lib().cmb200_host_alloc(n * 4)
        assert ptr, "page-locked lens buffer"
        lens_pin.append((ptr, np.ctypeslib.as_array((np.ctypeslib.ctypes.c_int32 * n).from_address(ptr))))

    # two input buffers: step k+1's pages must not be the buffer step k is still being copied from
    h_ptr_b = E.lib().cmb200_host_alloc(n * CHUNK)
    assert h_ptr_b, "second page-locked host buffer"
    np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (n * CHUNK)).from_address(h_ptr_b))[:] = h_pages
    h_ptr2 = (h_ptr, h_ptr_b)

    def pipelined(first_step: int, count: int):
        inflight = None
        for k in range(count):
            # host pages stay untouched until the step's ticket is done (mode 2): the call does not
            # wait for its own copies, so the copy engine never idles between steps
            tk = submit_step(first_step + k, 2, h_ptr2[k & 1], 2, lens=lens_pin[k & 1][0])
            if inflight is not None:
                eng.wait(inflight[0])                   # step k-1's stored lengths are on the host
            inflight = (tk, lens_pin[k & 1][1])
        eng.wait(inflight[0])
        return inflight[1]

    pipelined(0, args.warmup)
    sync_all()
    t0 = time.perf_counter()
    lens_p = pipelined(args.warmup, args.steps)
    sync_all()
    e2e_s = (time.perf_counter() - t0) / args.steps
    assert (lens_p == lens_e).all(), "pipelined puts stored different lengths"
    clocks = sampler.stop()
    if dist:
        tt = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = float(tt[0].item()), float(tt[1].item())
    e2e = world * n * CHUNK / GIB / e2e_s
    e2e_sync = world * n * CHUNK / GIB / e2e_sync_s

    # ---- roofline of the dominant kernel (k_encode) ----
    peak, peak_src = peaks()
    enc_ns = st1["encode_kernel_ns"] - st0["encode_kernel_ns"]
    enc_launches = st1["encode_kernel_launches"] - st0["encode_kernel_launches"]
    stored = float(lens[lens > 0].sum())
    alg_bytes_step = n * (CHUNK + 24 + 64) + stored          # 65 624 + c per chunk
    achieved = alg_bytes_step * args.steps / (enc_ns * 1e-9) / 1e9 if enc_ns else 0.0
    roofline = {"bound": "hbm", "kernel": "k_encode (LZ4 encode + EF128 fingerprint along the parse + record in place + slot publish)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_encode launch of this workload
                # (16 384 chunks), from the ncu --set full capture summarised in
                # profiles/r1_encode_notes.md (1.357 GB read + 0.647 GB written); scaled by chunk count
                # if the launch size differs
                "traffic": 2.004139e9 * (n / max(1, enc_launches // args.steps)) / 16384.0,
                "traffic_source": "profiles/r1_encode_blend_ncu_details.txt (ncu --set full capture of this launch shape, round 1; not re-measured per run)",
                "algorithmic_bytes_per_launch": alg_bytes_step / max(1, enc_launches // args.steps),
                "avg_launch_ms": enc_ns / 1e6 / max(1, enc_launches),
                "read_form_frac": n * CHUNK * args.steps / (enc_ns * 1e-9) / 1e9 / peak if enc_ns else 0.0,
                "stored_ratio": stored / (n * CHUNK)}
    launches = (st1["kernel_launches"] - st0["kernel_launches"]) // args.steps
    final = eng.stats()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        pages_s = h_pages.reshape(n, CHUNK)[: args.cpu_sample_chunks]
        if affinity_before:
            os.sched_setaffinity(0, affinity_before)        # the CPU baseline gets every host thread
        r = cpu_reference_run(pages_s, off[: args.cpu_sample_chunks], nh[: args.cpu_sample_chunks],
                              os.cpu_count() or 1)
        key = "put_gibs" if "put_gibs" in r else "codec_encode_gibs"
        cpu = {"value": r[key], "unit": "GiB/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"first {len(pages_s)} chunks ({len(pages_s) * CHUNK >> 20} MiB) of the same stream; "
                         f"{'cachemap_put on a tmpfs LMDB store' if key == 'put_gibs' else 'LZ4 encode only'}",
               "codec_encode_gibs": r["codec_encode_gibs"], "codec_decode_gibs": r["codec_decode_gibs"],
               "get_gibs": r.get("get_gibs")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world, n),
            "e2e": {"value": e2e, "unit": "GiB/s", "h2d_bytes_per_step": int(n * (CHUNK + 16 + 8)),
                    "d2h_bytes_per_step": int(n * 4),
                    "call": "cmb200_put_step (write-behind) + cmb200_wait, 2 steps in flight from 2 page-locked input "
                            "buffers (step k+1 submitted before step k's stored lengths are read); all K steps, "
                            "copies and reads inside one timed region",
                    "synchronous_call": {"value": e2e_sync, "unit": "GiB/s",
                                         "call": "cmb200_put_batch, one step at a time, barrier + synchronize around each"}},
            "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "index": {"local_entries": final["entries"], "remote_entries": final["remote_entries"],
                      "exchange": "1 all-gather of 32-byte key records per step (NCCL on the engine's stream; records packed and imported on the device)" if dist else "none (single GPU)"},
            "parity_spot_check": spot_check(eng, E, h_pages.reshape(n, CHUNK), nh, page_no, total_steps),
        }
        print(json.dumps(line))
    eng.dev_free(d_pages)
    E.lib().cmb200_host_free(h_ptr)
    E.lib().cmb200_host_free(h_ptr_b)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def spot_check(eng, E, pages, nh, page_no, total_steps) -> str:
    """Every throughput number is gated on parity: re-read a few stored records
    of the last timed step and decode them back with the engine's own get path."""
    idx = np.arange(0, len(pages), max(1, len(pages) // 16))[:16]
    l = page_no[idx] | (np.uint64(3 * (total_steps - 1)) << np.uint64(44))
    out, status = eng.get(nh[idx], l)
    ok = bool((status == E.HIT).all() and (out == pages[idx]).all())
    return "ok: 16 sampled records of the last step decode back to their pages" if ok else "FAILED"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=16384, help="chunks per GPU per step (16384 = 1 GiB)")
    ap.add_argument("--max-batch", type=int, default=16384, help="chunks per kernel launch (resident pages)")
    ap.add_argument("--cpu-sample-chunks", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
