#!/usr/bin/env python
"""bench.py — decode + blake3-verify throughput of the znippy hot path on B200 (BASELINE.json metric).

A "step" is one pass of the read-side hot path (decompress.rs:148-184: decode-or-skip, blake3, 32-byte compare) over
one batch: BASELINE configs[1], a single 2 GiB text-pattern file = 256 index rows of 8 MiB slices
(stream_packer.rs:31), each a Zstandard level-19 frame (written by libzstd 1.5.5, standing in for the reference
writer), plus its blake3 `checksum` column.

  value     device-resident: blobs and outputs already in HBM, zn_plan_run only          (GB/s of uncompressed bytes)
  e2e       zn_decode_verify_batch with HOST buffers: H2D of the blobs and D2H of the decoded bytes inside the timing
  roofline  dominant kernel, algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline   the reference's CPU worker loop (oracle/cpu_pipeline.c restatement: libzstd + blake3, all host
                 threads) on a bounded sample of the same rows

`--impl reference` times only that CPU loop.  N>1 (torchrun): rows shard by range, one process per GPU, no
collective on the data path (weak scaling: every rank gets its own 2 GiB file).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SLICE = 8 << 20
PHRASE = b"The quick brown fox jumps over the lazy dog. "  # perf_bench.rs:74-80
METRIC = "decode+blake3-verify GB/s (device)"


# ----------------------------------------------------------------------------- corpus (input preparation only)
def _libzstd():
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_isError.argtypes = [C.c_size_t]
    return z


def _zstd_compress(z, a: np.ndarray, level: int) -> bytes:
    cap = z.ZSTD_compressBound(a.size)
    out = np.empty(cap, np.uint8)
    r = z.ZSTD_compress(out.ctypes.data, cap, a.ctypes.data, a.size, level)
    if z.ZSTD_isError(r):
        raise RuntimeError("ZSTD_compress failed")
    return out[:r].tobytes()


def _digest(a: np.ndarray) -> bytes:
    import blake3  # official bindings (same upstream code as the reference's blake3 crate); input preparation only
    return blake3.blake3(a.tobytes()).digest()


def text_slice(phase: int, n: int) -> np.ndarray:
    reps = (n + phase) // len(PHRASE) + 2
    return np.frombuffer(PHRASE * reps, np.uint8)[phase: phase + n].copy()


def build_text_corpus(total_bytes: int, level: int = 19, first_byte: int = 0):
    """Index rows of text(total_bytes) cut at 8 MiB: returns (blobs list, out_len list, digests (n,32))."""
    z = _libzstd()
    cache = {}
    blobs, lens, digs = [], [], []
    off = 0
    while off < total_bytes:
        n = min(SLICE, total_bytes - off)
        key = ((first_byte + off) % len(PHRASE), n)
        if key not in cache:
            s = text_slice(key[0], n)
            cache[key] = (_zstd_compress(z, s, level), _digest(s))
        b, d = cache[key]
        blobs.append(b)
        lens.append(n)
        digs.append(d)
        off += n
    return blobs, lens, np.frombuffer(b"".join(digs), np.uint8).reshape(-1, 32).copy()


def lcg_random(n: int) -> np.ndarray:
    """perf_bench.rs:86-92 exactly (val = val*6364136223846793005 + 1 from 12345; byte = val >> 33), vectorised with
    the closed form x_k = a^k x_0 + c (a^k - 1)/(a - 1) in wrapping uint64 arithmetic."""
    a, x = np.uint64(6364136223846793005), np.uint64(12345)
    out = np.empty(n, np.uint8)
    blk = 1 << 22
    with np.errstate(over="ignore"):
        A = np.cumprod(np.full(blk, a, np.uint64))            # a^1 .. a^blk
        S = np.concatenate([[np.uint64(1)], np.cumsum(A[:-1]) + np.uint64(1)])  # 1 + a + .. + a^(k-1), k = 1..blk
        for o in range(0, n, blk):
            m = min(blk, n - o)
            xs = A[:m] * x + S[:m]
            out[o:o + m] = (xs >> np.uint64(33)).astype(np.uint8)
            x = xs[m - 1]
    return out


def build_rows(entries, level_for):
    """entries: (path, data ndarray, skip) -> index rows cut at 8 MiB (stream_packer.rs:169-202)."""
    z = _libzstd()
    cache = {}
    blobs, lens, digs, comp = [], [], [], []
    for path, data, skip in entries:
        for o in range(0, max(len(data), 1), SLICE):
            sl = data[o:o + SLICE]
            key = (hash(sl[:4096].tobytes()), hash(sl[-4096:].tobytes()), len(sl), skip)
            if key not in cache:
                cache[key] = (sl.tobytes() if skip else _zstd_compress(z, np.ascontiguousarray(sl), level_for(path)), _digest(sl))
            b, d = cache[key]
            blobs.append(b); lens.append(len(sl)); digs.append(d); comp.append(0 if skip else 1)
    return blobs, lens, np.frombuffer(b"".join(digs), np.uint8).reshape(-1, 32).copy(), np.array(comp, np.uint8)


def text2g_desc(gib: float, rows: int) -> str:
    return (f"configs[1]: single {gib:g} GiB text-pattern file per GPU = {rows} rows x 8 MiB slices, zstd L19 frames "
            "(libzstd 1.5.5), decode + blake3 + 32-byte compare, output materialised in HBM")


def build_workload(name: str, gib: float, rank: int):
    """Returns (blobs, lens, digests, compressed flags, description)."""
    if name == "text2g":
        total = int(gib * (1 << 30))
        blobs, lens, digs = build_text_corpus(total, first_byte=rank * total)
        return blobs, lens, digs, np.ones(len(blobs), np.uint8), text2g_desc(gib, len(blobs))
    if name == "small100k":  # configs[0] shape: 100 000 x text(10 240) (perf_bench.rs:133-141), one row per file
        s = text_slice(0, 10240)
        b, d = _zstd_compress(_libzstd(), s, 19), _digest(s)
        n = 100_000
        return [b] * n, [10240] * n, np.tile(np.frombuffer(d, np.uint8), (n, 1)), np.ones(n, np.uint8), (
            "configs[0] corpus on the GPU: 100 000 x 10 KiB text files (977 MiB), one zstd L19 frame per file")
    if name == "mixed":  # configs[2]: 500 MiB incompressible + the mixed store-as-is set (perf_bench.rs:120-181)
        rnd = lcg_random(500 << 20)
        ents = [("random.bin", rnd, False), ("pom.xml", text_slice(0, 32 << 10), False),
                ("app.jar", rnd[:200 << 20], True), ("sources.jar", text_slice(0, 100 << 20), True),
                ("javadoc.jar", text_slice(0, 80 << 20), True), ("metadata.xml", text_slice(0, 16 << 10), False),
                ("deps.tar.gz", rnd[:150 << 20], True)]
        blobs, lens, digs, comp = build_rows(ents, lambda p: 1 if p == "random.bin" else 19)
        return blobs, lens, digs, comp, (
            f"configs[2]: 500 MiB incompressible (zstd frames of raw blocks) + mixed repo set, {len(blobs)} rows, "
            f"{int((comp == 0).sum())} store-as-is; decode/gather + blake3 + compare, output materialised in HBM")
    if name == "realtext":  # entropy-coded corpus: python stdlib sources, 8 MiB slices, zstd level 3
        import sysconfig
        root = sysconfig.get_paths()["stdlib"]
        parts, tot = [], 0
        want = int(gib * (1 << 30))
        for fn in sorted(os.listdir(root)):
            if fn.endswith(".py"):
                b = open(os.path.join(root, fn), "rb").read()
                parts.append(b); tot += len(b)
        data = np.frombuffer(b"".join(parts), np.uint8)
        data = np.resize(data, want)
        z = _libzstd()
        blobs, lens, digs = [], [], []
        for o in range(0, want, SLICE):
            sl = np.ascontiguousarray(data[o:o + SLICE])
            blobs.append(_zstd_compress(z, sl, 3)); lens.append(len(sl)); digs.append(_digest(sl))
        return blobs, lens, np.frombuffer(b"".join(digs), np.uint8).reshape(-1, 32).copy(), np.ones(len(blobs), np.uint8), (
            f"real text (python stdlib sources, {tot >> 20} MiB cycled to {gib:g} GiB), {len(blobs)} rows x 8 MiB, zstd level 3 "
            "frames: Huffman literals + FSE-described sequence tables")
    if name == "realsmall":  # the reference's only published entropy-coded corpus shape (backlog.md:322): ~41 k files of ~24 KB
        import sysconfig
        root = sysconfig.get_paths()["stdlib"]
        data = np.frombuffer(b"".join(open(os.path.join(root, fn), "rb").read() for fn in sorted(os.listdir(root))
                                      if fn.endswith(".py")), np.uint8)
        rng = np.random.default_rng(5)
        z = _libzstd()
        n, uniq = 40_000, 2_000  # 2 000 distinct files (compressing 40 000 would dominate the bench's set-up time), cycled
        files = []
        for _ in range(uniq):
            ln = int(rng.integers(2_000, 48_000))
            o = int(rng.integers(0, data.size - ln))
            sl = np.ascontiguousarray(data[o:o + ln])
            files.append((_zstd_compress(z, sl, 3), ln, _digest(sl)))
        blobs = [files[i % uniq][0] for i in range(n)]
        lens = [files[i % uniq][1] for i in range(n)]
        digs = np.frombuffer(b"".join(files[i % uniq][2] for i in range(n)), np.uint8).reshape(-1, 32).copy()
        return blobs, lens, digs, np.ones(n, np.uint8), (
            f"{n} real-text files of 2-48 KB ({sum(lens) >> 20} MiB, python sources), one zstd level-3 frame per file")
    raise SystemExit(f"unknown workload {name}")


def pack(blobs, align=16):
    offs, cur = [], 0
    for b in blobs:
        offs.append(cur)
        cur += (len(b) + align - 1) // align * align
    buf = np.zeros(max(cur, 1), np.uint8)
    for o, b in zip(offs, blobs):
        buf[o:o + len(b)] = np.frombuffer(b, np.uint8)
    return buf, np.array(offs, np.uint64)


# ----------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop, self.t = index, [], False, None

    def _loop_nvml(self, pynvml):
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        R = pynvml
        bits = [("hw_slowdown", getattr(R, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(R, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = get_reasons(h)
            self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits])
            time.sleep(0.002)

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            return self._loop_nvml(pynvml)
        except Exception:
            pass
        while not self.stop:
            try:
                r = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full summary
    of this same workload (profiles/r1_ncu_traffic.json); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- reference / cpu baseline arm
def cpu_pipeline(blobs, lens, digs, min_seconds: float, passes_cap: int = 100000, comp=None):
    """The reference's read worker loop on host cores (oracle/cpu_pipeline.c: libzstd + SIMD blake3, atomic row
    cursor, N = ceil(0.9*cores) threads as common_config.rs:34).  Returns (GB/s, threads, sample description)."""
    import oracle as O
    cores = os.cpu_count() or 1
    threads = max(1, -(-cores * 9 // 10))
    buf, offs = pack(blobs, 1)
    n = len(blobs)
    bs = np.array([len(b) for b in blobs], np.uint64)
    us = np.array(lens, np.uint64)
    cf = np.ones(n, np.uint8) if comp is None else np.ascontiguousarray(comp, np.uint8)
    O.decompress_rows(buf, offs[:2], bs[:2], np.zeros(2, np.uint64), cf[:2], us[:2], digs[:2], threads)  # warm
    t0, done, passes = time.perf_counter(), 0, 0
    while True:
        st = O.decompress_rows(buf, offs, bs, np.zeros(n, np.uint64), cf, us, digs, threads)
        assert st.corrupt_rows == 0 and st.decode_errors == 0 and st.verified_bytes == int(us.sum())
        done += int(us.sum())
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or passes >= passes_cap:
            break
    return done / dt / 1e9, threads, f"{n} rows ({int(us.sum()) >> 20} MiB) of the workload, {passes} passes, {dt:.1f} s", dt / passes


def static_config(wl_desc: str, n_rows: int, out_bytes: int) -> dict:
    """The part of the line both arms share verbatim (the driver compares it): what is processed, never how."""
    return {"workload": wl_desc, "rows_per_gpu": n_rows,
            "l2": f"working set {out_bytes >> 20} MiB per step > 126 MB L2, no flush needed"}


def run_reference(args):
    """The reference's CPU worker loop over the SAME rows as our arm (same builder, same config), all host threads; a step
    is one pass over the whole workload unless that would take more than a few seconds (then every k-th row, stated)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    blobs, lens, digs, comp, wl_desc = build_workload(args.workload, args.gib, 0)
    n, out_bytes = len(blobs), int(sum(lens))
    stride = max(1, out_bytes // (4 << 30))  # bound a step to ~4 GiB of CPU work
    sb, sl, sd, sc = blobs[::stride], lens[::stride], digs[::stride], comp[::stride]
    sample = (f"all {n} rows ({out_bytes >> 20} MiB) per step" if stride == 1 else
              f"every {stride}th row: {len(sb)} rows ({sum(sl) >> 20} MiB) per step")
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_pipeline(sb[:64], sl[:64], sd[:64], 0.0, 1, comp=sc[:64])
    t_tot, b_tot, threads = 0.0, 0, 0
    for _ in range(args.steps):
        _, threads, _, per_pass = cpu_pipeline(sb, sl, sd, 0.0, 1, comp=sc)
        t_tot += per_pass
        b_tot += sum(sl)
    v = b_tot / t_tot / 1e9
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * t_tot / args.steps, 3), "higher_is_better": True,
        "scaling": "strong" if args.workload == "multirepo" else "weak",
        "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
        "config": static_config(wl_desc, n, out_bytes),
        "note": "reference CPU worker loop (decompress.rs:105-192) restated in C over libzstd 1.5.5 + SIMD blake3, output "
                "materialised in host memory; the Rust reference itself cannot be built in this image (no cargo, OpenZL "
                "fetched at build time)",
        "cpu_baseline": {"value": round(v, 3), "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (znippy_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from znippy_b200 import Ctx, Plan, codec

    total = int(args.gib * (1 << 30))
    # weak scaling: rank r holds rows [r*256, (r+1)*256) of an N x 2 GiB multi-file archive, no exchange
    blobs, lens, digs, comp, wl_desc = build_workload(args.workload, args.gib, rank)
    n = len(blobs)
    in_buf, in_off = pack(blobs, 16)
    in_len = np.array([len(b) for b in blobs], np.uint64)
    out_len = np.array(lens, np.uint64)
    # output rows 16-byte aligned, as the host-buffer API and the archive loops lay them out
    out_off = np.concatenate([[0], np.cumsum((out_len + np.uint64(15)) & ~np.uint64(15))])[:-1].astype(np.uint64)
    out_bytes = int(out_len.sum())
    out_span = int(out_off[-1] + out_len[-1]) if n else 0

    ctx = Ctx(local, staging_bytes=out_span + in_buf.size + (1 << 20))
    stream = torch.cuda.Stream()  # a real (non-default) stream: kernels and the timing events share it
    torch.cuda.set_stream(stream)
    d_in = torch.from_numpy(in_buf).cuda()
    d_out = torch.empty(out_span + 256, dtype=torch.uint8, device="cuda")
    plan = Plan.decode_verify(ctx, in_off, in_len, comp, out_off, out_len, digs)
    plan.set_overlap(args.groups)  # decode of row range g+1 overlaps blake3 of range g (znippy_cuda.h)

    def step():
        plan.run(d_in.data_ptr(), d_out.data_ptr(), stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    st, dg = plan.results()
    assert not st.any(), f"warm-up produced non-OK statuses: {np.unique(st)}"
    assert (dg == digs).all()

    # ---- device-resident timing: K steps between events on the launching stream
    stage_ms = np.zeros(4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local) as clk:
        sync_all()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        sync_all()
        dev_ms = e0.elapsed_time(e1)
        # per-kernel times: the same step with the stages back to back on one stream (no overlap), each bracketed by
        # CUDA events on that stream inside zn_plan_run
        plan.set_overlap(1)
        for _ in range(args.steps):
            step()
            stage_ms += np.array(plan.last_ms())
        stage_ms /= args.steps
    clocks = clk.summary()
    st, _ = plan.results()
    assert not st.any()
    serial_launches = plan.launches()
    plan.set_overlap(args.groups)
    step()
    torch.cuda.synchronize()
    launches = plan.launches() * args.steps

    # ---- end to end through the reference-facing call, HOST buffers in pinned memory (zn_ctx_pinned), every step:
    #   verify  (the metric: `znippy verify` / decompress_archive(save_data=false), decompress.rs:135-184 without :186-189)
    #           H2D blobs + expected digests, kernels, D2H statuses + digests
    #   extract (save_data=true): additionally D2H of every decoded byte — PCIe-bound by construction
    pinned = ctx.pinned()
    h_in = pinned[: in_buf.size]
    h_in[:] = in_buf
    h_out = pinned[in_buf.size + 4096 - in_buf.size % 4096:][:out_span]
    e2e_steps = max(3, min(args.steps, 20))
    codec.decode_verify_batch(h_in, in_off, in_len, comp, out_len, digs, None, None, ctx)  # warm (allocations)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        est, edg = codec.decode_verify_batch(h_in, in_off, in_len, comp, out_len, digs, None, None, ctx)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert not est.any() and (edg == digs).all()
    x_steps = 3
    codec.decode_verify_batch(h_in, in_off, in_len, comp, out_len, digs, h_out, out_off, ctx)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(x_steps):
        est, _ = codec.decode_verify_batch(h_in, in_off, in_len, comp, out_len, digs, h_out, out_off, ctx)
    torch.cuda.synchronize()
    x_s = (time.perf_counter() - t0) / x_steps
    assert not est.any()
    if args.workload == "text2g":  # spot-check returned bytes
        assert bytes(h_out[:45]) == text_slice((rank * total) % len(PHRASE), 45).tobytes()

    # ---- reduce: max time over ranks
    times = torch.tensor([dev_ms, e2e_s * 1e3, x_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, x_ms = float(times[0]), float(times[1]), float(times[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    ms_per_step = dev_ms / args.steps
    value = world * out_bytes / (ms_per_step * 1e-3) / 1e9
    e2e_value = world * out_bytes / (e2e_ms * 1e-3) / 1e9
    x_value = world * out_bytes / (x_ms * 1e-3) / 1e9
    blob_bytes = int(in_len.sum())
    # dominant kernel: the largest stage
    dec = comp.astype(bool)
    k_decode_bytes = int(in_len[dec].sum() + out_len[dec].sum())  # blob read + uncompressed written
    k_hash_bytes = out_bytes + 32 * (out_bytes // 1024)  # content read + one 32 B chaining value per chunk written
    kernels = {
        "k_decode": {"ms": round(float(stage_ms[1]), 4), "alg_bytes": k_decode_bytes,
                     "gbs": round(k_decode_bytes / (stage_ms[1] * 1e-3) / 1e9, 1) if stage_ms[1] > 0 else None},
        "k_b3_chunks": {"ms": round(float(stage_ms[2]), 4), "alg_bytes": k_hash_bytes,
                        "gbs": round(k_hash_bytes / (stage_ms[2] * 1e-3) / 1e9, 1) if stage_ms[2] > 0 else None},
        "k_b3_tree": {"ms": round(float(stage_ms[3]), 4)}}
    dom = "k_b3_chunks" if stage_ms[2] >= stage_ms[1] else "k_decode"
    note = "blake3 is int-ALU bound (~10.5 int ops/byte), see DESIGN.md; hbm frac reported as asked"
    if plan.fused():
        # decode and hash are ONE kernel (fused_ws.cuh): algorithmic bytes of decode+verify per SURVEY §8(d) = blob read +
        # content written + one 32 B chaining value per KiB; the hash's read of the fresh output is not algorithmic
        fused_bytes = k_decode_bytes + 32 * (out_bytes // 1024)
        kernels = {"k_decode_ws": {"ms": round(float(stage_ms[1]), 4), "alg_bytes": fused_bytes,
                                   "gbs": round(fused_bytes / (stage_ms[1] * 1e-3) / 1e9, 1)},
                   "k_b3_tree": {"ms": round(float(stage_ms[3]), 4)}}
        dom = "k_decode_ws"
        note = ("decode + blake3 chunk hashing fused in one warp-specialised kernel; its floor is the int-ALU pipe (the "
                "standalone hash kernel needs 0.99 ms for this batch at 89 % ALU utilisation), not HBM; hbm frac reported as asked")
    achieved = kernels[dom]["gbs"]
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(dom) if args.workload == "text2g" else None,
                "peak_source": peak_src,
                "frac_of_nominal_8000": round(achieved / 8000.0, 4), "kernels": kernels, "note": note}

    cpu = None
    if not args.no_cpu:
        k = min(n, 64 if args.workload not in ("small100k", "realsmall") else 20000)
        gbs, threads, desc, _ = cpu_pipeline(blobs[:k], lens[:k], digs[:k], args.cpu_seconds, comp=comp[:k])
        cpu = {"value": round(gbs, 3), "unit": "GB/s", "cores": threads, "kind": "port", "sample": desc}

    print(json.dumps({
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
        "config": {"workload": wl_desc, "rows_per_gpu": n, "schedule": ("stages back to back on one stream" if args.groups <= 1 else
                                f"{args.groups} row groups, decode(g+1) overlaps blake3(g) on 2 streams"),
                   "serial_ms_per_step": round(float(stage_ms[0]), 4), "serial_launches_per_step": serial_launches,
                   "l2": f"working set {out_bytes >> 20} MiB per step > 126 MB L2, no flush needed"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": round(e2e_value, 3), "unit": "GB/s", "h2d_bytes_per_step": blob_bytes + 32 * n + 33 * n,
                "d2h_bytes_per_step": 36 * n, "steps": e2e_steps, "ms_per_step": round(e2e_ms, 4),
                "api": "zn_decode_verify_batch, verify-only (out_base=NULL): pinned host blobs + digests in, statuses + "
                       "digests out, timed with the host clock around the calls",
                "extract": {"value": round(x_value, 3), "unit": "GB/s", "d2h_bytes_per_step": out_bytes + 36 * n,
                            "steps": x_steps, "note": "save_data=true: every decoded byte returns over PCIe"}},
        "roofline": roofline, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="text2g", choices=["text2g", "small100k", "mixed", "realtext", "realsmall"],
                    help="text2g = BASELINE configs[1] (the metric's config); the others are secondary report lines")
    ap.add_argument("--gib", type=float, default=2.0, help="uncompressed GiB per GPU (2 = BASELINE configs[1])")
    ap.add_argument("--groups", type=int, default=1, help="row groups of the overlapped decode/hash schedule (1 = serial)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
