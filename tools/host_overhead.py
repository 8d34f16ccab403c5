"""Development probe: where the host-side time of zn_decode_verify_batch goes for 100 000 small rows."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from znippy_b200 import Ctx, codec, _native as N
blobs, lens, digs, comp, _ = bench.build_workload("small100k", 2, 0)
n = len(blobs)
in_buf, in_off = bench.pack(blobs, 16)
in_len = np.array([len(b) for b in blobs], np.uint64)
out_len = np.array(lens, np.uint64)
ctx = Ctx(0, staging_bytes=in_buf.size + (1 << 20))
h_in = ctx.pinned()[:in_buf.size]; h_in[:] = in_buf
for _ in range(3): codec.decode_verify_batch(h_in, in_off, in_len, comp, out_len, digs, None, None, ctx)
t0 = time.perf_counter()
for _ in range(10): codec.decode_verify_batch(h_in, in_off, in_len, comp, out_len, digs, None, None, ctx)
t_py = (time.perf_counter() - t0) / 10
st = np.zeros(n, np.uint32); dg = np.zeros((n, 32), np.uint8)
L = N.lib()
args = (ctx.handle, N.ptr(h_in), N.ptr(in_off), N.ptr(in_len), N.ptr(comp), N.ptr(out_len), N.ptr(digs.reshape(-1)), None, None, n, N.ptr(st), N.ptr(dg))
t0 = time.perf_counter()
for _ in range(10): L.zn_decode_verify_batch(*args)
t_c = (time.perf_counter() - t0) / 10
print(f"python wrapper {t_py*1e3:.2f} ms, C call alone {t_c*1e3:.2f} ms, kernels ~1.1 ms")
