"""CPU-side view of the GPU zstd compressor (host emulation build of compress.cuh): sizes and block statistics."""
import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, oracle as O
L = C.CDLL(os.path.join('tests', 'host_emu', 'libhostemu.so'))
L.zn_hostemu_compress.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
L.zn_hostemu_compress.restype = C.c_long
z = O.libzstd()
for name, d in [('real3.5m', O.real_text(3_500_000)), ('real700k', O.real_text(700000)), ('small', O.gen_small_alphabet(200000)),
                ('text1m', O.gen_text(1 << 20)), ('bin8m', O.gen_binary(8 << 20))]:
    d = np.ascontiguousarray(d)
    out = np.zeros(d.size + d.size // 2 + 4096, np.uint8)
    n = L.zn_hostemu_compress(int(sys.argv[1]) if len(sys.argv) > 1 else 1, d.ctypes.data, d.size, out.ctypes.data)
    b = out[:n].tobytes()
    assert z.decompress(b, d.size) == d.tobytes()
    rc, o, st = O.zstd_decompress(b, d.size, want_stats=True)
    print(name, d.size, 'ours', n, 'libzstd1', len(z.compress(d, 1)), 'libzstd3', len(z.compress(d, 3)),
          {k: st[k] for k in st if k.startswith(('blocks', 'seq', 'lit', 'mode', 'rep'))})
