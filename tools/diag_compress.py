import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O
from znippy_b200 import codec
z=O.libzstd(); l=O.liblz4()
for name,d in [('real3.5m',O.real_text(3_500_000)),('real700k',O.real_text(700000)),('small',O.gen_small_alphabet(200000)),('text1m',O.gen_text(1<<20)),('bin8m',O.gen_binary(8<<20))]:
    b=codec.CompressCtx(3).compress(d); b4=codec.CompressCtx(3,codec.CODEC_LZ4).compress(d)
    assert z.decompress(b,len(d))==d.tobytes()
    rc,o,st=O.zstd_decompress(b,len(d),want_stats=True)
    print(name,len(d),'zstd',len(b),'lz4',len(b4),'libzstd1',len(z.compress(d,1)),'libzstd3',len(z.compress(d,3)),'liblz4',len(l.compress_frame(d)),{k:st[k] for k in ('blocks_raw','blocks_compressed','sequences','literal_bytes','lit_huf_4stream','lit_raw')})
