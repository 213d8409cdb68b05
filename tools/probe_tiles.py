"""Runs arcb200_tile_stream_probe: what does the layout of the trailing matrix (column-major tiles as
32 x 256 B segments vs 8 KB contiguous tiles) and the prefetch depth cost the symmetric product?"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arc_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
lda = 2368
buf = torch.randn(148 * lda * lda, dtype=torch.float64, device=dev)
scratch = torch.zeros(8, dtype=torch.float64, device=dev)
tf = ctypes.c_double(0.0)
_lib.check(lib.arcb200_dmma_peak(0, _lib.stream_ptr(0), _lib.ptr(scratch), ctypes.addressof(tf)), "peak")
out = {"dmma_tflops": tf.value, "lda": lda}
h = (ctypes.c_double * 3)()
for mode, name in ((0, "column_major_256B_segments"), (1, "tile_major_8KB"), (2, "column_major_block_barrier_per_step")):
    for depth in (1, 2):
        _lib.check(lib.arcb200_tile_stream_probe(0, _lib.stream_ptr(0), _lib.ptr(buf), buf.numel(), lda, 600, mode, depth,
                                                 tf.value, ctypes.addressof(h)), "probe")
        out["%s_depth%d" % (name, depth)] = {"ms": h[0], "frac_of_dmma_rate": h[1], "GBs": h[2]}
print(json.dumps(out))
