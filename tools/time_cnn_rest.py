"""Graph-replay time of the coefficient network's non-conv kernels + splat1 at small batch."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdrnet_b200 import _lib
lib = _lib.load()

def timed(call):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        call(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20): call()
        ts = []
        for _ in range(5):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) / 20 * 1e3)
    return statistics.median(ts)

cs = lambda: torch.cuda.current_stream().cuda_stream
for B in (1, 2, 8):
    out = []
    x = torch.randn(B, 256, 256, 3, device="cuda"); w = torch.randn(3, 3, 3, 8, device="cuda"); b = torch.randn(8, device="cuda")
    o = torch.empty(B, 128, 128, 8, device="cuda")
    out.append(("splat1", timed(lambda: lib.hdrnet_conv2d_nhwc_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), B, 256, 256, 3, 8, 3, 2, 1, cs()))))
    for name, I, O in (("fc1", 1024, 256), ("fc2", 256, 128), ("fc3", 128, 64)):
        xi = torch.randn(B, I, device="cuda"); wi = torch.randn(I, O, device="cuda"); bi = torch.randn(O, device="cuda"); oi = torch.empty(B, O, device="cuda")
        out.append((name, timed(lambda: lib.hdrnet_fc_f32(xi.data_ptr(), wi.data_ptr(), bi.data_ptr(), oi.data_ptr(), B, I, O, 1, cs()))))
    loc = torch.randn(B, 16, 16, 64, device="cuda"); gf = torch.randn(B, 64, device="cuda"); wp = torch.randn(64, 96, device="cuda"); bp = torch.randn(96, device="cuda")
    grid = torch.empty(B, 16, 16, 8, 3, 4, device="cuda")
    out.append(("fuse_predict", timed(lambda: lib.hdrnet_fuse_predict_f32(loc.data_ptr(), gf.data_ptr(), wp.data_ptr(), bp.data_ptr(), grid.data_ptr(), B, 16, 16, 64, 8, 3, 4, cs()))))
    print(f"batch {B}: " + "  ".join(f"{k} {v:.1f}" for k, v in out), flush=True)
